// message <-> tf2 conversions the reference uses (found through argument-dependent look-up, as with the real package)
#pragma once
#include <geometry_msgs/Point.h>
#include <geometry_msgs/PoseWithCovarianceStamped.h>
#include <geometry_msgs/TransformStamped.h>
#include <tf2/LinearMath/Transform.h>
#include <tf2/exceptions.h>
#include <tf2/transform_datatypes.h>
namespace tf2 {
inline void fromMsg(const geometry_msgs::Transform& in, Transform& out) {
    out = Transform(Quaternion(in.rotation.x, in.rotation.y, in.rotation.z, in.rotation.w), Vector3(in.translation.x, in.translation.y, in.translation.z));
}
inline geometry_msgs::Quaternion toMsg(const Quaternion& in) {
    geometry_msgs::Quaternion out;
    out.w = in.getW();
    out.x = in.getX();
    out.y = in.getY();
    out.z = in.getZ();
    return out;
}
inline geometry_msgs::Pose& toMsg(const Transform& in, geometry_msgs::Pose& out) {
    out.position.x = in.getOrigin().getX();
    out.position.y = in.getOrigin().getY();
    out.position.z = in.getOrigin().getZ();
    out.orientation = toMsg(in.getRotation());
    return out;
}
inline geometry_msgs::Transform toMsg(const Transform& in) {
    geometry_msgs::Transform out;
    out.translation.x = in.getOrigin().getX();
    out.translation.y = in.getOrigin().getY();
    out.translation.z = in.getOrigin().getZ();
    out.rotation = toMsg(in.getRotation());
    return out;
}
}  // namespace tf2
