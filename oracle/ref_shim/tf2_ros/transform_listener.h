// tf2_ros::Buffer stand-in: look-ups come from a table the harness fills; a missing entry throws like the real buffer.
#pragma once
#include <geometry_msgs/TransformStamped.h>
#include <map>
#include <ros/ros.h>
#include <tf2/exceptions.h>
namespace tf2_ros {
struct TfTable {
    std::map<std::pair<std::string, std::string>, geometry_msgs::Transform> t;  // (target, source) -> transform
    static TfTable& get() {
        static TfTable x;
        return x;
    }
};
class Buffer {
   public:
    explicit Buffer(ros::Duration = ros::Duration(10.0)) {}
    geometry_msgs::TransformStamped lookupTransform(const std::string& target, const std::string& source, const ros::Time& time, const ros::Duration = ros::Duration(0.0)) const {
        auto it = TfTable::get().t.find(std::make_pair(target, source));
        if (it == TfTable::get().t.end()) throw tf2::LookupException("no transform from " + source + " to " + target);
        geometry_msgs::TransformStamped out;
        out.header.frame_id = target;
        out.header.stamp = time;
        out.child_frame_id = source;
        out.transform = it->second;
        return out;
    }
};
class TransformListener {
   public:
    explicit TransformListener(Buffer&) {}
};
}  // namespace tf2_ros
