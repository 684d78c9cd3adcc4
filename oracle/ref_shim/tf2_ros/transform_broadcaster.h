#pragma once
#include <geometry_msgs/TransformStamped.h>
#include <ros/ros.h>
namespace tf2_ros {
class TransformBroadcaster {
   public:
    void sendTransform(const geometry_msgs::TransformStamped& t) { ros::Capture::get().put(t); }
};
}  // namespace tf2_ros
