#pragma once
#include <ros/ros.h>
namespace std_msgs {
struct Header {
    unsigned int seq = 0;
    ros::Time stamp;
    std::string frame_id;
};
}  // namespace std_msgs
