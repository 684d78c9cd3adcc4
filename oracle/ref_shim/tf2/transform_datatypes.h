// tf2::Stamped stand-in: a value with a time stamp and a frame.
#pragma once
#include <ros/ros.h>
#include <string>
namespace tf2 {
template <typename T>
class Stamped : public T {
   public:
    ros::Time stamp_;
    std::string frame_id_;
    Stamped() : frame_id_("NO_ID_STAMPED_DEFAULT_CONSTRUCTION") {}
    Stamped(const T& input, const ros::Time& timestamp, const std::string& frame_id) : T(input), stamp_(timestamp), frame_id_(frame_id) {}
    Stamped(const Stamped<T>& s) : T(s), stamp_(s.stamp_), frame_id_(s.frame_id_) {}
    Stamped& operator=(const Stamped<T>& s) {
        T::operator=(s);
        stamp_ = s.stamp_;
        frame_id_ = s.frame_id_;
        return *this;
    }
    void setData(const T& input) { *static_cast<T*>(this) = input; }
};
}  // namespace tf2
