#pragma once
#include <fiducial_msgs/FiducialMapEntry.h>
#include <vector>
namespace fiducial_msgs {
struct FiducialMapEntryArray {
    std::vector<FiducialMapEntry> fiducials;
};
}  // namespace fiducial_msgs
