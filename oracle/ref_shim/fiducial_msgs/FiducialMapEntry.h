#pragma once
namespace fiducial_msgs {
struct FiducialMapEntry {
    int fiducial_id = 0;
    double x = 0, y = 0, z = 0, rx = 0, ry = 0, rz = 0;
};
}  // namespace fiducial_msgs
