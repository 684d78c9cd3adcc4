#pragma once
namespace fiducial_slam {
struct AddFiducial {
    struct Request {
        int fiducial_id = 0;
    };
    struct Response {};
};
}  // namespace fiducial_slam
