// C++ node glue over the C-ABI: what a maintainer drops into aruco_detect / fiducial_slam instead of
// the OpenCV calls.  Header-only, no ROS headers: the message structs below mirror fiducial_msgs
// field-for-field so that, inside a ROS build, filling the real messages is a member-wise copy
// (see INTEGRATION.md).  Method names follow the reference node
// (aruco_detect/src/aruco_detect.cpp, fiducial_slam/src/fiducial_slam.cpp).
#pragma once
#include <algorithm>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <fstream>
#include <sstream>
#include <cstdio>
#include <cmath>
#include <vector>

#include "../../include/fiducials_b200.h"

namespace fid_glue {

struct Header {  // std_msgs/Header
    uint32_t seq = 0;
    uint32_t stamp_sec = 0, stamp_nsec = 0;
    std::string frame_id;
};
struct Fiducial {  // fiducial_msgs/msg/Fiducial.msg:3-14
    int32_t fiducial_id = 0;
    int32_t direction = 0;
    double x0 = 0, y0 = 0, x1 = 0, y1 = 0, x2 = 0, y2 = 0, x3 = 0, y3 = 0;
};
struct FiducialArray {  // FiducialArray.msg
    Header header;
    std::vector<Fiducial> fiducials;
};
struct Transform {  // geometry_msgs/Transform
    double tx = 0, ty = 0, tz = 0;
    double qx = 0, qy = 0, qz = 0, qw = 1;
};
struct FiducialTransform {  // FiducialTransform.msg:2-6
    int32_t fiducial_id = 0;
    Transform transform;
    double image_error = 0, object_error = 0, fiducial_area = 0;
};
struct FiducialTransformArray {  // FiducialTransformArray.msg:3-5
    Header header;
    int32_t image_seq = 0;
    std::vector<FiducialTransform> transforms;
};
struct FiducialMapEntry {  // FiducialMapEntry.msg:2-10
    int32_t fiducial_id = 0;
    double x = 0, y = 0, z = 0, rx = 0, ry = 0, rz = 0;
};
struct FiducialMapEntryArray {
    std::vector<FiducialMapEntry> fiducials;
};

inline void check(int status, const char* what) {
    if (status != FID_OK) throw std::runtime_error(std::string(what) + ": " + fid_strerror(status));
}

// aruco_detect's FiducialsNode, minus ROS transport.
class FiducialsNode {
   public:
    FiducialsNode(int dictionary, double fiducial_len_, int max_width, int max_height, int device = 0) : fiducial_len(fiducial_len_) {
        fid_params p;
        check(fid_default_params(&p), "fid_default_params");  // aruco_detect.cpp:690-727
        p.dictionary = dictionary;                            // :611
        check(fid_create(&p, device, max_width, max_height, 1, &det), "fid_create");
    }
    ~FiducialsNode() {
        if (det) fid_destroy(det);
    }
    FiducialsNode(const FiducialsNode&) = delete;
    FiducialsNode& operator=(const FiducialsNode&) = delete;

    // camInfoCallback, aruco_detect.cpp:307-330
    void camInfoCallback(const double K[9], const double* D, int nD, const std::string& frame_id) {
        if (haveCamInfo) return;
        bool all_zero = true;
        for (int i = 0; i < 9; i++) all_zero = all_zero && K[i] == 0.0;
        if (all_zero) return;  // :313
        for (int i = 0; i < 9; i++) cam.K[i] = K[i];
        for (int i = 0; i < 5; i++) cam.D[i] = i < nD ? D[i] : 0.0;  // :317-323
        haveCamInfo = true;
        frameId = frame_id;
    }

    // imageCallback, aruco_detect.cpp:332-395: bgr = cv_bridge::toCvCopy(msg, BGR8) pixels (:348)
    bool imageCallback(const uint8_t* bgr, int width, int height, size_t stride, const Header& hdr, FiducialArray* fva) {
        if (!enable_detections) return false;  // :334
        fva->header = hdr;
        fva->header.frame_id = frameId;
        fva->fiducials.clear();
        ids.assign(FID_MAX_MARKERS, 0);
        corners.assign(FID_MAX_MARKERS * 8, 0.f);
        int n = 0;
        if (fid_detect(det, bgr, width, height, stride, FID_MAX_MARKERS, &n, ids.data(), corners.data()) != FID_OK) return false;  // frame dropped (:389-394)
        ids.resize(n);
        corners.resize((size_t)n * 8);
        for (int i = 0; i < n; i++) {
            if (std::count(ignoreIds.begin(), ignoreIds.end(), ids[i]) != 0) continue;  // :359-364
            Fiducial f;
            f.fiducial_id = ids[i];
            const float* c = &corners[(size_t)i * 8];
            f.x0 = c[0]; f.y0 = c[1]; f.x1 = c[2]; f.y1 = c[3]; f.x2 = c[4]; f.y2 = c[5]; f.x3 = c[6]; f.y3 = c[7];  // :366-376
            fva->fiducials.push_back(f);
        }
        last = hdr;
        return true;
    }

    // poseEstimateCallback, aruco_detect.cpp:397-538 (uses the member ids/corners like the reference)
    bool poseEstimateCallback(FiducialTransformArray* fta) {
        fta->header = last;
        fta->header.frame_id = frameId;
        fta->image_seq = (int32_t)last.seq;
        fta->transforms.clear();
        frameNum++;
        if (!doPoseEstimation) return true;
        if (!haveCamInfo) return false;  // :417-422
        std::vector<int32_t> oi;
        std::vector<double> ol;
        for (auto& kv : fiducialLens) {  // :239-244
            oi.push_back(kv.first);
            ol.push_back(kv.second);
        }
        std::vector<fid_transform> out(ids.size());
        if (fid_pose(det, (int)ids.size(), ids.data(), corners.data(), &cam, fiducial_len, (int)oi.size(), oi.data(), ol.data(), out.data()) != FID_OK) return true;
        for (const fid_transform& t : out) {
            if (std::count(ignoreIds.begin(), ignoreIds.end(), t.fiducial_id) != 0) continue;  // :440
            FiducialTransform ft;
            ft.fiducial_id = t.fiducial_id;
            ft.transform = Transform{t.translation[0], t.translation[1], t.translation[2], t.rotation[0], t.rotation[1], t.rotation[2], t.rotation[3]};
            ft.image_error = t.image_error;
            ft.object_error = t.object_error;
            ft.fiducial_area = t.fiducial_area;
            fta->transforms.push_back(ft);
        }
        return true;
    }

    double fiducial_len;
    bool doPoseEstimation = true, enable_detections = true, haveCamInfo = false;
    std::vector<int> ignoreIds;              // :540-571
    std::map<int, double> fiducialLens;      // :627-660
    std::string frameId;
    int frameNum = 0;

   private:
    fid_detector* det = nullptr;
    fid_camera cam{};
    std::vector<int32_t> ids;
    std::vector<float> corners;
    Header last;
};

// fiducial_slam's FiducialSlam + Map, minus ROS transport and tf (the two tf lookups of
// Map::updatePose, map.cpp:258-273, are passed in by the caller; nullptr = lookup failed).
class FiducialSlam {
   public:
    explicit FiducialSlam(int max_fiducials = 512, int device = 0) {
        fid_map_params p;
        check(fid_map_default_params(&p), "fid_map_default_params");
        p.max_fiducials = max_fiducials;
        check(fid_map_create(&p, device, &map), "fid_map_create");
        cap = max_fiducials;
    }
    ~FiducialSlam() {
        if (map) fid_map_destroy(map);
    }
    FiducialSlam(const FiducialSlam&) = delete;
    FiducialSlam& operator=(const FiducialSlam&) = delete;

    // transformCallback, fiducial_slam.cpp:79-105 + Map::update, map.cpp:152-176
    bool transformCallback(const FiducialTransformArray& msg, const fid_tf* T_baseCam, const fid_tf* T_camBase, fid_robot_pose* robot) {
        std::vector<fid_transform> obs(msg.transforms.size());
        for (size_t i = 0; i < obs.size(); i++) {
            const FiducialTransform& ft = msg.transforms[i];
            fid_transform& o = obs[i];
            o.fiducial_id = ft.fiducial_id;
            o.translation[0] = ft.transform.tx; o.translation[1] = ft.transform.ty; o.translation[2] = ft.transform.tz;
            o.rotation[0] = ft.transform.qx; o.rotation[1] = ft.transform.qy; o.rotation[2] = ft.transform.qz; o.rotation[3] = ft.transform.qw;
            o.image_error = ft.image_error; o.object_error = ft.object_error; o.fiducial_area = ft.fiducial_area;
        }
        return fid_map_update(map, 0, (int)obs.size(), obs.data(), T_baseCam, T_camBase, robot) == FID_OK;
    }

    // publishMap, map.cpp:629-654
    FiducialMapEntryArray publishMap() {
        std::vector<fid_map_entry> e(cap);
        int n = 0;
        check(fid_map_entries(map, 0, cap, &n, e.data()), "fid_map_entries");
        FiducialMapEntryArray out;
        for (int i = 0; i < n; i++) out.fiducials.push_back(FiducialMapEntry{e[i].fiducial_id, e[i].x, e[i].y, e[i].z, e[i].rx, e[i].ry, e[i].rz});
        return out;
    }

    // Map::saveMap, map.cpp:541-566
    bool saveMap(const std::string& filename) {
        std::vector<fid_map_entry> e(cap);
        int n = 0, np = 0;
        check(fid_map_entries(map, 0, cap, &n, e.data()), "fid_map_entries");
        std::vector<int32_t> pairs;
        fid_map_links(map, 0, 0, &np, nullptr);  // FID_ERR_CAPACITY by design: this call only asks for the count
        pairs.resize(2 * (size_t)np + 2);
        check(fid_map_links(map, 0, np, &np, pairs.data()), "fid_map_links");
        FILE* fp = fopen(filename.c_str(), "w");
        if (fp == NULL) return false;
        for (int i = 0; i < n; i++) {
            fprintf(fp, "%d %lf %lf %lf %lf %lf %lf %lf %d", e[i].fiducial_id, e[i].x, e[i].y, e[i].z, e[i].rx * 180.0 / M_PI, e[i].ry * 180.0 / M_PI, e[i].rz * 180.0 / M_PI,
                    e[i].variance, e[i].num_obs);
            for (int k = 0; k < np; k++)
                if (pairs[2 * k] == e[i].fiducial_id) fprintf(fp, " %d", pairs[2 * k + 1]);
            fprintf(fp, "\n");
        }
        fclose(fp);
        return true;
    }

    // Map::loadMap(filename), map.cpp:572-625: one fiducial per line, nine leading fields (id, x y z, roll pitch yaw
    // in degrees, variance, numObs), then the linked ids up to a tab or the end of the line; a line whose nine
    // fields do not parse is skipped (the reference logs "Invalid line").
    bool loadMap(const std::string& filename) {
        std::ifstream in(filename);
        if (!in) return false;
        std::vector<fid_map_file_entry> rows;
        std::vector<int32_t> pairs;
        std::string line;
        while (std::getline(in, line)) {
            // sscanf("%d %lf %lf %lf %lf %lf %lf %lf %d%[^\t\n]"): the nine numbers may be separated by any white space (tabs
            // included); only the link list that follows them ends at the first tab
            std::istringstream fields(line);
            fid_map_file_entry r{};
            if (!(fields >> r.fiducial_id >> r.x >> r.y >> r.z >> r.roll_deg >> r.pitch_deg >> r.yaw_deg >> r.variance >> r.num_obs)) continue;
            rows.push_back(r);
            std::string rest;
            std::getline(fields, rest);
            std::istringstream links(rest.substr(0, rest.find('\t')));
            for (int32_t linked; links >> linked;) {
                pairs.push_back(r.fiducial_id);
                pairs.push_back(linked);
            }
        }
        if (fid_map_load(map, 0, (int)rows.size(), rows.data()) != FID_OK) return false;
        return fid_map_add_links(map, 0, (int)pairs.size() / 2, pairs.data()) == FID_OK;
    }

   private:
    fid_map* map = nullptr;
    int cap = 0;
};

}  // namespace fid_glue
