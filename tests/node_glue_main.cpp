// Compile/link check (and, on a GPU box, a run) of the C++ node glue: reads a raw BGR8 frame
// written by the python test, runs imageCallback + poseEstimateCallback + transformCallback and
// prints the messages as text for the python side to compare with the oracle.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../fiducials_b200/csrc/node_glue.hpp"

int main(int argc, char** argv) {
    if (argc < 6) {
        std::fprintf(stderr, "usage: %s frame.bgr width height dictionary fiducial_len [fx fy cx cy k1 k2 p1 p2 k3]\n", argv[0]);
        return 2;
    }
    const int W = std::atoi(argv[2]), H = std::atoi(argv[3]), dict = std::atoi(argv[4]);
    const double len = std::atof(argv[5]);
    std::vector<uint8_t> bgr((size_t)W * H * 3);
    FILE* f = std::fopen(argv[1], "rb");
    if (!f || std::fread(bgr.data(), 1, bgr.size(), f) != bgr.size()) return 3;
    std::fclose(f);
    try {
        fid_glue::FiducialsNode node(dict, len, W, H);
        double K[9] = {0.73 * W, 0, W / 2.0, 0, 0.73 * W, H / 2.0, 0, 0, 1};
        double D[5] = {0, 0, 0, 0, 0};
        if (argc >= 15) {
            K[0] = std::atof(argv[6]); K[4] = std::atof(argv[7]); K[2] = std::atof(argv[8]); K[5] = std::atof(argv[9]);
            for (int i = 0; i < 5; i++) D[i] = std::atof(argv[10 + i]);
        }
        node.camInfoCallback(K, D, 5, "camera");
        fid_glue::FiducialArray fva;
        fid_glue::FiducialTransformArray fta;
        fid_glue::Header hdr;
        hdr.seq = 7;
        if (!node.imageCallback(bgr.data(), W, H, (size_t)W * 3, hdr, &fva)) return 4;
        if (!node.poseEstimateCallback(&fta)) return 5;
        for (const auto& v : fva.fiducials) std::printf("V %d %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g\n", v.fiducial_id, v.x0, v.y0, v.x1, v.y1, v.x2, v.y2, v.x3, v.y3);
        for (const auto& t : fta.transforms)
            std::printf("T %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", t.fiducial_id, t.transform.tx, t.transform.ty, t.transform.tz, t.transform.qx,
                        t.transform.qy, t.transform.qz, t.transform.qw, t.image_error, t.object_error, t.fiducial_area);
        fid_glue::FiducialSlam slam(64);
        fid_tf ident{{0, 0, 0}, {0, 0, 0, 1}};
        fid_robot_pose robot{};
        for (int i = 0; i < 13; i++) slam.transformCallback(fta, &ident, &ident, &robot);
        for (const auto& e : slam.publishMap().fiducials) std::printf("M %d %.17g %.17g %.17g %.17g %.17g %.17g\n", e.fiducial_id, e.x, e.y, e.z, e.rx, e.ry, e.rz);
        std::printf("R %d %.17g %.17g %.17g\n", robot.valid, robot.t[0], robot.t[1], robot.t[2]);
        // published pose: covariance override + 2-D squashing of map -> odom (map.cpp:337-379)
        {
            fid_robot_pose rp{};
            rp.valid = 1;
            rp.t[0] = 1.25; rp.t[1] = -0.5; rp.t[2] = 0.3;
            rp.q[0] = 0.18257418583505536; rp.q[1] = 0.3651483716701107; rp.q[2] = 0.5477225575051661; rp.q[3] = 0.7302967433402214;
            rp.variance = 0.125;
            fid_tf odom{{0.4, 0.2, 0.0}, {0, 0, 0.3826834323650898, 0.9238795325112867}};
            double cov[36];
            slam.robotPoseCovariance(rp, cov);
            std::printf("C %.17g %.17g %.17g\n", cov[0], cov[7], cov[35]);
            slam.setCovarianceDiagonal({1, 2, 3, 4, 5, 6});
            slam.robotPoseCovariance(rp, cov);
            std::printf("C %.17g %.17g %.17g\n", cov[0], cov[7], cov[35]);
            slam.setCovarianceDiagonal({1, 2, 0, 4, 5, 6});
            slam.robotPoseCovariance(rp, cov);
            std::printf("C %.17g %.17g %.17g\n", cov[0], cov[7], cov[35]);
            fid_glue::Transform p2 = slam.poseTf(rp, &odom);
            std::printf("P %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", p2.tx, p2.ty, p2.tz, p2.qx, p2.qy, p2.qz, p2.qw);
            slam.publish_6dof_pose = true;
            fid_glue::Transform p6 = slam.poseTf(rp, nullptr);
            std::printf("P %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", p6.tx, p6.ty, p6.tz, p6.qx, p6.qy, p6.qz, p6.qw);
        }
        // saveMap / loadMap round trip through the reference's text format (map.cpp:541-625)
        const std::string path = std::string(argv[1]) + ".map.txt";
        if (!slam.saveMap(path)) return 6;
        fid_glue::FiducialSlam slam2(64);
        if (!slam2.loadMap(path) || !slam2.saveMap(path + "2")) return 7;
        std::printf("F %s\n", path.c_str());
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
