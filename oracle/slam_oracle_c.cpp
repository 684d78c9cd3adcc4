/* C restatement of the fiducial_slam per-message update, for timing the reference algorithm on the
 * host cores (bench.py cpu_baseline for config C5) and as a second checker.  TEST INFRASTRUCTURE ONLY.
 *
 * The arithmetic is shared verbatim with the CUDA path through fiducials_b200/csrc/slam.cuh, whose
 * functions are plain C++ when compiled by g++ (FID_HD expands to `inline`); this file only adds a C
 * ABI around them.  slam.cuh follows fiducial_slam/src/map.cpp:152-320,415-485 and
 * transform_with_variance.cpp:9-85 line by line and is itself pinned to the numpy restatement
 * (oracle/slam_oracle.py) by tests/test_hostsim_detect.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../fiducials_b200/csrc/slam.cuh"

using namespace fid;

extern "C" {

/* Replays n_msgs messages into a fresh map; obs: rows of 10 doubles (id,t3,q4,object_error,area);
 * offsets[n_msgs+1]; seed entries: rows of (id,x,y,z,qx,qy,qz,qw,var).  Returns number of fiducials;
 * entries_out rows (id, t3, R9 row-major, var). */
int slam_c_replay(int capacity, int n_seed, const double* seed, int n_msgs, const int32_t* offsets, const double* obs, const double* base_cam, const double* cam_base,
                  double* entries_out, double* robot_out) {
    MapState st;
    memset(&st, 0, sizeof(st));
    st.capacity = capacity;
    st.origin_fid = -1;
    st.fiducial_to_add = -1;
    MapEntry* e = (MapEntry*)calloc((size_t)capacity, sizeof(MapEntry));
    for (int i = 0; i < n_seed; i++) {
        e[i].id = (int)seed[9 * i];
        e[i].num_obs = 0;
        q_to_m(seed + 9 * i + 4, e[i].pose.R);
        e[i].pose.t[0] = seed[9 * i + 1];
        e[i].pose.t[1] = seed[9 * i + 2];
        e[i].pose.t[2] = seed[9 * i + 3];
        e[i].pose.var = seed[9 * i + 8];
    }
    st.n = n_seed;
    Twv bc, cb;
    if (base_cam) {
        q_to_m(base_cam + 3, bc.R);
        bc.t[0] = base_cam[0];
        bc.t[1] = base_cam[1];
        bc.t[2] = base_cam[2];
        bc.var = 0;
    }
    if (cam_base) {
        q_to_m(cam_base + 3, cb.R);
        cb.t[0] = cam_base[0];
        cb.t[1] = cam_base[1];
        cb.t[2] = cam_base[2];
        cb.var = 0;
    }
    int max_n = 1;
    for (int k = 0; k < n_msgs; k++) max_n = offsets[k + 1] - offsets[k] > max_n ? offsets[k + 1] - offsets[k] : max_n;
    Obs* buf = (Obs*)calloc((size_t)max_n, sizeof(Obs));
    double* var_scratch = (double*)calloc((size_t)max_n, sizeof(double));
    int* slot_scratch = (int*)calloc((size_t)max_n, sizeof(int));
    RobotPose rp;
    for (int k = 0; k < n_msgs; k++) {
        int n = offsets[k + 1] - offsets[k];
        for (int i = 0; i < n; i++) {
            const double* o = obs + 10 * (size_t)(offsets[k] + i);
            buf[i].id = (int)o[0];
            buf[i].t[0] = o[1];
            buf[i].t[1] = o[2];
            buf[i].t[2] = o[3];
            buf[i].q[0] = o[4];
            buf[i].q[1] = o[5];
            buf[i].q[2] = o[6];
            buf[i].q[3] = o[7];
            buf[i].object_error = o[8];
            buf[i].area = o[9];
        }
        map_update(st, e, NULL, buf, n, base_cam ? &bc : NULL, cam_base ? &cb : NULL, 1e9, 0, 0.01, &rp, NULL, var_scratch, slot_scratch);
        if (robot_out) {
            double* r = robot_out + 9 * (size_t)k;
            r[0] = rp.valid;
            r[1] = rp.t[0];
            r[2] = rp.t[1];
            r[3] = rp.t[2];
            r[4] = rp.q[0];
            r[5] = rp.q[1];
            r[6] = rp.q[2];
            r[7] = rp.q[3];
            r[8] = rp.var;
        }
    }
    for (int i = 0; i < st.n; i++) {
        double* o = entries_out + 14 * (size_t)i;
        o[0] = e[i].id;
        for (int k = 0; k < 3; k++) o[1 + k] = e[i].pose.t[k];
        for (int k = 0; k < 9; k++) o[4 + k] = e[i].pose.R[k];
        o[13] = e[i].pose.var;
    }
    const int n = st.n;
    free(e);
    free(buf);
    free(var_scratch);
    free(slot_scratch);
    return n;
}
}
