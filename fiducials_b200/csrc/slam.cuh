// fiducial_slam per-message update, double precision, one map instance per thread.
//
// Follows the reference line by line (it is a sequential, order-dependent scalar-variance fold --
// SURVEY.md fact 6 / Appendix B), with tf2 LinearMath semantics restated:
//   FiducialSlam::transformCallback  fiducial_slam/src/fiducial_slam.cpp:79-105
//   Observation::Observation         fiducial_slam/src/map.cpp:53-59
//   Map::update                      map.cpp:152-176
//   Map::updateMap                   map.cpp:181-225
//   Map::updatePose                  map.cpp:247-391 (arithmetic :275-320, :347)
//   Map::autoInit, findClosestObs    map.cpp:415-485
//   TransformWithVariance::update    fiducial_slam/src/transform_with_variance.cpp:43-78
//   normalizeDavid / probabiltyAtPoint  transform_with_variance.cpp:14-38
//   operator*= (variances add)       include/fiducial_slam/transform_with_variance.h:26-34
#pragma once
#include "common.cuh"

namespace fid {

struct Twv {  // TransformWithVariance: tf2::Transform (row-major basis + origin) + scalar variance
    double R[9];
    double t[3];
    double var;
};

FID_HD void q_to_m(const double q[4], double m[9]) {  // tf2::Matrix3x3::setRotation
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double d = x * x + y * y + z * z + w * w;
    const double s = 2.0 / d;
    const double xs = x * s, ys = y * s, zs = z * s;
    const double wx = w * xs, wy = w * ys, wz = w * zs;
    const double xx = x * xs, xy = x * ys, xz = x * zs;
    const double yy = y * ys, yz = y * zs, zz = z * zs;
    m[0] = 1.0 - (yy + zz);
    m[1] = xy - wz;
    m[2] = xz + wy;
    m[3] = xy + wz;
    m[4] = 1.0 - (xx + zz);
    m[5] = yz - wx;
    m[6] = xz - wy;
    m[7] = yz + wx;
    m[8] = 1.0 - (xx + yy);
}

FID_HD void m_to_q(const double m[9], double q[4]) {  // tf2::Matrix3x3::getRotation
    const double tr = m[0] + m[4] + m[8];
    if (tr > 0.0) {
        double s = sqrt(tr + 1.0);
        q[3] = s * 0.5;
        s = 0.5 / s;
        q[0] = (m[7] - m[5]) * s;
        q[1] = (m[2] - m[6]) * s;
        q[2] = (m[3] - m[1]) * s;
    } else {
        const int i = m[0] < m[4] ? (m[4] < m[8] ? 2 : 1) : (m[0] < m[8] ? 2 : 0);
        const int j = (i + 1) % 3, k = (i + 2) % 3;
        double s = sqrt(m[i * 3 + i] - m[j * 3 + j] - m[k * 3 + k] + 1.0);
        q[i] = s * 0.5;
        s = 0.5 / s;
        q[3] = (m[k * 3 + j] - m[j * 3 + k]) * s;
        q[j] = (m[j * 3 + i] + m[i * 3 + j]) * s;
        q[k] = (m[k * 3 + i] + m[i * 3 + k]) * s;
    }
}

FID_HD void get_rpy(const double m[9], double* roll, double* pitch, double* yaw) {  // tf2::Matrix3x3::getRPY
    if (fabs(m[6]) >= 1.0) {
        *yaw = 0.0;
        const double delta = atan2(m[7], m[8]);
        if (m[6] < 0) {
            *pitch = 3.14159265358979323846 / 2.0;
            *roll = delta;
        } else {
            *pitch = -3.14159265358979323846 / 2.0;
            *roll = delta;
        }
        return;
    }
    *pitch = -asin(m[6]);
    const double c = cos(*pitch);
    *roll = atan2(m[7] / c, m[8] / c);
    *yaw = atan2(m[3] / c, m[0] / c);
}

FID_HD Twv twv_mul(const Twv& a, const Twv& b) {  // operator*=: compose, variances add
    Twv o;
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) o.R[i * 3 + j] = a.R[i * 3] * b.R[j] + a.R[i * 3 + 1] * b.R[3 + j] + a.R[i * 3 + 2] * b.R[6 + j];
        o.t[i] = (a.R[i * 3] * b.t[0] + a.R[i * 3 + 1] * b.t[1] + a.R[i * 3 + 2] * b.t[2]) + a.t[i];
    }
    o.var = a.var + b.var;
    return o;
}

FID_HD Twv twv_inverse(const Twv& a) {  // tf2::Transform::inverse
    Twv o;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) o.R[i * 3 + j] = a.R[j * 3 + i];
    const double n[3] = {-a.t[0], -a.t[1], -a.t[2]};
    for (int i = 0; i < 3; i++) o.t[i] = o.R[i * 3] * n[0] + o.R[i * 3 + 1] * n[1] + o.R[i * 3 + 2] * n[2];
    o.var = a.var;
    return o;
}

FID_HD double prob_at_point(double x, double u, double var) {
    return (1.0 / (sqrt(var) * sqrt(2.0 * 3.14159265358979323846))) * exp(-((x - u) * (x - u)) / (2.0 * var));
}

FID_HD double normalize_david(double new_mean, double mean1, double var1, double mean2, double var2) {
    const double p1 = prob_at_point(new_mean, mean1, var1);
    const double p2 = prob_at_point(new_mean, mean2, var2);
    const double p = sqrt(p1 * p1 + p2 * p2);
    const double r = 1.0 / (p * sqrt(2.0 * 3.14159265358979323846));
    double nv = r * r;
    nv = nv < 1e3 ? nv : 1e3;  // std::min(newVar, 1e3) -- also maps +inf; NaN falls through like std::min
    nv = nv > 1e-8 ? nv : 1e-8;
    return nv;
}

FID_HD void twv_update(Twv& self, const Twv& n) {  // TransformWithVariance::update
    double q1[4], q2[4];
    m_to_q(self.R, q1);
    m_to_q(n.R, q2);
    const double v1 = self.var, v2 = n.var;
    const double p1[3] = {self.t[0], self.t[1], self.t[2]};
    const double k = v1 / (v1 + v2);
    for (int i = 0; i < 3; i++) self.t[i] = p1[i] + k * (n.t[i] - p1[i]);
    // tf2::Quaternion::slerp
    const double dot = q1[0] * q2[0] + q1[1] * q2[1] + q1[2] * q2[2] + q1[3] * q2[3];
    const double s = sqrt((q1[0] * q1[0] + q1[1] * q1[1] + q1[2] * q1[2] + q1[3] * q1[3]) * (q2[0] * q2[0] + q2[1] * q2[1] + q2[2] * q2[2] + q2[3] * q2[3]));
    double c = (dot < 0 ? -dot : dot) / s;
    c = c < -1.0 ? -1.0 : (c > 1.0 ? 1.0 : c);
    const double theta = acos(c);
    double q[4];
    if (theta != 0.0) {
        const double d = 1.0 / sin(theta);
        const double s0 = sin((1.0 - k) * theta), s1 = sin(k * theta);
        for (int i = 0; i < 4; i++) q[i] = dot < 0 ? (q1[i] * s0 + -q2[i] * s1) * d : (q1[i] * s0 + q2[i] * s1) * d;
    } else {
        for (int i = 0; i < 4; i++) q[i] = q1[i];
    }
    const double nrm = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; i++) q[i] /= nrm;
    q_to_m(q, self.R);
    const double d0 = n.t[0] - p1[0], d1 = n.t[1] - p1[1], d2 = n.t[2] - p1[2];
    const double mean2 = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    const double e0 = self.t[0] - p1[0], e1 = self.t[1] - p1[1], e2 = self.t[2] - p1[2];
    const double mean = sqrt(e0 * e0 + e1 * e1 + e2 * e2);
    self.var = normalize_david(mean, 0.0, v1, mean2, v2);
}

// ---- map state -------------------------------------------------------------------------------------
struct MapEntry {
    int32_t id;
    int32_t num_obs;
    Twv pose;
};

struct MapState {  // Map members map.h:118-134 that carry arithmetic state
    int32_t n;           // number of fiducials
    int32_t capacity;
    int32_t frame_num;
    int32_t initial_frame_num;
    int32_t origin_fid;
    int32_t initializing;
    int32_t read_only;
    int32_t overflow;    // set if an insertion did not fit
    int32_t hash_size;   // power of two >= 2*capacity; 0 = no hash table (linear search)
    int32_t hash_valid;  // 0 = rebuild the id -> slot table from the entries before the next update
    int32_t fiducial_to_add;     // add_fiducial service request, -1 = none (Map::fiducialToAdd, map.cpp:821-828)
    int32_t add_have_map_base;   // the map -> base tf lookup of handleAddFiducial (map.cpp:514-517) succeeded
    double add_map_base[7];      // ... x y z qx qy qz qw
};

// id -> slot open-addressing table (linear probing); key = id + 1 (0 = empty), value = slot.
struct MapHash {
    int32_t* keys;
    int32_t* vals;
    int32_t size;
};

FID_HD uint32_t map_hash_of(int id) { return (uint32_t)id * 2654435761u; }

FID_HD void map_hash_insert(const MapHash& h, int id, int slot) {
    uint32_t p = map_hash_of(id) & (uint32_t)(h.size - 1);
    while (h.keys[p] != 0 && h.keys[p] != id + 1) p = (p + 1) & (uint32_t)(h.size - 1);
    h.keys[p] = id + 1;
    h.vals[p] = slot;
}

FID_HD void map_hash_rebuild(const MapHash& h, const struct MapEntry* e, int n);

struct Obs {  // one FiducialTransform as consumed by transformCallback
    int32_t id;
    int32_t pad;
    double t[3];
    double q[4];  // x y z w
    double object_error;
    double area;
};

struct RobotPose {  // T_mapBase after updatePose
    int32_t valid;
    int32_t n_estimates;
    double t[3];
    double q[4];
    double var;
};


FID_HD void map_hash_rebuild(const MapHash& h, const MapEntry* e, int n) {
    for (int i = 0; i < h.size; i++) h.keys[i] = 0;
    for (int i = 0; i < n; i++) map_hash_insert(h, e[i].id, i);
}

FID_HD int map_find(const MapState& st, const MapEntry* e, int id, const MapHash* h = nullptr) {
    if (h && h->size > 0) {
        uint32_t p = map_hash_of(id) & (uint32_t)(h->size - 1);
        while (h->keys[p] != 0) {
            if (h->keys[p] == id + 1) return h->vals[p];
            p = (p + 1) & (uint32_t)(h->size - 1);
        }
        return -1;
    }
    for (int i = 0; i < st.n; i++)
        if (e[i].id == id) return i;
    return -1;
}

FID_HD bool isnan3(const double t[3]) { return t[0] != t[0] || t[1] != t[1] || t[2] != t[2]; }

// T_camFid of one observation with its variance (Observation::Observation, map.cpp:53-59; transformCallback :91-97)
FID_HD Twv obs_cam_fid(const Obs& o, double weighting_scale, int use_area) {
    Twv t;
    q_to_m(o.q, t.R);
    t.t[0] = o.t[0];
    t.t[1] = o.t[1];
    t.t[2] = o.t[2];
    t.var = use_area ? weighting_scale / o.area : weighting_scale * o.object_error;
    return t;
}

// links: capacity x capacity bit matrix (row = slot, bit = other slot), may be null.
// var_scratch / slot_scratch: n_obs entries each of caller-owned scratch (the per-observation variance updatePose writes back
// for updateMap, map.cpp:298, and the slots of the link pass) -- a message may hold any number of observations, like the
// reference's std::vector<Observation>.
FID_HD void map_update_core(MapState& st, MapEntry* e, uint32_t* links, const Obs* obs_in, int n_obs, const Twv* T_baseCam /*null = lookup failed*/,
                            const Twv* T_camBase /*null = lookup failed*/, double weighting_scale, int use_area, double systematic_error, RobotPose* robot,
                            const MapHash* hash, double* var_scratch, int* slot_scratch) {
    if (hash && hash->size > 0 && !st.hash_valid) {
        map_hash_rebuild(*hash, e, st.n);
        st.hash_valid = 1;
    }
    robot->valid = 0;
    robot->n_estimates = 0;
    // transformCallback: observations with T_camFid, T_fidCam, variance (recomputed from the message where needed)
    st.frame_num++;
    if (n_obs > 0 && st.n == 0) st.initializing = 1;
    if (st.initializing) {
        // ---- autoInit
        if (st.n == 0) {
            int idx = -1;
            double smallest = -1.0;
            for (int i = 0; i < n_obs; i++) {
                const double* ct = obs_in[i].t;
                const double d = ct[0] * ct[0] + ct[1] * ct[1] + ct[2] * ct[2];
                if (smallest < 0 || d < smallest) {
                    smallest = d;
                    idx = i;
                }
            }
            if (idx >= 0 && st.capacity > 0) {
                st.origin_fid = obs_in[idx].id;
                const Twv cf = obs_cam_fid(obs_in[idx], weighting_scale, use_area);
                Twv T = cf;
                if (T_baseCam) {
                    T = twv_mul(*T_baseCam, cf);
                    T.var = cf.var;
                }
                e[0].id = obs_in[idx].id;
                e[0].num_obs = 0;
                e[0].pose = T;
                st.n = 1;
                if (hash && hash->size > 0) map_hash_insert(*hash, obs_in[idx].id, 0);
            } else if (idx >= 0) {
                st.overflow = 1;
            }
        } else {
            for (int i = 0; i < n_obs; i++) {
                if (obs_in[i].id == st.origin_fid) {
                    const Twv cf = obs_cam_fid(obs_in[i], weighting_scale, use_area);
                    Twv T = cf;
                    if (T_baseCam) {
                        T = twv_mul(*T_baseCam, cf);
                        T.var = cf.var;
                    }
                    const int slot = map_find(st, e, st.origin_fid, hash);
                    if (slot >= 0) {
                        twv_update(e[slot].pose, T);
                        e[slot].num_obs++;
                    }
                    break;
                }
            }
        }
        if (st.frame_num - st.initial_frame_num > 10 && st.origin_fid != -1) {
            st.initializing = 0;
            const int slot = map_find(st, e, st.origin_fid, hash);
            if (slot >= 0) e[slot].pose.var = 0.0;
        }
        return;
    }
    // ---- updatePose
    if (n_obs == 0 || !T_baseCam) return;
    Twv camBase;
    if (T_camBase) {
        camBase = *T_camBase;
        camBase.var = 1.0;
    } else {  // default-constructed tf2::Transform is uninitialised in the reference; use identity
        for (int i = 0; i < 9; i++) camBase.R[i] = (i % 4 == 0) ? 1.0 : 0.0;
        camBase.t[0] = camBase.t[1] = camBase.t[2] = 0.0;
        camBase.var = 0.0;
    }
    Twv baseCam = *T_baseCam;
    baseCam.var = 1.0;
    int n_est = 0;
    Twv mapBase;
    for (int i = 0; i < n_obs; i++) {
        const Twv cf = obs_cam_fid(obs_in[i], weighting_scale, use_area);
        var_scratch[i] = cf.var;
        const int slot = map_find(st, e, obs_in[i].id, hash);
        if (slot < 0) continue;
        Twv fidCam = twv_inverse(cf);
        Twv p = twv_mul(e[slot].pose, fidCam);
        p = twv_mul(p, camBase);
        double roll, pitch, yaw;
        get_rpy(p.R, &roll, &pitch, &yaw);
        const double* c = cf.t;
        const double zr = p.t[2] / c[2];
        const double s1 = (zr * zr) * (c[0] * c[0] + c[1] * c[1]);
        const double len2 = p.t[0] * p.t[0] + p.t[1] * p.t[1] + p.t[2] * p.t[2];
        const double sr = sin(roll), spt = sin(pitch);
        const double s2 = len2 * (sr * sr);
        const double s3 = len2 * (spt * spt);
        p.var = s1 + s2 + s3 + systematic_error;
        var_scratch[i] = p.var;  // write-back used by updateMap
        if (isnan3(p.t)) continue;
        if (n_est == 0) {
            mapBase = p;
        } else {
            twv_update(mapBase, p);
        }
        n_est++;
    }
    if (n_est == 0) return;
    robot->valid = 1;
    robot->n_estimates = n_est;
    robot->t[0] = mapBase.t[0];
    robot->t[1] = mapBase.t[1];
    robot->t[2] = mapBase.t[2];
    m_to_q(mapBase.R, robot->q);
    robot->var = mapBase.var;
    const Twv mapCam = twv_mul(mapBase, baseCam);
    // ---- updateMap
    if (n_obs > 1 && !st.read_only) {
        for (int i = 0; i < n_obs; i++) {
            Twv cf = obs_cam_fid(obs_in[i], weighting_scale, use_area);
            cf.var = var_scratch[i];
            const Twv mapFid = twv_mul(mapCam, cf);
            if (isnan3(mapFid.t)) continue;
            int slot = map_find(st, e, obs_in[i].id, hash);
            if (slot < 0) {
                if (st.n >= st.capacity) {
                    st.overflow = 1;
                    continue;
                }
                slot = st.n++;
                e[slot].id = obs_in[i].id;
                e[slot].num_obs = 0;
                e[slot].pose = mapFid;
                if (hash && hash->size > 0) map_hash_insert(*hash, obs_in[i].id, slot);
            }
            if (e[slot].pose.var != 0) {
                twv_update(e[slot].pose, mapFid);
                e[slot].num_obs += 2;
            }
        }
        // links (map.cpp:217-222): every fiducial seen in this frame links to every other one
        if (links) {
            const int wpr = (st.capacity + 31) / 32;
            int* slots = slot_scratch;
            for (int i = 0; i < n_obs; i++) slots[i] = map_find(st, e, obs_in[i].id, hash);
            for (int i = 0; i < n_obs; i++) {
                if (slots[i] < 0) continue;
                for (int j = 0; j < n_obs; j++) {
                    if (obs_in[j].id == obs_in[i].id || slots[j] < 0) continue;
                    links[(size_t)slots[i] * wpr + (slots[j] >> 5)] |= 1u << (slots[j] & 31);
                }
            }
        }
    }
}

// Map::handleAddFiducial, map.cpp:489-535 (called at the end of every Map::update, :173)
FID_HD void map_handle_add_fiducial(MapState& st, MapEntry* e, const Obs* obs_in, int n_obs, const Twv* T_baseCam, double weighting_scale, int use_area,
                                    const MapHash* hash) {
    if (st.fiducial_to_add == -1) return;
    if (map_find(st, e, st.fiducial_to_add, hash) >= 0) {  // "already in map - ignoring add request"
        st.fiducial_to_add = -1;
        return;
    }
    for (int i = 0; i < n_obs; i++) {
        if (obs_in[i].id != st.fiducial_to_add) continue;
        Twv T = obs_cam_fid(obs_in[i], weighting_scale, use_area);
        const double var = T.var;
        if (T_baseCam) T = twv_mul(*T_baseCam, T);  // tf2::Transform * TransformWithVariance: the variance stays T's
        if (st.add_have_map_base) {
            Twv mb;
            q_to_m(st.add_map_base + 3, mb.R);
            mb.t[0] = st.add_map_base[0];
            mb.t[1] = st.add_map_base[1];
            mb.t[2] = st.add_map_base[2];
            mb.var = 0.0;
            T = twv_mul(mb, T);
        }
        T.var = var;
        if (st.n >= st.capacity) {
            st.overflow = 1;
        } else {
            const int slot = st.n++;
            e[slot].id = obs_in[i].id;
            e[slot].num_obs = 0;
            e[slot].pose = T;
            if (hash && hash->size > 0) map_hash_insert(*hash, obs_in[i].id, slot);
            // fiducials[originFid].pose.variance = 0.0 -- only where originFid names a fiducial (with originFid == -1 the
            // reference's operator[] inserts a default-constructed Fiducial with uninitialised members: not restated)
            const int os = st.origin_fid != -1 ? map_find(st, e, st.origin_fid, hash) : -1;
            if (os >= 0) e[os].pose.var = 0.0;
        }
        st.initializing = 0;
        st.fiducial_to_add = -1;
        return;
    }
}

FID_HD void map_update(MapState& st, MapEntry* e, uint32_t* links, const Obs* obs_in, int n_obs, const Twv* T_baseCam /*null = lookup failed*/,
                       const Twv* T_camBase /*null = lookup failed*/, double weighting_scale, int use_area, double systematic_error, RobotPose* robot,
                       const MapHash* hash, double* var_scratch, int* slot_scratch) {
    map_update_core(st, e, links, obs_in, n_obs, T_baseCam, T_camBase, weighting_scale, use_area, systematic_error, robot, hash, var_scratch, slot_scratch);
    map_handle_add_fiducial(st, e, obs_in, n_obs, T_baseCam, weighting_scale, use_area, hash);
}

}  // namespace fid
