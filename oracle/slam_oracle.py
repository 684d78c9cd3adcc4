"""CPU oracle for the fiducial_slam per-message path.  TEST INFRASTRUCTURE ONLY.

A double-precision restatement of

  * FiducialSlam::transformCallback   fiducial_slam/src/fiducial_slam.cpp:79-105
  * Observation::Observation          fiducial_slam/src/map.cpp:53-59
  * Map::update                       map.cpp:152-176
  * Map::updateMap                    map.cpp:181-225
  * Map::updatePose                   map.cpp:247-391 (arithmetic :275-320, :347)
  * Map::autoInit / findClosestObs    map.cpp:415-485
  * Map::publishMap                   map.cpp:629-654
  * TransformWithVariance             fiducial_slam/src/transform_with_variance.cpp:9-85 and
                                      include/fiducial_slam/transform_with_variance.h:26-59

with the semantics of tf2 LinearMath (ROS geometry2; not vendored in the reference):
Transform composition / inverse, Matrix3x3::setRotation(q) / getRotation / getRPY,
Quaternion::slerp / normalized / setRPY.  Pure python floats (IEEE double) so that the
operation order matches the C++ expression order; numpy only for containers.

tf lookups (map.cpp:258-273, :455, :470) are inputs: ``T_baseCam`` / ``T_camBase`` are
passed in by the caller, exactly as the C-ABI does (include/fiducials_b200.h).
"""
from __future__ import annotations

import math

import numpy as np
from dataclasses import dataclass, field

SYSTEMATIC_ERROR = 0.01  # map.cpp:50


# ----------------------------- tf2 LinearMath restatement ------------------------------
def q_to_m(q):
    """tf2::Matrix3x3::setRotation(q)."""
    x, y, z, w = q
    d = x * x + y * y + z * z + w * w
    s = 2.0 / d
    xs, ys, zs = x * s, y * s, z * s
    wx, wy, wz = w * xs, w * ys, w * zs
    xx, xy, xz = x * xs, x * ys, x * zs
    yy, yz, zz = y * ys, y * zs, z * zs
    return [
        [1.0 - (yy + zz), xy - wz, xz + wy],
        [xy + wz, 1.0 - (xx + zz), yz - wx],
        [xz - wy, yz + wx, 1.0 - (xx + yy)],
    ]


def m_to_q(m):
    """tf2::Matrix3x3::getRotation(q)."""
    tr = m[0][0] + m[1][1] + m[2][2]
    t = [0.0, 0.0, 0.0, 0.0]
    if tr > 0.0:
        s = math.sqrt(tr + 1.0)
        t[3] = s * 0.5
        s = 0.5 / s
        t[0] = (m[2][1] - m[1][2]) * s
        t[1] = (m[0][2] - m[2][0]) * s
        t[2] = (m[1][0] - m[0][1]) * s
    else:
        i = (2 if m[1][1] < m[2][2] else 1) if m[0][0] < m[1][1] else (2 if m[0][0] < m[2][2] else 0)
        j = (i + 1) % 3
        k = (i + 2) % 3
        s = math.sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0)
        t[i] = s * 0.5
        s = 0.5 / s
        t[3] = (m[k][j] - m[j][k]) * s
        t[j] = (m[j][i] + m[i][j]) * s
        t[k] = (m[k][i] + m[i][k]) * s
    return t


def q_slerp(a, b, t):
    """tf2::Quaternion::slerp (shortest path, acos of |dot|/sqrt(len2*len2))."""
    dot = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3]
    s = math.sqrt((a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3]) * (b[0] * b[0] + b[1] * b[1] + b[2] * b[2] + b[3] * b[3]))
    c = (-dot if dot < 0 else dot) / s
    c = -1.0 if c < -1.0 else (1.0 if c > 1.0 else c)  # tf2Acos clamps
    theta = math.acos(c)
    if theta != 0.0:
        d = 1.0 / math.sin(theta)
        s0 = math.sin((1.0 - t) * theta)
        s1 = math.sin(t * theta)
        if dot < 0:
            return [(a[i] * s0 + -b[i] * s1) * d for i in range(4)]
        return [(a[i] * s0 + b[i] * s1) * d for i in range(4)]
    return list(a)


def q_normalized(q):
    n = math.sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3])
    return [q[0] / n, q[1] / n, q[2] / n, q[3] / n]


def get_rpy(m):
    """tf2::Matrix3x3::getRPY (solution 1)."""
    if abs(m[2][0]) >= 1.0:
        yaw = 0.0
        delta = math.atan2(m[2][1], m[2][2])
        if m[2][0] < 0:
            pitch = math.pi / 2.0
            roll = delta
        else:
            pitch = -math.pi / 2.0
            roll = delta
        return roll, pitch, yaw
    pitch = -math.asin(m[2][0])
    c = math.cos(pitch)
    roll = math.atan2(m[2][1] / c, m[2][2] / c)
    yaw = math.atan2(m[1][0] / c, m[0][0] / c)
    return roll, pitch, yaw


def q_from_rpy(roll, pitch, yaw):
    """tf2::Quaternion::setRPY."""
    hy, hp, hr = yaw * 0.5, pitch * 0.5, roll * 0.5
    cy, sy = math.cos(hy), math.sin(hy)
    cp, sp = math.cos(hp), math.sin(hp)
    cr, sr = math.cos(hr), math.sin(hr)
    return [sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy]


def _mm(A, B):
    return [[A[i][0] * B[0][j] + A[i][1] * B[1][j] + A[i][2] * B[2][j] for j in range(3)] for i in range(3)]


def _mv(A, v):
    return [A[i][0] * v[0] + A[i][1] * v[1] + A[i][2] * v[2] for i in range(3)]


# ----------------------------- TransformWithVariance -----------------------------------
@dataclass
class TWV:
    """TransformWithVariance (transform_with_variance.h:10-63): tf2::Transform + scalar variance."""

    R: list
    t: list
    var: float = 0.0

    @staticmethod
    def from_qt(q_xyzw, t, var=0.0):
        return TWV(q_to_m(list(q_xyzw)), [float(t[0]), float(t[1]), float(t[2])], float(var))

    @staticmethod
    def identity(var=0.0):
        return TWV([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]], [0.0, 0.0, 0.0], var)

    def copy(self):
        return TWV([row[:] for row in self.R], self.t[:], self.var)

    def mul(self, o: "TWV", add_var=True):
        """operator*= (h:26-34): tf2 compose (R1 R2, R1 t2 + t1); variances add."""
        Rt = _mv(self.R, o.t)
        return TWV(_mm(self.R, o.R), [Rt[0] + self.t[0], Rt[1] + self.t[1], Rt[2] + self.t[2]], self.var + (o.var if add_var else 0.0))

    def inverse(self):
        """tf2::Transform::inverse: (R^T, R^T * -t)."""
        Rt = [[self.R[j][i] for j in range(3)] for i in range(3)]
        return TWV(Rt, _mv(Rt, [-self.t[0], -self.t[1], -self.t[2]]), self.var)

    def update(self, n: "TWV"):
        """TransformWithVariance::update, transform_with_variance.cpp:43-78."""
        p1, q1, v1 = self.t, m_to_q(self.R), self.var
        p2, q2, v2 = n.t, m_to_q(n.R), n.var
        k = v1 / (v1 + v2)  # kalman_gain :9
        self.t = [p1[i] + k * (p2[i] - p1[i]) for i in range(3)]
        self.R = q_to_m(q_normalized(q_slerp(q1, q2, k)))
        d = [p2[i] - p1[i] for i in range(3)]
        mean2 = math.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
        e = [self.t[i] - p1[i] for i in range(3)]
        mean = math.sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2])
        self.var = normalize_david(mean, 0.0, v1, mean2, v2)


def probability_at_point(x, u, var):
    """probabiltyAtPoint, transform_with_variance.cpp:14-17 -- IEEE semantics (a zero variance gives inf / NaN, not an exception)."""
    with np.errstate(all="ignore"):
        x, u, var = np.float64(x), np.float64(u), np.float64(var)
        return float((np.float64(1.0) / (np.sqrt(var) * np.sqrt(np.float64(2.0 * math.pi)))) * np.exp(-((x - u) * (x - u)) / (np.float64(2.0) * var)))


def normalize_david(new_mean, mean1, var1, mean2, var2):
    """normalizeDavid, transform_with_variance.cpp:23-38."""
    p1 = probability_at_point(new_mean, mean1, var1)
    p2 = probability_at_point(new_mean, mean2, var2)
    with np.errstate(all="ignore"):  # std::pow / division overflow to +inf, NaN propagates (found by tests/test_map_ref.py)
        p = np.sqrt(np.float64(p1) * np.float64(p1) + np.float64(p2) * np.float64(p2))
        x = np.float64(1.0) / (p * np.sqrt(np.float64(2.0 * math.pi)))
        nv = float(x * x)
    nv = 1e3 if 1e3 < nv else nv    # std::min(newVar, 1e3): NaN stays NaN
    nv = 1e-8 if nv < 1e-8 else nv  # std::max(newVar, 1e-8)
    return nv


def average_transforms(t1: TWV, t2: TWV) -> TWV:
    """averageTransforms, transform_with_variance.cpp:81-85."""
    out = t1.copy()
    out.update(t2)
    return out


# ----------------------------- Map ------------------------------------------------------
@dataclass
class Observation:
    """Observation, map.cpp:53-59."""

    fid: int
    T_camFid: TWV
    T_fidCam: TWV = None

    def __post_init__(self):
        if self.T_fidCam is None:
            self.T_fidCam = self.T_camFid.inverse()


@dataclass
class Fiducial:
    """Fiducial, map.h:74-89."""

    id: int
    pose: TWV
    numObs: int = 0
    links: set = field(default_factory=set)


def observations_from_transforms(transforms, weighting_scale=1e9, use_area=False):
    """transformCallback, fiducial_slam.cpp:79-105.  ``transforms`` = iterable of dicts with
    fiducial_id, translation, rotation (xyzw), object_error, fiducial_area."""
    obs = []
    for ft in transforms:
        var = weighting_scale / ft["fiducial_area"] if use_area else weighting_scale * ft["object_error"]
        obs.append(Observation(int(ft["fiducial_id"]), TWV.from_qt(ft["rotation"], ft["translation"], var)))
    return obs


class Map:
    """Map (map.h:92-160) -- state + update arithmetic only (no ROS I/O)."""

    def __init__(self, read_only=False):
        self.fiducials: dict[int, Fiducial] = {}
        self.frameNum = 0
        self.initialFrameNum = 0
        self.originFid = -1
        self.isInitializingMap = False
        self.readOnly = read_only
        self.fiducialToAdd = -1          # add_fiducial service (map.cpp:821-828)
        self.addMapBase = None           # tf lookup map -> base at the time of handleAddFiducial (map.cpp:514-517), None = failed

    def load_entry(self, fid, x, y, z, roll_deg, pitch_deg, yaw_deg, variance, num_obs=0, links=()):
        """One line of loadMap, map.cpp:595-606 (degrees -> setRPY)."""
        q = q_from_rpy(math.radians(roll_deg), math.radians(pitch_deg), math.radians(yaw_deg))
        f = Fiducial(int(fid), TWV.from_qt(q, [x, y, z], variance), int(num_obs), set(links))
        self.fiducials[int(fid)] = f

    # map.cpp:152-176
    def update(self, obs, T_baseCam: TWV | None, T_camBase: TWV | None):
        """Returns the robot pose T_mapBase (TWV) or None.  ``T_baseCam`` None == tf lookup failed."""
        self.frameNum += 1
        robot = None
        if len(obs) > 0 and len(self.fiducials) == 0:
            self.isInitializingMap = True
        if self.isInitializingMap:
            self.auto_init(obs, T_baseCam)
        else:
            n, T_mapCam, robot = self.update_pose(obs, T_baseCam, T_camBase)
            if n > 0 and len(obs) > 1 and not self.readOnly:
                self.update_map(obs, T_mapCam)
        self.handle_add_fiducial(obs, T_baseCam)  # :173
        return robot

    # map.cpp:489-535
    def handle_add_fiducial(self, obs, T_baseCam: TWV | None):
        if self.fiducialToAdd == -1:
            return
        if self.fiducialToAdd in self.fiducials:
            self.fiducialToAdd = -1
            return
        for o in obs:
            if o.fid == self.fiducialToAdd:
                T = o.T_camFid.copy()
                if T_baseCam is not None:  # T.setData(T_baseCam * T): tf2::Transform * TransformWithVariance keeps T's variance
                    var = T.var
                    T = T_baseCam.mul(T, add_var=False)
                    T.var = var
                if self.addMapBase is not None:
                    var = T.var
                    T = self.addMapBase.mul(T, add_var=False)
                    T.var = var
                self.fiducials[o.fid] = Fiducial(o.fid, T)
                # `fiducials[originFid].pose.variance = 0.0`: with originFid == -1 (map from a file) the reference's operator[]
                # inserts a default-constructed Fiducial with uninitialised members under key -1; only the defined case is restated
                if self.originFid in self.fiducials:
                    self.fiducials[self.originFid].pose.var = 0.0
                self.isInitializingMap = False
                self.fiducialToAdd = -1
                return

    # map.cpp:181-225
    def update_map(self, obs, T_mapCam: TWV):
        for o in obs:
            T_mapFid = T_mapCam.mul(o.T_camFid)
            if any(math.isnan(v) for v in T_mapFid.t):
                continue
            if o.fid not in self.fiducials:
                self.fiducials[o.fid] = Fiducial(o.fid, T_mapFid.copy())
            f = self.fiducials[o.fid]
            if f.pose.var != 0:
                f.pose.update(T_mapFid)  # Fiducial::update :62-65 -> numObs++
                f.numObs += 2  # ... and numObs++ again at :215
            for other in obs:
                if other.fid != f.id:
                    f.links.add(other.fid)

    # map.cpp:247-391
    def update_pose(self, obs, T_baseCam: TWV | None, T_camBase: TWV | None):
        if len(obs) == 0:
            return 0, None, None
        cam_base = T_camBase.copy() if T_camBase is not None else TWV.identity()
        cam_base.var = 1.0 if T_camBase is not None else 0.0  # :261 (default-constructed otherwise)
        if T_baseCam is None:
            return 0, None, None  # :272
        base_cam = T_baseCam.copy()
        base_cam.var = 1.0  # :269
        n = 0
        T_mapBase = None
        for o in obs:
            if o.fid in self.fiducials:
                fid = self.fiducials[o.fid]
                p = fid.pose.mul(o.T_fidCam)  # :279
                p = p.mul(cam_base)  # :284
                pos = p.t
                roll, pitch, _yaw = get_rpy(p.R)  # :287
                c = o.T_camFid.t
                s1 = math.pow(pos[2] / c[2], 2) * (math.pow(c[0], 2) + math.pow(c[1], 2))  # :293-294
                len2 = pos[0] * pos[0] + pos[1] * pos[1] + pos[2] * pos[2]
                s2 = len2 * math.pow(math.sin(roll), 2)
                s3 = len2 * math.pow(math.sin(pitch), 2)
                p.var = s1 + s2 + s3 + SYSTEMATIC_ERROR  # :297
                o.T_camFid.var = p.var  # :298 write-back
                if math.isnan(pos[0]) or math.isnan(pos[1]) or math.isnan(pos[2]):
                    continue
                if n == 0:
                    T_mapBase = p
                else:
                    T_mapBase = average_transforms(T_mapBase, p)  # :315
                n += 1
        if n == 0:
            return 0, None, None
        T_mapCam = T_mapBase.mul(base_cam)  # :347
        return n, T_mapCam, T_mapBase

    # map.cpp:415-485
    def auto_init(self, obs, T_baseCam: TWV | None):
        if len(self.fiducials) == 0:
            idx = -1
            smallest = -1.0
            for i, o in enumerate(obs):
                d = o.T_camFid.t[0] ** 2 + o.T_camFid.t[1] ** 2 + o.T_camFid.t[2] ** 2
                if smallest < 0 or d < smallest:
                    smallest = d
                    idx = i
            if idx == -1:
                return
            o = obs[idx]
            self.originFid = o.fid
            T = o.T_camFid.copy()
            if T_baseCam is not None:
                T = T_baseCam.mul(T, add_var=False)
                T.var = o.T_camFid.var  # operator*(tf2::Transform, TWV) keeps rhs variance (h:55-59)
            self.fiducials[o.fid] = Fiducial(o.fid, T)
        else:
            for o in obs:
                if o.fid == self.originFid:
                    T = o.T_camFid.copy()
                    if T_baseCam is not None:
                        T = T_baseCam.mul(T, add_var=False)
                        T.var = o.T_camFid.var
                    f = self.fiducials[self.originFid]
                    f.pose.update(T)
                    f.numObs += 1  # Fiducial::update :62-65
                    break
        if self.frameNum - self.initialFrameNum > 10 and self.originFid != -1:
            self.isInitializingMap = False
            self.fiducials[self.originFid].pose.var = 0.0

    # map.cpp:629-654
    def entries(self):
        """[(fiducial_id, x, y, z, rx, ry, rz)] ascending id (std::map order)."""
        out = []
        for fid in sorted(self.fiducials):
            f = self.fiducials[fid]
            r, p, y = get_rpy(f.pose.R)
            out.append((fid, f.pose.t[0], f.pose.t[1], f.pose.t[2], r, p, y))
        return out


def save_map_text(m: "Map") -> str:
    """Map::saveMap, map.cpp:541-566: one line per fiducial in ascending id,
    `id x y z rx ry rz(deg) variance numObs link link ...` with C's %lf (six decimals)."""
    lines = []
    for fid in sorted(m.fiducials):
        f = m.fiducials[fid]
        r, p, y = get_rpy(f.pose.R)
        rad2deg = lambda a: a * 180.0 / math.pi  # helpers.h:8
        head = "%d %f %f %f %f %f %f %f %d" % (f.id, f.pose.t[0], f.pose.t[1], f.pose.t[2], rad2deg(r), rad2deg(p), rad2deg(y), f.pose.var, f.numObs)
        lines.append(head + "".join(" %d" % k for k in sorted(f.links)))
    return "".join(line + "\n" for line in lines)


def load_map_text(m: "Map", text: str) -> int:
    """Map::loadMap, map.cpp:572-625: sscanf("%d %lf %lf %lf %lf %lf %lf %lf %d%[^\t\n]*s"); a line is accepted when
    the nine leading fields parse (nElems 9 or 10), the rest of the line is a blank-separated list of linked ids.
    Returns the number of entries read."""
    n = 0
    for line in text.splitlines():
        tok = line.split("\t")[0].split()
        try:
            fid = int(tok[0])
            vals = [float(v) for v in tok[1:8]]
            num_obs = int(tok[8])
            if len(vals) != 7:
                raise ValueError
        except (ValueError, IndexError):
            continue  # ROS_WARN("Invalid line")
        links = [int(v) for v in tok[9:]]
        m.load_entry(fid, *vals, num_obs=num_obs, links=links)
        n += 1
    return n


def merge_maps(tables):
    """Deterministic merge of per-rank map tables (NEW -- no reference counterpart; parity
    unpinned, checked only against this restatement).  ``tables`` = list (rank order) of lists of
    (id, TWV, numObs); result: dict id -> (TWV, numObs).  Ranks are folded in rank order, ids
    ascending, fused with TransformWithVariance::update; variance-0 (pinned) entries win."""
    merged: dict[int, tuple] = {}
    for table in tables:
        for fid, pose, num_obs in sorted(table, key=lambda e: e[0]):
            if fid not in merged:
                merged[fid] = (pose.copy(), int(num_obs))
                continue
            cur, n = merged[fid]
            if cur.var == 0.0:
                merged[fid] = (cur, n + int(num_obs))
            elif pose.var == 0.0:
                merged[fid] = (pose.copy(), n + int(num_obs))
            else:
                cur.update(pose)
                merged[fid] = (cur, n + int(num_obs))
    return merged


# ----------------------------- published pose (map.cpp:337-379) --------------------------
def pose_covariance(variance, covariance_diagonal=None):
    """toPose (transform_with_variance.h:69-84) + the covariance_diagonal override (map.cpp:110-125,341-345): the 36-entry
    row-major covariance of the PoseWithCovarianceStamped on /fiducial_pose.  The override is ignored unless it has six non-zero values."""
    cov = [0.0] * 36
    diag = [variance] * 6
    if covariance_diagonal is not None and len(covariance_diagonal) == 6 and all(v != 0 for v in covariance_diagonal):
        diag = [float(v) for v in covariance_diagonal]
    for i in range(6):
        cov[i * 6 + i] = diag[i]
    return cov


def set_rpy_matrix(roll, pitch, yaw):
    """tf2::Matrix3x3::setRPY == setEulerYPR(yaw, pitch, roll)."""
    ci, cj, ch = math.cos(roll), math.cos(pitch), math.cos(yaw)
    si, sj, sh = math.sin(roll), math.sin(pitch), math.sin(yaw)
    cc, cs, sc, ss = ci * ch, ci * sh, si * ch, si * sh
    return [[cj * ch, sj * sc - cs, sj * cc + ss], [cj * sh, sj * ss + cc, sj * cs - sc], [-sj, cj * si, cj * ci]]


def published_pose_tf(T_mapBase: TWV, T_odomBase: TWV | None = None, publish_6dof_pose=False):
    """The map -> odom (or map -> base) transform of updatePose's tail, map.cpp:351-379: outPose = basePose * odom^-1 when the
    odom lookup succeeded, then squashed to x, y, yaw unless publish_6dof_pose."""
    out = T_mapBase.copy()
    if T_odomBase is not None:
        var = out.var
        out = out.mul(T_odomBase.inverse(), add_var=False)
        out.var = var
    if not publish_6dof_pose:
        out.t[2] = 0.0
        _r, _p, yaw = get_rpy(out.R)
        out.R = set_rpy_matrix(0.0, 0.0, yaw)
    return out
