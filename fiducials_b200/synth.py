"""Deterministic synthetic frames for the parity tests and bench.py (numpy only).

Follows SURVEY.md section 8(d): markers of a predefined ArUco dictionary with a white quiet
zone, laid on a grid, each warped by a random in-plane rotation plus per-corner jitter, on a
light background with a slow sinusoidal gradient, Gaussian blur (sigma 1) and +-4 integer
noise, gray replicated to BGR8.  Camera intrinsics mimic the reference test's CameraInfo
(aruco_detect/test/aruco_images_test.cpp:20-29): fx = fy = 0.73 W, principal point at the
centre, plumb_bob distortion D from that test.
"""
from __future__ import annotations

import math
import os

import numpy as np

_DICT_NPZ = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dictionaries.npz")
_dict_cache: dict = {}

# aruco_detect/test/aruco_images_test.cpp:24-27
REFERENCE_TEST_D = (0.1349735087283542, -0.2335869827451621, 0.0006697030315075139, 0.004846737465872353, 0.0)

# BASELINE.json configs: name -> (W, H, n_markers, dictionary enum)
CONFIGS = {
    "C1": (640, 480, 4, 6),
    "C2": (1920, 1080, 16, 10),
    "C3": (1280, 720, 8, 6),
    "C4": (3840, 2160, 64, 10),
}


def dictionary_info(dict_id: int):
    """(marker_size, n_markers, max_correction_bits, bytes uint8[n, 4*nbytes])."""
    if not _dict_cache:
        z = np.load(_DICT_NPZ)
        _dict_cache["info"] = {int(r[0]): (int(r[1]), int(r[2]), int(r[3])) for r in z["info"]}
        names = {0: "4x4", 1: "4x4", 2: "4x4", 3: "4x4", 4: "5x5", 5: "5x5", 6: "5x5", 7: "5x5", 8: "6x6", 9: "6x6", 10: "6x6", 11: "6x6", 12: "7x7", 13: "7x7",
                 14: "7x7", 15: "7x7", 16: "original", 17: "apriltag16h5", 18: "apriltag25h9", 19: "apriltag36h10", 20: "apriltag36h11", 21: "mip36h12"}
        for e, nm in names.items():
            _dict_cache[("bytes", e)] = z["bytes_" + nm]
    if dict_id not in _dict_cache["info"]:
        raise ValueError("unsupported dictionary enum %d (supported: 0..21)" % dict_id)
    ms, n, mc = _dict_cache["info"][dict_id]
    return ms, n, mc, _dict_cache[("bytes", dict_id)][:n]


def marker_bits(dict_id: int, marker_id: int) -> np.ndarray:
    """(ms+2) x (ms+2) uint8 cell image incl. the black border: 1 = white cell."""
    ms, n, _, tab = dictionary_info(dict_id)
    nb = tab.shape[1] // 4
    rot0 = tab[marker_id, :nb]
    bits = np.zeros(ms * ms, np.uint8)
    full = (ms * ms) // 8
    k = 0
    for byte_i in range(nb):
        nbits = 8 if byte_i < full else ms * ms - 8 * full
        for j in range(nbits):
            bits[k] = (int(rot0[byte_i]) >> (nbits - 1 - j)) & 1
            k += 1
    out = np.zeros((ms + 2, ms + 2), np.uint8)
    out[1:-1, 1:-1] = bits.reshape(ms, ms)
    return out


def camera_for(W: int, H: int):
    f = 0.73 * W
    K = np.array([[f, 0.0, W / 2.0], [0.0, f, H / 2.0], [0.0, 0.0, 1.0]], np.float64)
    return K, np.array(REFERENCE_TEST_D, np.float64)


def _homography(src, dst):
    A = []
    b = []
    for (x, y), (u, v) in zip(src, dst):
        A.append([x, y, 1, 0, 0, 0, -u * x, -u * y])
        A.append([0, 0, 0, x, y, 1, -v * x, -v * y])
        b += [u, v]
    h = np.linalg.solve(np.array(A, np.float64), np.array(b, np.float64))
    return np.append(h, 1.0).reshape(3, 3)


def _blur(img: np.ndarray, sigma: float) -> np.ndarray:
    r = int(math.ceil(3 * sigma))
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2).astype(np.float32)
    k /= k.sum()
    H, W = img.shape
    p = np.pad(img, ((0, 0), (r, r)), mode="edge")
    t = np.zeros_like(img)
    for i, w in enumerate(k):
        t += w * p[:, i : i + W]
    p = np.pad(t, ((r, r), (0, 0)), mode="edge")
    out = np.zeros_like(img)
    for i, w in enumerate(k):
        out += w * p[i : i + H, :]
    return out


def _rodrigues(rvec):
    th = float(np.linalg.norm(rvec))
    if th < 1e-12:
        return np.eye(3)
    k = np.asarray(rvec, np.float64) / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + math.sin(th) * Kx + (1 - math.cos(th)) * (Kx @ Kx)


def project_points(obj, R, t, K, D):
    """Pinhole + plumb_bob projection (numpy), obj [n,3] -> [n,2]."""
    P = obj @ R.T + t
    x, y = P[:, 0] / P[:, 2], P[:, 1] / P[:, 2]
    k1, k2, p1, p2, k3 = [float(v) for v in D[:5]]
    r2 = x * x + y * y
    cd = 1 + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2
    xd = x * cd + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    yd = y * cd + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    return np.stack([xd * K[0, 0] + K[0, 2], yd * K[1, 1] + K[1, 2]], 1)


def make_frame(W: int, H: int, n_markers: int, dict_id: int, seed: int = 0, jitter: float = 0.08, first_id: int = 0, supersample: int = 2, marker_len: float = 0.14,
               max_tilt_deg: float = 35.0, mode: str = "pose"):
    """Returns (bgr uint8[H,W,3], truth list of (id, float64[4,2] TL,TR,BR,BL image corners)).

    mode "pose": each marker is a square of side ``marker_len`` at a sampled 3-D pose (in-plane angle
    U[0,2pi), tilt <= max_tilt_deg about a random in-plane axis, depth chosen so that it fills about
    0.55 of its grid cell) projected with camera_for(W,H) incl. distortion.  mode "jitter": in-plane
    rotation plus per-corner jitter (not a consistent perspective view)."""
    rng = np.random.default_rng(seed)
    ms, nmk, _, _ = dictionary_info(dict_id)
    cells = ms + 2
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    phase = rng.uniform(0, 2 * math.pi, 2)
    img = 190.0 + 20.0 * np.sin(2 * math.pi * xx / (1.7 * W) + phase[0]) * np.cos(2 * math.pi * yy / (1.3 * H) + phase[1])
    img = img.astype(np.float32)
    cols = int(math.ceil(math.sqrt(n_markers * W / H)))
    rows = int(math.ceil(n_markers / cols))
    cw, ch = W / cols, H / rows
    truth = []
    ss = supersample
    offs = (np.arange(ss) + 0.5) / ss - 0.5
    for i in range(n_markers):
        mid = (first_id + i) % nmk
        bits = marker_bits(dict_id, mid)
        side = 0.55 * min(cw, ch)
        cx = (i % cols + 0.5) * cw + rng.uniform(-0.05, 0.05) * cw
        cy = (i // cols + 0.5) * ch + rng.uniform(-0.05, 0.05) * ch
        ang = rng.uniform(0, 2 * math.pi)
        ca, sa = math.cos(ang), math.sin(ang)
        base = np.array([[-0.5, -0.5], [0.5, -0.5], [0.5, 0.5], [-0.5, 0.5]]) * side
        quad = np.stack([cx + ca * base[:, 0] - sa * base[:, 1], cy + sa * base[:, 0] + ca * base[:, 1]], 1)
        jit = rng.uniform(-jitter, jitter, (4, 2)) * side
        if mode == "pose":
            Kc, Dc = camera_for(W, H)
            f = Kc[0, 0]
            z = f * marker_len / side
            axis_ang = rng.uniform(0, 2 * math.pi)
            tilt = math.radians(rng.uniform(0, max_tilt_deg))
            R = _rodrigues(np.array([math.cos(axis_ang), math.sin(axis_ang), 0.0]) * tilt) @ _rodrigues(np.array([0.0, 0.0, ang]))
            # marker frame: x right, y up, z out of the marker; camera looks along +z with y down
            R = R @ np.diag([1.0, -1.0, -1.0])
            t = np.array([(cx - Kc[0, 2]) * z / f, (cy - Kc[1, 2]) * z / f, z])
            hl = marker_len / 2.0
            obj = np.array([[-hl, hl, 0], [hl, hl, 0], [hl, -hl, 0], [-hl, -hl, 0]])
            quad = project_points(obj, R, t, Kc, Dc)
        else:
            quad += jit
        truth.append((mid, quad.copy()))
        # marker plane coordinates: marker occupies [0,1]^2, quiet zone extends by q on every side
        q = 1.0 / 6.0
        Hm = _homography([(0, 0), (1, 0), (1, 1), (0, 1)], quad)
        Hinv = np.linalg.inv(Hm)
        outer = (Hm @ np.array([[-q, -q, 1], [1 + q, -q, 1], [1 + q, 1 + q, 1], [-q, 1 + q, 1]]).T).T
        outer = outer[:, :2] / outer[:, 2:3]
        x0 = max(int(math.floor(outer[:, 0].min())) - 1, 0)
        x1 = min(int(math.ceil(outer[:, 0].max())) + 2, W)
        y0 = max(int(math.floor(outer[:, 1].min())) - 1, 0)
        y1 = min(int(math.ceil(outer[:, 1].max())) + 2, H)
        if x1 <= x0 or y1 <= y0:
            continue
        py, px = np.mgrid[y0:y1, x0:x1].astype(np.float64)
        acc = np.zeros(py.shape, np.float64)
        cov = np.zeros(py.shape, np.float64)
        for oy in offs:
            for ox in offs:
                X = px + ox
                Y = py + oy
                w = Hinv[2, 0] * X + Hinv[2, 1] * Y + Hinv[2, 2]
                u = (Hinv[0, 0] * X + Hinv[0, 1] * Y + Hinv[0, 2]) / w
                v = (Hinv[1, 0] * X + Hinv[1, 1] * Y + Hinv[1, 2]) / w
                inside_q = (u >= -q) & (u < 1 + q) & (v >= -q) & (v < 1 + q)
                inside_m = (u >= 0) & (u < 1) & (v >= 0) & (v < 1)
                ci = np.clip((u * cells).astype(np.int64), 0, cells - 1)
                cj = np.clip((v * cells).astype(np.int64), 0, cells - 1)
                val = np.where(inside_m, np.where(bits[cj, ci] > 0, 235.0, 20.0), 235.0)
                acc += np.where(inside_q, val, 0.0)
                cov += inside_q
        n_s = float(ss * ss)
        a = cov / n_s
        region = img[y0:y1, x0:x1]
        img[y0:y1, x0:x1] = (region * (1.0 - a) + acc / n_s).astype(np.float32)
    img = _blur(img, 1.0)
    img += rng.integers(-4, 5, img.shape).astype(np.float32)
    g = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    return np.repeat(g[:, :, None], 3, axis=2), truth


def make_config_stream(name: str, n_frames: int, seed: int = 0, realizations: int = 4):
    """n distinct frames of a BASELINE config: ceil(n/realizations) marker layouts, each with
    `realizations` independent noise draws (cheap to generate, distinct threshold planes)."""
    W, H, n, d = CONFIGS[name]
    out = np.empty((n_frames, H, W, 3), np.uint8)
    truths = []
    rng = np.random.default_rng(seed + 7919)
    base = None
    for i in range(n_frames):
        if i % realizations == 0:
            base, truth = make_frame(W, H, n, d, seed * 100003 + i, first_id=(i * 7) % 200)
        g = base[:, :, 0].astype(np.int16)
        if i % realizations:
            g = g + rng.integers(-3, 4, g.shape, dtype=np.int16)
        out[i] = np.clip(g, 0, 255).astype(np.uint8)[:, :, None]
        truths.append(truth)
    K, D = camera_for(W, H)
    return out, truths, K, D, d


def _q_from_rpy(roll, pitch, yaw):
    hy, hp, hr = yaw * 0.5, pitch * 0.5, roll * 0.5
    cy, sy, cp, sp, cr, sr = math.cos(hy), math.sin(hy), math.cos(hp), math.sin(hp), math.cos(hr), math.sin(hr)
    return np.array([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy])


def _q_to_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def _R_to_q(R):
    w = math.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    x = math.copysign(math.sqrt(max(0.0, 1 + R[0, 0] - R[1, 1] - R[2, 2])) / 2, R[2, 1] - R[1, 2])
    y = math.copysign(math.sqrt(max(0.0, 1 - R[0, 0] + R[1, 1] - R[2, 2])) / 2, R[0, 2] - R[2, 0])
    z = math.copysign(math.sqrt(max(0.0, 1 - R[0, 0] - R[1, 1] + R[2, 2])) / 2, R[1, 0] - R[0, 1])
    return np.array([x, y, z, w])


def make_c5_sequence(n_frames: int = 1000, seed: int = 0, cols: int = 25, rows: int = 20, visible: int = 10):
    """BASELINE.json config C5: 500 fiducials on a 25x20 ceiling grid (1 m pitch,
    z = 2.5 m, rpy = (180,0,180) deg like fiducial_slam/scripts/init_map.py:30), a camera on a
    lawn-mower path below, `visible` nearest fiducials per frame with noise sigma_t = 5 mm,
    sigma_R = 0.5 deg and object_error ~ U[1e-4, 1e-2].  Returns (messages, seed_entry) where
    messages[k] is a list of FiducialTransform-like dicts and seed_entry the map-file row that pins
    fiducial 0 (variance 0).  The camera is the base frame (identity T_baseCam)."""
    rng = np.random.default_rng(seed)
    pos = np.array([[float(i % cols), float(i // cols), 2.5] for i in range(cols * rows)])
    q_fid = _q_from_rpy(math.pi, 0.0, math.pi)
    R_fid = _q_to_R(q_fid)
    msgs = []
    for k in range(n_frames):
        s = k / max(1, n_frames - 1) * rows  # lawn mower: sweep x back and forth while y advances
        row = min(int(s), rows - 1)
        fx = s - row
        cam = np.array([fx * (cols - 1) if row % 2 == 0 else (1 - fx) * (cols - 1), row + 0.0, 0.0])
        yaw = 0.3 * math.sin(0.05 * k)
        R_cam = _q_to_R(_q_from_rpy(0.0, 0.0, yaw))
        d = np.linalg.norm(pos[:, :2] - cam[:2], axis=1)
        vis = np.argsort(d, kind="stable")[:visible]
        order = rng.permutation(vis)
        m = []
        for i in order:
            t = R_cam.T @ (pos[i] - cam) + rng.normal(0, 0.005, 3)
            dR = _q_to_R(_q_from_rpy(*rng.normal(0, math.radians(0.5), 3)))
            q = _R_to_q(R_cam.T @ R_fid @ dR)
            m.append(dict(fiducial_id=int(i), translation=t, rotation=q, image_error=0.1, object_error=float(rng.uniform(1e-4, 1e-2)), fiducial_area=1500.0))
        msgs.append(m)
    seed_entry = [0, pos[0][0], pos[0][1], pos[0][2], 180.0, 0.0, 180.0, 0.0, 0]
    return msgs, seed_entry


def make_config_frame(name: str, seed: int = 0):
    W, H, n, d = CONFIGS[name]
    bgr, truth = make_frame(W, H, n, d, seed)
    K, D = camera_for(W, H)
    return bgr, truth, K, D, d
