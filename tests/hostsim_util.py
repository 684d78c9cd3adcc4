"""ctypes loader for tests/hostsim/libfid_hostsim.so (CPU harness for the FID_HD device functions)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def load():
    global _lib
    if _lib is None:
        so = os.path.join(_HERE, "hostsim", "libfid_hostsim.so")
        src = os.path.join(_HERE, "hostsim", "hostsim.cpp")
        csrc = os.path.join(os.path.dirname(_HERE), "fiducials_b200", "csrc")
        newest = max([os.path.getmtime(src)] + [os.path.getmtime(os.path.join(csrc, f)) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))])
        if not os.path.exists(so) or os.path.getmtime(so) < newest:
            subprocess.check_call([os.path.join(_HERE, "hostsim", "build.sh")])
        _lib = C.CDLL(so)
    return _lib


def find_contours(plane, min_len=1, max_len=1 << 30, mode=1):
    lib = load()
    plane = np.ascontiguousarray(plane, np.uint8)
    H, W = plane.shape
    max_pts = 4 * plane.size + 16
    pts = np.zeros((max_pts, 2), np.int16)
    lens = np.zeros(plane.size + 16, np.int32)
    nstarts = C.c_int64(0)
    n = lib.hs_find_contours_mode(plane.ctypes.data_as(C.c_void_p), W, H, min_len, min(max_len, 1 << 30), pts.ctypes.data_as(C.c_void_p), C.c_int64(max_pts),
                                  lens.ctypes.data_as(C.c_void_p), len(lens), C.byref(nstarts), mode)
    assert n >= 0
    out = []
    off = 0
    for i in range(n):
        out.append(pts[off:off + lens[i]].astype(np.int32))
        off += lens[i]
    return out, nstarts.value


def approx_poly(pts, eps):
    lib = load()
    p = np.ascontiguousarray(pts, np.int16)
    out = np.zeros((16, 2), np.int16)
    lib.hs_approx_poly.restype = C.c_int
    n = lib.hs_approx_poly(p.ctypes.data_as(C.c_void_p), len(p), C.c_double(eps), out.ctypes.data_as(C.c_void_p))
    return n, out[:max(n, 0)].astype(np.int32)


def is_convex(pts):
    lib = load()
    p = np.ascontiguousarray(pts, np.int16)
    return bool(lib.hs_is_convex(p.ctypes.data_as(C.c_void_p), len(p)))


def candidates(planes, dict_id):
    lib = load()
    planes = np.ascontiguousarray(planes, np.uint8)
    S, H, W = planes.shape
    cap = 8192
    quads = np.zeros((cap, 8), np.int32)
    scale = np.zeros(cap, np.int32)
    clen = np.zeros(cap, np.int32)
    n = lib.hs_candidates(planes.ctypes.data_as(C.c_void_p), W, H, dict_id, quads.ctypes.data_as(C.c_void_p), scale.ctypes.data_as(C.c_void_p),
                          clen.ctypes.data_as(C.c_void_p), cap)
    assert n >= 0, n
    return quads[:n].reshape(n, 4, 2), scale[:n], clen[:n]


def detect(gray, planes, dict_id, refine=True):
    lib = load()
    gray = np.ascontiguousarray(gray, np.uint8)
    planes = np.ascontiguousarray(planes, np.uint8)
    S, H, W = planes.shape
    ids = np.zeros(256, np.int32)
    corners = np.zeros((256, 8), np.float32)
    stats = np.zeros(4, np.int32)
    n = lib.hs_detect(gray.ctypes.data_as(C.c_void_p), planes.ctypes.data_as(C.c_void_p), W, H, dict_id, int(refine), ids.ctypes.data_as(C.c_void_p),
                      corners.ctypes.data_as(C.c_void_p), 256, stats.ctypes.data_as(C.c_void_p))
    assert n >= 0, n
    return ids[:n].copy(), corners[:n].reshape(n, 4, 2).copy(), stats


def corner_subpix(gray, pts, win=5, max_iters=30, eps=0.01):
    lib = load()
    gray = np.ascontiguousarray(gray, np.uint8)
    p = np.ascontiguousarray(pts, np.float32).copy()
    H, W = gray.shape
    lib.hs_corner_subpix(gray.ctypes.data_as(C.c_void_p), W, H, p.ctypes.data_as(C.c_void_p), len(p), win, max_iters, C.c_double(eps))
    return p


def pose(corners, K, D, lens, default_len):
    lib = load()
    c = np.ascontiguousarray(corners, np.float32).reshape(-1, 8)
    n = len(c)
    K = np.ascontiguousarray(K, np.float64).reshape(9)
    D = np.ascontiguousarray(D, np.float64).reshape(-1)[:5]
    lens = np.ascontiguousarray(lens, np.float32)
    out = np.zeros((n, 16))
    lib.hs_pose(n, c.ctypes.data_as(C.c_void_p), K.ctypes.data_as(C.c_void_p), D.ctypes.data_as(C.c_void_p), lens.ctypes.data_as(C.c_void_p),
                C.c_double(default_len), out.ctypes.data_as(C.c_void_p))
    return out


class HsMap:
    def __init__(self, capacity=64, read_only=False):
        self.lib = load()
        self.lib.hs_map_create.restype = C.c_void_p
        self.h = C.c_void_p(self.lib.hs_map_create(capacity, int(read_only)))
        self.cap = capacity

    def load(self, fid, x, y, z, r, p, yw, var, num_obs=0):
        self.lib.hs_map_load(self.h, int(fid), *[C.c_double(v) for v in (x, y, z, r, p, yw, var)], int(num_obs))

    def update(self, transforms, base_cam, cam_base):
        obs = np.zeros((len(transforms), 10))
        for i, t in enumerate(transforms):
            obs[i, 0] = t["fiducial_id"]
            obs[i, 1:4] = t["translation"]
            obs[i, 4:8] = t["rotation"]
            obs[i, 8] = t["object_error"]
            obs[i, 9] = t["fiducial_area"]
        robot = np.zeros(10)
        bc = None if base_cam is None else np.ascontiguousarray(base_cam, np.float64)
        cb = None if cam_base is None else np.ascontiguousarray(cam_base, np.float64)
        self.lib.hs_map_update(self.h, len(transforms), obs.ctypes.data_as(C.c_void_p), None if bc is None else bc.ctypes.data_as(C.c_void_p),
                               None if cb is None else cb.ctypes.data_as(C.c_void_p), robot.ctypes.data_as(C.c_void_p))
        return robot

    def entries(self):
        out = np.zeros((self.cap, 9))
        n = self.lib.hs_map_entries(self.h, out.ctypes.data_as(C.c_void_p))
        return out[:n]

    def __del__(self):
        try:
            self.lib.hs_map_destroy(self.h)
        except Exception:
            pass
