"""Host-side mirror of the reference's two nodes over the C-ABI (include/fiducials_b200.h).

``FiducialsNode`` follows aruco_detect/src/aruco_detect.cpp (FiducialsNode: camInfoCallback :307-330,
imageCallback :332-395, poseEstimateCallback :397-538, ignore list :540-571, length overrides
:627-660); ``FiducialSlam`` follows fiducial_slam/src/fiducial_slam.cpp (transformCallback :79-105)
and Map (map.cpp).  Same names, same argument meaning, same error behaviour (a frame that cannot be
processed is dropped and an empty/None result returned, never an exception from the callbacks);
the arithmetic happens on the GPU inside libfiducials_b200.so.  This python layer exists for the
parity tests and bench.py; the C++ twin for a real ROS node is fiducials_b200/csrc/node_glue.hpp.
"""
from __future__ import annotations

import ctypes as C
import math
import re
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

from . import _lib
from .msgs import (Detection2D, Detection2DArray, ObjectHypothesisWithPose, Fiducial, FiducialArray, FiducialMapEntry, FiducialMapEntryArray, FiducialTransform, FiducialTransformArray, Header, Transform)

MAXM = _lib.FID_MAX_MARKERS


def default_params(**overrides) -> "_lib.fid_params":
    lib = _lib.load()
    p = _lib.fid_params()
    _lib.check(lib.fid_default_params(C.byref(p)))
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError("unknown detector parameter %r" % k)
        setattr(p, k, v)
    return p


def _camera(K, D) -> "_lib.fid_camera":
    cam = _lib.fid_camera()
    K = np.asarray(K, np.float64).reshape(9)
    D = np.asarray(D, np.float64).reshape(-1)
    for i in range(9):
        cam.K[i] = float(K[i])
    for i in range(5):
        cam.D[i] = float(D[i]) if i < len(D) else 0.0  # first five coefficients (:317-323)
    return cam


class Detector:
    """Thin RAII wrapper of fid_detector*."""

    def __init__(self, params=None, device=0, max_width=1920, max_height=1080, max_batch=1):
        self.lib = _lib.load()
        self.params = params if params is not None else default_params()
        self.bpp = 3
        self.h = C.c_void_p()
        _lib.check(self.lib.fid_create(C.byref(self.params), device, max_width, max_height, max_batch, C.byref(self.h)), "fid_create")
        self.max_batch = max_batch
        self.max_width, self.max_height = max_width, max_height

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.fid_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, params):
        _lib.check(self.lib.fid_set_params(self.h, C.byref(params)))
        self.params = params

    ENCODINGS = {"bgr8": 0, "rgb8": 1, "mono8": 2}

    def set_input_encoding(self, encoding: str):
        """fid_set_input_encoding: the camera's own sensor_msgs/Image encoding ("bgr8", "rgb8", "mono8"; frames are then
        [n,H,W,3] or [n,H,W]) instead of the BGR8 copy cv_bridge makes for the reference (aruco_detect.cpp:348)."""
        _lib.check(self.lib.fid_set_input_encoding(self.h, self.ENCODINGS[encoding]), "fid_set_input_encoding")
        self.bpp = 1 if encoding == "mono8" else 3

    def detect(self, bgr: np.ndarray):
        """fid_detect: (ids int32[n], corners float32[n,4,2])."""
        bgr = np.ascontiguousarray(bgr, np.uint8)
        H, W = bgr.shape[:2]
        ids = np.zeros(MAXM, np.int32)
        corners = np.zeros((MAXM, 8), np.float32)
        n = C.c_int(0)
        _lib.check(self.lib.fid_detect(self.h, bgr.ctypes.data_as(C.c_void_p), W, H, W * self.bpp, MAXM, C.byref(n), ids.ctypes.data_as(C.c_void_p),
                                       corners.ctypes.data_as(C.c_void_p)), "fid_detect")
        return ids[: n.value].copy(), corners[: n.value].reshape(-1, 4, 2).copy()

    def pose(self, ids, corners, K, D, fiducial_len, overrides: Optional[Dict[int, float]] = None):
        ids = np.ascontiguousarray(ids, np.int32)
        corners = np.ascontiguousarray(corners, np.float32).reshape(-1, 8)
        n = len(ids)
        out = (_lib.fid_transform * max(n, 1))()
        cam = _camera(K, D)
        oi, ol, no = _overrides(overrides)
        _lib.check(self.lib.fid_pose(self.h, n, ids.ctypes.data_as(C.c_void_p), corners.ctypes.data_as(C.c_void_p), C.byref(cam), float(fiducial_len), no,
                                     oi.ctypes.data_as(C.c_void_p), ol.ctypes.data_as(C.c_void_p), C.cast(out, C.c_void_p)), "fid_pose")
        return [out[i] for i in range(n)]

    def detect_pose_batch(self, frames, K=None, D=None, fiducial_len=0.14, overrides=None, on_device=False, n_frames=None, width=None, height=None):
        """frames: uint8 array [n,H,W,3] (host) or an integer device address when on_device.
        Returns counts[n], ids[n,MAXM], corners[n,MAXM,4,2], transforms (ctypes array n*MAXM or None)."""
        if on_device:
            n, H, W = int(n_frames), int(height), int(width)
            ptr = C.c_void_p(int(frames))
        else:
            frames = np.ascontiguousarray(frames, np.uint8)
            n, H, W = frames.shape[:3]
            ptr = frames.ctypes.data_as(C.c_void_p)
        cam = _camera(K, D) if K is not None else None
        key = (n, cam is not None)
        if getattr(self, "_out_key", None) != key:  # output buffers are reused between calls of the same shape
            self._out = (np.zeros(n, np.int32), np.zeros((n, MAXM), np.int32), np.zeros((n, MAXM, 8), np.float32),
                         (_lib.fid_transform * (n * MAXM))() if cam is not None else None)
            self._out_key = key
        counts, ids, corners, tfs = self._out
        oi, ol, no = _overrides(overrides)
        st = self.lib.fid_detect_pose_batch(self.h, n, ptr, 1 if on_device else 0, W, H, W * self.bpp, W * self.bpp * H, C.byref(cam) if cam is not None else None, float(fiducial_len), no,
                                            oi.ctypes.data_as(C.c_void_p), ol.ctypes.data_as(C.c_void_p), MAXM, counts.ctypes.data_as(C.c_void_p),
                                            ids.ctypes.data_as(C.c_void_p), corners.ctypes.data_as(C.c_void_p), C.cast(tfs, C.c_void_p) if tfs is not None else None)
        _lib.check(st, "fid_detect_pose_batch")
        return counts, ids, corners.reshape(n, MAXM, 4, 2), tfs

    def submit_batch(self, frames, K=None, D=None, fiducial_len=0.14, overrides=None, on_device=False, n_frames=None, width=None, height=None):
        """fid_submit_batch: queue a batch and return at once (see detect_pose_batch for the arguments).  Host frames must
        stay alive and unchanged until the matching collect_batch."""
        if on_device:
            n, H, W = int(n_frames), int(height), int(width)
            ptr = C.c_void_p(int(frames))
        else:
            assert frames.dtype == np.uint8 and frames.flags["C_CONTIGUOUS"]
            n, H, W = frames.shape[:3]
            ptr = frames.ctypes.data_as(C.c_void_p)
        cam = _camera(K, D) if K is not None else None
        oi, ol, no = _overrides(overrides)
        _lib.check(self.lib.fid_submit_batch(self.h, n, ptr, 1 if on_device else 0, W, H, W * self.bpp, W * self.bpp * H, C.byref(cam) if cam is not None else None, float(fiducial_len), no,
                                             oi.ctypes.data_as(C.c_void_p), ol.ctypes.data_as(C.c_void_p)), "fid_submit_batch")
        if not hasattr(self, "_pending"):
            self._pending = []
        self._pending.append((n, cam is not None, frames))

    def collect_batch(self, out=None):
        """fid_collect_batch: results of the oldest submitted batch, as detect_pose_batch returns them.  `out` = a tuple
        returned by an earlier call of the same shape, to reuse its buffers."""
        n, with_pose, _keep = self._pending.pop(0)
        if out is None:
            out = (np.zeros(n, np.int32), np.zeros((n, MAXM), np.int32), np.zeros((n, MAXM, 4, 2), np.float32), (_lib.fid_transform * (n * MAXM))() if with_pose else None)
        counts, ids, corners, tfs = out
        _lib.check(self.lib.fid_collect_batch(self.h, MAXM, counts.ctypes.data_as(C.c_void_p), ids.ctypes.data_as(C.c_void_p), corners.ctypes.data_as(C.c_void_p),
                                              C.cast(tfs, C.c_void_p) if tfs is not None else None), "fid_collect_batch")
        return counts, ids, corners, tfs

    def debug_threshold(self, bgr):
        bgr = np.ascontiguousarray(bgr, np.uint8)
        H, W = bgr.shape[:2]
        gray = np.zeros((H, W), np.uint8)
        planes = np.zeros((16, H, W), np.uint8)
        ns = C.c_int(0)
        _lib.check(self.lib.fid_debug_threshold(self.h, bgr.ctypes.data_as(C.c_void_p), W, H, W * self.bpp, gray.ctypes.data_as(C.c_void_p), planes.ctypes.data_as(C.c_void_p),
                                                C.byref(ns)))
        return gray, planes[: ns.value]

    def debug_candidates(self):
        cap = 4096
        quads = np.zeros((cap, 8), np.int32)
        scale = np.zeros(cap, np.int32)
        clen = np.zeros(cap, np.int32)
        n = C.c_int(0)
        _lib.check(self.lib.fid_debug_candidates(self.h, cap, C.byref(n), quads.ctypes.data_as(C.c_void_p), scale.ctypes.data_as(C.c_void_p),
                                                 clen.ctypes.data_as(C.c_void_p)))
        return quads[: n.value].reshape(-1, 4, 2), scale[: n.value], clen[: n.value]

    STAGES = ["h2d", "threshold", "masks_starts", "walk", "emit", "approx", "group", "identify", "subpix_pose", "_", "d2h", "walk_r0", "walk_r1", "walk_r2", "walk_r3", "walk_r4", "walk_r5", "walk_r6", "walk_r7"]

    def last_stage_ms(self):
        ms = np.zeros(32, np.float32)
        n = C.c_int(0)
        _lib.check(self.lib.fid_last_stage_ms(self.h, ms.ctypes.data_as(C.c_void_p), 32, C.byref(n)))
        return {k: float(ms[i]) for i, k in enumerate(self.STAGES) if k != "_"}

    COUNTERS = ["start_cracks", "contours_in_range", "contour_points", "quad_candidates", "selected", "markers", "kernel_launches"]

    def last_counters(self):
        c = np.zeros(8, np.int64)
        n = C.c_int(0)
        _lib.check(self.lib.fid_last_counters(self.h, c.ctypes.data_as(C.c_void_p), 8, C.byref(n)))
        return {k: int(c[i]) for i, k in enumerate(self.COUNTERS)}


def _overrides(overrides):
    if not overrides:
        return np.zeros(1, np.int32), np.zeros(1, np.float64), 0
    ks = sorted(overrides)
    return np.array(ks, np.int32), np.array([overrides[k] for k in ks], np.float64), len(ks)


class FiducialsNode:
    """aruco_detect's node, minus ROS transport."""

    def __init__(self, dictionary=7, fiducial_len=0.14, ignore_fiducials: Iterable[int] = (), fiducial_len_override: Optional[Dict[int, float]] = None,
                 do_pose_estimation=True, device=0, max_width=1920, max_height=1080, max_batch=1, doCornerRefinement=True, cornerRefinementSubPix=True, **detector_params):
        # doCornerRefinement / cornerRefinementSubPix -> cornerRefinementMethod NONE / SUBPIX / CONTOUR (:700-711, configCallback :274-281)
        detector_params.setdefault("cornerRefinementMethod", (1 if cornerRefinementSubPix else 2) if doCornerRefinement else 0)
        self.fiducial_len = float(fiducial_len)  # :615
        self.doPoseEstimation = do_pose_estimation  # :614
        self.ignoreIds = set(int(i) for i in ignore_fiducials)  # :540-571
        self.fiducialLens = dict(fiducial_len_override or {})  # :627-660
        self.det = Detector(default_params(dictionary=dictionary, **detector_params), device, max_width, max_height, max_batch)
        self.haveCamInfo = False
        self.K = None
        self.D = None
        self.frameId = ""
        self.frameNum = 0
        self.enable_detections = True
        self.vis_msgs = False  # pnh.param vis_msgs (:614)
        self.ids = np.zeros(0, np.int32)
        self.corners = np.zeros((0, 4, 2), np.float32)
        self._last_header = Header()

    # :307-330
    def camInfoCallback(self, K, D, frame_id=""):
        if self.haveCamInfo:
            return
        K = np.asarray(K, np.float64).reshape(3, 3)
        if np.all(K == 0.0):  # :313 "CameraInfo message has invalid intrinsics, K matrix all zeros"
            return
        self.K = K
        self.D = np.asarray(D, np.float64).reshape(-1)[:5]
        self.haveCamInfo = True
        self.frameId = frame_id

    # :332-395
    def imageCallback(self, bgr, header: Optional[Header] = None) -> Optional[FiducialArray]:
        if not self.enable_detections:
            return None  # :334
        header = header or Header()
        fva = FiducialArray(header=Header(header.seq, header.stamp, self.frameId))
        try:
            self.ids, self.corners = self.det.detect(bgr)  # :350
        except _lib.FidError:
            return None  # frame dropped (:389-394)
        for i, fid in enumerate(self.ids.tolist()):
            if fid in self.ignoreIds:
                continue  # :359-364
            c = self.corners[i]
            fva.fiducials.append(Fiducial(fid, 0, *[float(v) for v in c.reshape(-1)]))  # :366-376
        self._last_header = header
        return fva

    # :397-538
    def poseEstimateCallback(self, msg: Optional[FiducialArray] = None) -> Optional[FiducialTransformArray]:
        header = msg.header if msg is not None else self._last_header
        fta = FiducialTransformArray(header=Header(0, header.stamp, self.frameId), image_seq=header.seq)
        self.frameNum += 1
        if not self.doPoseEstimation:
            return fta
        if not self.haveCamInfo:
            return None  # :417-422
        try:
            tfs = self.det.pose(self.ids, self.corners, self.K, self.D, self.fiducial_len, self.fiducialLens)
        except _lib.FidError:
            return fta
        if self.vis_msgs:  # :403, :462-478: vision_msgs/Detection2DArray instead of FiducialTransformArray
            vma = Detection2DArray(header=Header(0, header.stamp, self.frameId))
            for t in tfs:
                if t.fiducial_id in self.ignoreIds:
                    continue
                vma.detections.append(Detection2D([ObjectHypothesisWithPose(int(t.fiducial_id), math.exp(-2.0 * float(t.object_error)), tuple(t.translation), tuple(t.rotation))]))
            return vma
        for t in tfs:
            if t.fiducial_id in self.ignoreIds:
                continue  # :440
            fta.transforms.append(_to_msg(t))
        return fta

    def process_batch(self, frames, first_seq=0) -> List[FiducialTransformArray]:
        """Throughput path: detect + pose for a stack of frames in one C-ABI call."""
        if not self.haveCamInfo:
            return []
        counts, ids, corners, tfs = self.det.detect_pose_batch(frames, self.K, self.D, self.fiducial_len, self.fiducialLens)
        out = []
        for f in range(len(counts)):
            fta = FiducialTransformArray(header=Header(0, (0, 0), self.frameId), image_seq=first_seq + f)
            for m in range(int(counts[f])):
                t = tfs[f * MAXM + m]
                if t.fiducial_id not in self.ignoreIds:
                    fta.transforms.append(_to_msg(t))
            out.append(fta)
        return out


def _to_msg(t) -> FiducialTransform:
    return FiducialTransform(int(t.fiducial_id), Transform(tuple(t.translation), tuple(t.rotation)), float(t.image_error), float(t.object_error), float(t.fiducial_area))


def _tf(T) -> Optional["_lib.fid_tf"]:
    if T is None:
        return None
    t = _lib.fid_tf()
    for i in range(3):
        t.t[i] = float(T[i])
    for i in range(4):
        t.q[i] = float(T[3 + i])
    return t


class FiducialSlam:
    """fiducial_slam's node, minus ROS transport: transformCallback + Map state on the device."""

    def __init__(self, device=0, max_fiducials=512, n_instances=1, weighting_scale=1e9, use_fiducial_area_as_weight=False, read_only_map=False):
        self.lib = _lib.load()
        p = _lib.fid_map_params()
        _lib.check(self.lib.fid_map_default_params(C.byref(p)))
        p.max_fiducials = max_fiducials
        p.n_instances = n_instances
        p.weighting_scale = weighting_scale
        p.use_fiducial_area_as_weight = int(use_fiducial_area_as_weight)
        p.read_only_map = int(read_only_map)
        self.p = p
        self.h = C.c_void_p()
        _lib.check(self.lib.fid_map_create(C.byref(p), device, C.byref(self.h)), "fid_map_create")

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.fid_map_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def loadMap(self, entries: Sequence[Sequence[float]], instance=0):
        """entries: rows of the map file: id x y z roll pitch yaw(deg) variance numObs (map.cpp:556-562)."""
        arr = (_lib.fid_map_file_entry * len(entries))()
        for i, e in enumerate(entries):
            arr[i].fiducial_id = int(e[0])
            arr[i].x, arr[i].y, arr[i].z, arr[i].roll_deg, arr[i].pitch_deg, arr[i].yaw_deg, arr[i].variance = [float(v) for v in e[1:8]]
            arr[i].num_obs = int(e[8]) if len(e) > 8 else 0
        _lib.check(self.lib.fid_map_load(self.h, instance, len(entries), C.cast(arr, C.c_void_p)))

    def links(self, instance=0):
        """fid_map_links: {fiducial_id: sorted linked ids} (Fiducial::links, map.h:87)."""
        cap = self.p.max_fiducials
        pairs = np.zeros((cap * cap, 2), np.int32)
        n = C.c_int(0)
        _lib.check(self.lib.fid_map_links(self.h, instance, cap * cap, C.byref(n), pairs.ctypes.data_as(C.c_void_p)))
        out: Dict[int, List[int]] = {}
        for a, b in pairs[: n.value]:
            out.setdefault(int(a), []).append(int(b))
        return out

    def saveMap(self, filename, instance=0):
        """Map::saveMap, map.cpp:541-566: `id x y z rx ry rz(deg) variance numObs links...`, %lf formatting."""
        links = self.links(instance)
        with open(filename, "w") as fp:
            for e in self.entries(instance):
                rad2deg = lambda a: a * 180.0 / math.pi  # helpers.h:8
                fp.write("%d %f %f %f %f %f %f %f %d" % (e.fiducial_id, e.x, e.y, e.z, rad2deg(e.rx), rad2deg(e.ry), rad2deg(e.rz), e.variance, e.num_obs))
                fp.write("".join(" %d" % k for k in links.get(e.fiducial_id, [])) + "\n")
        return True

    def loadMapFile(self, filename, instance=0):
        """Map::loadMap(filename), map.cpp:572-625.  Returns the number of entries read (invalid lines are skipped
        like the reference's ROS_WARN("Invalid line"))."""
        rows, pairs = [], []
        with open(filename) as fp:
            for line in fp:
                # sscanf("%d %lf ... %d%[^\t\n]"): nine numbers separated by any white space (tabs too), then the links up to a tab
                m = re.match(r"\s*(\S+)\s+(\S+)\s+(\S+)\s+(\S+)\s+(\S+)\s+(\S+)\s+(\S+)\s+(\S+)\s+([+-]?\d+)([^\t\n]*)", line)
                try:
                    if m is None:
                        raise ValueError
                    tok = m.groups()
                    row = [int(tok[0])] + [float(v) for v in tok[1:8]] + [int(tok[8])]
                    links = [int(v) for v in tok[9].split()]
                except ValueError:
                    continue
                rows.append(row)
                pairs += [(row[0], v) for v in links]
        self.loadMap(rows, instance)
        if pairs:
            arr = np.ascontiguousarray(np.array(pairs, np.int32))
            _lib.check(self.lib.fid_map_add_links(self.h, instance, len(pairs), arr.ctypes.data_as(C.c_void_p)))
        return len(rows)

    @staticmethod
    def _obs(transforms):
        arr = (_lib.fid_transform * max(len(transforms), 1))()
        for i, ft in enumerate(transforms):
            if isinstance(ft, FiducialTransform):
                arr[i].fiducial_id = ft.fiducial_id
                arr[i].translation[:] = ft.transform.translation
                arr[i].rotation[:] = ft.transform.rotation
                arr[i].image_error, arr[i].object_error, arr[i].fiducial_area = ft.image_error, ft.object_error, ft.fiducial_area
            else:
                arr[i].fiducial_id = int(ft["fiducial_id"])
                arr[i].translation[:] = [float(v) for v in ft["translation"]]
                arr[i].rotation[:] = [float(v) for v in ft["rotation"]]
                arr[i].image_error, arr[i].object_error, arr[i].fiducial_area = float(ft["image_error"]), float(ft["object_error"]), float(ft["fiducial_area"])
        return arr

    def transformCallback(self, msg, T_baseCam=None, T_camBase=None, instance=0):
        """msg: FiducialTransformArray or list of transforms.  T_* = 7-vectors (x y z qx qy qz qw) the
        host got from tf (map.cpp:258-273), None when the lookup failed.  Returns fid_robot_pose."""
        transforms = msg.transforms if isinstance(msg, FiducialTransformArray) else list(msg)
        arr = self._obs(transforms)
        bc, cb = _tf(T_baseCam), _tf(T_camBase)
        robot = _lib.fid_robot_pose()
        _lib.check(self.lib.fid_map_update(self.h, instance, len(transforms), C.cast(arr, C.c_void_p), C.byref(bc) if bc is not None else None,
                                           C.byref(cb) if cb is not None else None, C.byref(robot)), "fid_map_update")
        return robot

    def replay(self, messages_per_instance, T_baseCam=None, T_camBase=None):
        """messages_per_instance[i] = list of messages (lists of transforms) for instance i; all
        instances must have the same number of messages.  One kernel launch."""
        ni = self.p.n_instances
        assert len(messages_per_instance) == ni
        n_msgs = len(messages_per_instance[0])
        flat = []
        offsets = np.zeros((ni, n_msgs + 1), np.int32)
        for i, msgs in enumerate(messages_per_instance):
            assert len(msgs) == n_msgs
            for k, m in enumerate(msgs):
                offsets[i, k] = len(flat)
                flat.extend(m)
            offsets[i, n_msgs] = len(flat)
        arr = self._obs(flat)
        robots = (_lib.fid_robot_pose * (ni * n_msgs))()
        bc, cb = _tf(T_baseCam), _tf(T_camBase)
        _lib.check(self.lib.fid_map_update_sequence(self.h, n_msgs, offsets.ctypes.data_as(C.c_void_p), C.cast(arr, C.c_void_p), C.byref(bc) if bc is not None else None,
                                                    C.byref(cb) if cb is not None else None, C.cast(robots, C.c_void_p)), "fid_map_update_sequence")
        return robots

    def update_frames(self, counts, tfs, T_baseCam=None, T_camBase=None, instance=0, asynchronous=False):
        """One message per frame straight from Detector.detect_pose_batch's dense output.  With
        asynchronous=True the fold is only enqueued (it overlaps the next detection); sync() waits."""
        counts = np.ascontiguousarray(counts, np.int32)
        bc, cb = _tf(T_baseCam), _tf(T_camBase)
        if asynchronous:
            _lib.check(self.lib.fid_map_update_frames_async(self.h, instance, len(counts), counts.ctypes.data_as(C.c_void_p), C.cast(tfs, C.c_void_p), MAXM,
                                                            C.byref(bc) if bc is not None else None, C.byref(cb) if cb is not None else None), "fid_map_update_frames_async")
            return None
        robot = _lib.fid_robot_pose()
        _lib.check(self.lib.fid_map_update_frames(self.h, instance, len(counts), counts.ctypes.data_as(C.c_void_p), C.cast(tfs, C.c_void_p), MAXM,
                                                  C.byref(bc) if bc is not None else None, C.byref(cb) if cb is not None else None, C.byref(robot)), "fid_map_update_frames")
        return robot

    def sync(self):
        _lib.check(self.lib.fid_map_sync(self.h))

    TRANSFORM_DTYPE = np.dtype([("fiducial_id", "<i4"), ("reserved", "<i4"), ("translation", "<f8", 3), ("rotation", "<f8", 4), ("image_error", "<f8"),
                                ("object_error", "<f8"), ("fiducial_area", "<f8"), ("rvec", "<f8", 3)])

    def replay_raw(self, offsets: np.ndarray, obs: np.ndarray, T_baseCam=None, T_camBase=None):
        """offsets int32 [n_instances, n_msgs+1] into obs (TRANSFORM_DTYPE records); one launch."""
        offsets = np.ascontiguousarray(offsets, np.int32)
        obs = np.ascontiguousarray(obs)
        assert obs.dtype == self.TRANSFORM_DTYPE and offsets.shape[0] == self.p.n_instances
        bc, cb = _tf(T_baseCam), _tf(T_camBase)
        _lib.check(self.lib.fid_map_update_sequence(self.h, offsets.shape[1] - 1, offsets.ctypes.data_as(C.c_void_p), obs.ctypes.data_as(C.c_void_p),
                                                    C.byref(bc) if bc is not None else None, C.byref(cb) if cb is not None else None, None), "fid_map_update_sequence")

    def entries(self, instance=0):
        cap = self.p.max_fiducials
        arr = (_lib.fid_map_entry * cap)()
        n = C.c_int(0)
        _lib.check(self.lib.fid_map_entries(self.h, instance, cap, C.byref(n), C.cast(arr, C.c_void_p)))
        return [arr[i] for i in range(n.value)]

    def publishMap(self, instance=0) -> FiducialMapEntryArray:  # map.cpp:629-654
        return FiducialMapEntryArray([FiducialMapEntry(e.fiducial_id, e.x, e.y, e.z, e.rx, e.ry, e.rz) for e in self.entries(instance)])

    def clear(self, instance=0):
        _lib.check(self.lib.fid_map_clear(self.h, instance))

    def addFiducial(self, fiducial_id, T_mapBase=None, instance=0):
        """add_fiducial service (addFiducialCallback, map.cpp:821-828): handled by the next transformCallback that observes the id
        (handleAddFiducial, map.cpp:489-535).  T_mapBase = the tf lookup map -> base (7-vector) or None when it fails."""
        mb = _tf(T_mapBase)
        _lib.check(self.lib.fid_map_add_fiducial(self.h, instance, int(fiducial_id), C.byref(mb) if mb is not None else None))

    # ---- published pose (host-side message packing, map.cpp:337-379) ----
    covariance_diagonal = None   # rosparam covariance_diagonal (map.cpp:110-125): six non-zero values or ignored
    publish_6dof_pose = False    # map.cpp:107

    def robotPoseCovariance(self, robot):
        return robot_pose_covariance(robot.variance, self.covariance_diagonal)

    def poseTf(self, robot, T_odomBase=None):
        return pose_tf(robot.t[:], robot.q[:], T_odomBase, self.publish_6dof_pose)

    def refine(self, messages, instance=0, **params):
        """fid_map_refine: batch SE(3) Gauss-Newton over the relative poses the recorded messages contain (new).
        messages = list of messages (lists of FiducialTransform-like dicts).  Returns fid_refine_stats."""
        flat, offsets = [], [0]
        for m in messages:
            flat.extend(m)
            offsets.append(len(flat))
        arr = self._obs(flat)
        off = np.ascontiguousarray(np.array(offsets, np.int32))
        p = _lib.fid_refine_params()
        _lib.check(self.lib.fid_map_refine_default_params(C.byref(p)))
        for k, v in params.items():
            if not hasattr(p, k):
                raise AttributeError("unknown refine parameter %r" % k)
            setattr(p, k, v)
        st = _lib.fid_refine_stats()
        _lib.check(self.lib.fid_map_refine(self.h, instance, len(messages), off.ctypes.data_as(C.c_void_p), C.cast(arr, C.c_void_p), C.byref(p), C.byref(st)), "fid_map_refine")
        return st

    # multi-GPU merged view (new): local instances are never overwritten by a merge
    def export_table(self, instance=0) -> np.ndarray:
        cap = self.p.max_fiducials
        arr = (_lib.fid_map_record * cap)()
        _lib.check(self.lib.fid_map_export(self.h, instance, C.cast(arr, C.c_void_p)))
        return np.frombuffer(arr, dtype=np.uint8).copy()

    def merge_tables(self, tables: np.ndarray, n_tables: int):
        """Fold n_tables gathered tables (rank order) into the merged view (rebuilt from scratch: idempotent)."""
        tables = np.ascontiguousarray(tables, np.uint8)
        _lib.check(self.lib.fid_map_merge(self.h, n_tables, tables.ctypes.data_as(C.c_void_p)))

    def merged_entries(self):
        cap = self.p.max_fiducials
        arr = (_lib.fid_map_entry * cap)()
        n = C.c_int(0)
        _lib.check(self.lib.fid_map_merged_entries(self.h, cap, C.byref(n), C.cast(arr, C.c_void_p)))
        return [arr[i] for i in range(n.value)]

    def adopt_merged(self, instance=0):
        _lib.check(self.lib.fid_map_adopt_merged(self.h, instance))

    @property
    def table_bytes(self) -> int:
        return self.p.max_fiducials * C.sizeof(_lib.fid_map_record)

    def cuda_stream(self) -> int:
        st = C.c_void_p()
        _lib.check(self.lib.fid_map_stream(self.h, C.byref(st)))
        return st.value or 0

    def export_async(self, device_ptr: int, instance=0):
        _lib.check(self.lib.fid_map_export_async(self.h, instance, C.c_void_p(device_ptr)))

    def merge_device_async(self, device_ptr: int, n_tables: int):
        _lib.check(self.lib.fid_map_merge_device_async(self.h, n_tables, C.c_void_p(device_ptr)))


def robot_pose_covariance(variance, covariance_diagonal=None):
    """Row-major 6x6 covariance of the PoseWithCovarianceStamped on /fiducial_pose: toPose fills the diagonal with the scalar
    variance (transform_with_variance.h:69-84); a covariance_diagonal of six non-zero values overrides it (map.cpp:110-125,341-345)."""
    cov = [0.0] * 36
    diag = [float(variance)] * 6
    if covariance_diagonal is not None and len(covariance_diagonal) == 6 and all(v != 0 for v in covariance_diagonal):
        diag = [float(v) for v in covariance_diagonal]
    for i in range(6):
        cov[i * 6 + i] = diag[i]
    return cov


def pose_tf(t, q, T_odomBase=None, publish_6dof_pose=False):
    """The transform broadcast as map -> odom (or map -> base): outPose = basePose * odom^-1 when the odom lookup succeeded
    (map.cpp:351-365), squashed to x, y, yaw unless publish_6dof_pose (map.cpp:369-379).  Returns (t[3], q_xyzw[4])."""
    t = np.array(t, np.float64)
    R = _q_to_R(q)
    if T_odomBase is not None:
        Ro = _q_to_R(T_odomBase[3:7])
        to = np.array(T_odomBase[:3], np.float64)
        Rinv = Ro.T
        tinv = Rinv @ (-to)
        t = R @ tinv + t
        R = R @ Rinv
    if not publish_6dof_pose:
        t[2] = 0.0
        yaw = _get_rpy(R)[2]
        R = _set_rpy(0.0, 0.0, yaw)
    return t, _R_to_q(R)


# ---- tf2 LinearMath pieces used by the host-side message packing above ------------------------------------
def _q_to_R(q):  # tf2::Matrix3x3::setRotation
    x, y, z, w = [float(v) for v in q]
    d = x * x + y * y + z * z + w * w
    s = 2.0 / d
    xs, ys, zs = x * s, y * s, z * s
    wx, wy, wz, xx, xy, xz, yy, yz, zz = w * xs, w * ys, w * zs, x * xs, x * ys, x * zs, y * ys, y * zs, z * zs
    return np.array([[1.0 - (yy + zz), xy - wz, xz + wy], [xy + wz, 1.0 - (xx + zz), yz - wx], [xz - wy, yz + wx, 1.0 - (xx + yy)]])


def _R_to_q(m):  # tf2::Matrix3x3::getRotation
    tr = m[0][0] + m[1][1] + m[2][2]
    q = [0.0] * 4
    if tr > 0.0:
        s = math.sqrt(tr + 1.0)
        q[3] = s * 0.5
        s = 0.5 / s
        q[0], q[1], q[2] = (m[2][1] - m[1][2]) * s, (m[0][2] - m[2][0]) * s, (m[1][0] - m[0][1]) * s
    else:
        i = (2 if m[1][1] < m[2][2] else 1) if m[0][0] < m[1][1] else (2 if m[0][0] < m[2][2] else 0)
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0)
        q[i] = s * 0.5
        s = 0.5 / s
        q[3] = (m[k][j] - m[j][k]) * s
        q[j] = (m[j][i] + m[i][j]) * s
        q[k] = (m[k][i] + m[i][k]) * s
    return np.array(q)


def _get_rpy(m):  # tf2::Matrix3x3::getRPY
    if abs(m[2][0]) >= 1.0:
        delta = math.atan2(m[2][1], m[2][2])
        return (delta, math.pi / 2.0, 0.0) if m[2][0] < 0 else (delta, -math.pi / 2.0, 0.0)
    pitch = -math.asin(m[2][0])
    c = math.cos(pitch)
    return math.atan2(m[2][1] / c, m[2][2] / c), pitch, math.atan2(m[1][0] / c, m[0][0] / c)


def _set_rpy(roll, pitch, yaw):  # tf2::Matrix3x3::setRPY
    ci, cj, ch = math.cos(roll), math.cos(pitch), math.cos(yaw)
    si, sj, sh = math.sin(roll), math.sin(pitch), math.sin(yaw)
    cc, cs, sc, ss = ci * ch, ci * sh, si * ch, si * sh
    return np.array([[cj * ch, sj * sc - cs, sj * cc + ss], [cj * sh, sj * ss + cc, sj * cs - sc], [-sj, cj * si, cj * ci]])


class JpegDecoder:
    """fid_jpeg_*: the cv::imdecode that compressed_image_transport runs in front of imageCallback (aruco_detect.cpp:332,348;
    launch default transport `compressed`), with the Huffman stage on host threads and everything after it on the device.
    decode(list of bytes-like JPEG streams, device pointer) writes BGR8 frames into device memory
    (FiducialsNode.detector_device_alloc / fid_device_alloc) for submit_batch(..., on_device=True)."""

    def __init__(self, max_width, max_height, max_batch, device=0, n_threads=0):
        self.lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self.lib.fid_jpeg_create(device, max_width, max_height, max_batch, n_threads, C.byref(h)), "fid_jpeg_create")
        self.h = h

    def decode(self, streams, device_ptr, width, height, row_stride=None, frame_stride=None, sync=True):
        n = len(streams)
        bufs = [np.frombuffer(bytes(s) if not isinstance(s, np.ndarray) else s, np.uint8) for s in streams]
        ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
        sizes = (C.c_size_t * n)(*[b.size for b in bufs])
        status = np.zeros(n, np.int32)
        rs = row_stride or width * 3
        fs = frame_stride or rs * height
        _lib.check(self.lib.fid_jpeg_decode_batch(self.h, n, ptrs, sizes, width, height, C.c_void_p(int(device_ptr)), rs, fs, status.ctypes.data_as(C.c_void_p)), "fid_jpeg_decode_batch")
        self._keep = bufs
        if sync:
            self.sync()
        return status

    def sync(self):
        _lib.check(self.lib.fid_jpeg_sync(self.h))

    def stats(self):
        a, b, c = C.c_double(0), C.c_double(0), C.c_double(0)
        _lib.check(self.lib.fid_jpeg_last_stats(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return {"host_decode_ms": a.value, "h2d_bytes": b.value, "device_ms": c.value}

    def close(self):
        if self.h:
            self.lib.fid_jpeg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
