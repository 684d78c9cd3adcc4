#!/usr/bin/env python
"""BASELINE.json config C5: 500-fiducial / 10k-observation map update.  Reports map updates/s of
(a) the reference algorithm on one CPU core (oracle/_ref/libslam_oracle.so, the C++ restatement --
the reference is strictly sequential, so one core is what it can use per map), and
(b) the CUDA fold for 1, 64 and 1024 independent map instances replayed in one launch.
The reference node itself is capped at 20 updates/s by design (fiducial_slam.cpp:116,138-143)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fiducials_b200 import synth  # noqa: E402


def cpu_replay(msgs, seed_entry, reps=5):
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libslam_oracle.so"))
    flat = np.array([[m["fiducial_id"], *m["translation"], *m["rotation"], m["object_error"], m["fiducial_area"]] for msg in msgs for m in msg], np.float64)
    off = np.zeros(len(msgs) + 1, np.int32)
    off[1:] = np.cumsum([len(m) for m in msgs])
    q = synth._q_from_rpy(np.radians(seed_entry[4]), np.radians(seed_entry[5]), np.radians(seed_entry[6]))
    seed = np.array([[seed_entry[0], seed_entry[1], seed_entry[2], seed_entry[3], *q, seed_entry[7]]], np.float64)
    ident = np.array([0, 0, 0, 0, 0, 0, 1], np.float64)
    out = np.zeros((512, 14))
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        n = lib.slam_c_replay(512, 1, seed.ctypes.data_as(C.c_void_p), len(msgs), off.ctypes.data_as(C.c_void_p), flat.ctypes.data_as(C.c_void_p),
                              ident.ctypes.data_as(C.c_void_p), ident.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), None)
        best = min(best, time.perf_counter() - t0)
    cpu_replay.last_entries = out[:n].copy()
    return best, n


def _ref_map(seed_entry):
    """The reference's own Map (oracle/_ref/libmap_ref.so: map.cpp + transform_with_variance.cpp compiled from the reference
    checkout) seeded with the pinned fiducial, or None when the library was not built."""
    from oracle import map_ref

    if not map_ref.available():
        return None
    text = "%d %.17g %.17g %.17g %.17g %.17g %.17g %.17g 0\n" % tuple(seed_entry[:8])
    return map_ref.RefMap(initial_map_text=text)


IDENT7 = [0, 0, 0, 0, 0, 0, 1]


def cpu_fold_kind():
    """("reference" | "port", description) of what cpu_fold_seconds times."""
    from oracle import map_ref

    if map_ref.available():
        return "reference", "oracle/_ref/libmap_ref.so = the reference's fiducial_slam/src/map.cpp + transform_with_variance.cpp compiled unmodified (g++ -O2) against stand-in ROS/tf2 headers, Map::update per message incl. its publishMap / marker messages"
    return "port", "oracle/_ref/libslam_oracle.so (map.cpp / transform_with_variance.cpp restated, g++ -O2)"


def cpu_fold_seconds(msgs, seed_entry, reps=3):
    """Seconds for the whole sequence on one host core (bench.py --workload C5 cpu_baseline)."""
    best = 1e9
    for _ in range(reps):
        ref = _ref_map(seed_entry)
        if ref is None:
            return cpu_replay(msgs, seed_entry, reps)[0]
        t0 = time.perf_counter()
        ref.replay(msgs, IDENT7, IDENT7)
        best = min(best, time.perf_counter() - t0)
        ref.close()
    return best


def cpu_port_info(msgs, seed_entry):
    """The arithmetic-only restatement (oracle/_ref/libslam_oracle.so) beside the compiled reference: the reference's Map::update
    also rebuilds and publishes the whole FiducialMapEntryArray and four rviz markers per observation on every message."""
    t, n = cpu_replay(msgs, seed_entry, 3)
    n_obs = sum(len(m) for m in msgs)
    return {"value": n_obs / t, "unit": "observations/s", "ms_per_sequence": t * 1e3,
            "what": "oracle/_ref/libslam_oracle.so: the same fold without the reference's per-message publishMap / marker messages (map.cpp:629-654, 669-775)"}


def cpu_fold_entries(msgs, seed_entry):
    """Map after the sequence as rows (id, x, y, z, roll, pitch, yaw), ids ascending (publishMap read-out)."""
    ref = _ref_map(seed_entry)
    if ref is not None:
        ref.replay(msgs, IDENT7, IDENT7)
        rows = [(int(r[0]), *r[1:7]) for r in ref.entries()]
        ref.close()
        return sorted(rows)
    from oracle import slam_oracle as so

    cpu_replay(msgs, seed_entry, 1)
    rows = []
    for r in cpu_replay.last_entries:
        R = [list(r[4:7]), list(r[7:10]), list(r[10:13])]
        rr = so.get_rpy(R)
        rows.append((int(r[0]), r[1], r[2], r[3], rr[0], rr[1], rr[2]))
    return sorted(rows)


def main():
    from fiducials_b200.node import FiducialSlam

    msgs, seed_entry = synth.make_c5_sequence(1000, seed=0)
    n_obs = sum(len(m) for m in msgs)
    res = {"config": "C5: 500 fiducials, %d frames, %d observations" % (len(msgs), n_obs)}
    t, n = cpu_replay(msgs, seed_entry)
    res["cpu_1core"] = {"seconds": t, "updates_per_s": len(msgs) / t, "observations_per_s": n_obs / t, "map_entries": n}
    ident = [0, 0, 0, 0, 0, 0, 1]
    one = np.zeros(n_obs, FiducialSlam.TRANSFORM_DTYPE)
    k = 0
    for msg in msgs:
        for m in msg:
            one[k]["fiducial_id"] = m["fiducial_id"]
            one[k]["translation"] = m["translation"]
            one[k]["rotation"] = m["rotation"]
            one[k]["image_error"], one[k]["object_error"], one[k]["fiducial_area"] = m["image_error"], m["object_error"], m["fiducial_area"]
            k += 1
    off1 = np.zeros(len(msgs) + 1, np.int64)
    off1[1:] = np.cumsum([len(m) for m in msgs])
    for ni in (1, 64, 1024, 8192):
        slam = FiducialSlam(max_fiducials=512, n_instances=ni)
        obs = np.tile(one, ni)
        offsets = (off1[None, :] + (np.arange(ni) * n_obs)[:, None]).astype(np.int32)
        best = 1e9
        for rep in range(3):
            for i in range(ni):
                slam.clear(i)
            slam.loadMap([seed_entry], instance=0)
            if ni > 1:  # load the seed into every instance through one export/merge-free path: replicate by loadMap
                for i in range(1, ni):
                    slam.loadMap([seed_entry], instance=i)
            t0 = time.perf_counter()
            slam.replay_raw(offsets, obs, ident, ident)
            best = min(best, time.perf_counter() - t0)
        res["gpu_%d_instances" % ni] = {"seconds": best, "updates_per_s": ni * len(msgs) / best, "observations_per_s": ni * n_obs / best, "map_entries": len(slam.entries(0))}
        slam.close()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
