"""Wall-clock breakdown of one bench step (debug aid): python tools/step_breakdown.py FRAMES SLOT"""
import ctypes as C, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fiducials_b200 import _lib, synth
from fiducials_b200.node import MAXM, Detector, FiducialSlam, default_params
nf, slot = int(sys.argv[1]), int(sys.argv[2])
lib = _lib.load()
W, H, nm, d = synth.CONFIGS["C2"]
frames, truths, K, D, _ = synth.make_config_stream("C2", nf, seed=0, realizations=8)
det = Detector(default_params(dictionary=d), 0, W, H, slot)
slam = FiducialSlam(max_fiducials=512)
dptr = C.c_void_p(); _lib.check(lib.fid_device_alloc(det.h, frames.nbytes, C.byref(dptr))); _lib.check(lib.fid_memcpy_h2d(det.h, dptr, frames.ctypes.data_as(C.c_void_p), frames.nbytes))
ident = [0, 0, 0, 0, 0, 0, 1]
for it in range(4):
    t0 = time.perf_counter()
    counts, ids, corners, tfs = det.detect_pose_batch(dptr.value, K, D, 0.14, on_device=True, n_frames=nf, width=W, height=H)
    t1 = time.perf_counter()
    slam.update_frames(counts, tfs, ident, ident)
    t2 = time.perf_counter()
    st = det.last_stage_ms()
    print("iter", it, "detect %.1f ms  map %.1f ms  stage sum %.1f  entries %d markers %d" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, sum(v for k, v in st.items() if not k.startswith("walk_r")), len(slam.entries()), int(counts.sum())))
print({k: round(v, 3) for k, v in det.last_stage_ms().items() if v > 0.002})
print("counters", det.last_counters() if hasattr(det, "last_counters") else None)
