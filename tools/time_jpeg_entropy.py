#!/usr/bin/env python
"""Host entropy stage of the JPEG ingest (jpeg_host.hpp) on one core, next to cv2.imdecode's full decode, on a C2 frame."""
import ctypes as C
import os
import sys
import time

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hostsim_util as hs  # noqa: E402
from fiducials_b200 import synth  # noqa: E402

lib = hs.load()
lib.hs_jpeg_entropy.restype = C.c_longlong
frame = synth.make_config_stream("C2", 1, seed=0)[0][0]
for q in (80, 90):
    buf = cv2.imencode(".jpg", frame, [cv2.IMWRITE_JPEG_QUALITY, q])[1]
    data = np.frombuffer(bytes(buf), np.uint8)
    lib.hs_jpeg_entropy(data.ctypes.data_as(C.c_void_p), C.c_longlong(data.size), 2)
    t0 = time.perf_counter()
    nv = lib.hs_jpeg_entropy(data.ctypes.data_as(C.c_void_p), C.c_longlong(data.size), 10)
    t = (time.perf_counter() - t0) / 10
    cv2.setNumThreads(1)
    cv2.imdecode(buf, cv2.IMREAD_COLOR)
    t0 = time.perf_counter()
    for _ in range(10):
        cv2.imdecode(buf, cv2.IMREAD_COLOR)
    tc = (time.perf_counter() - t0) / 10
    print("quality %d: %d bytes, %d values; entropy stage %.2f ms (%.0f MB/s), cv2.imdecode %.2f ms" % (q, data.size, nv, t * 1e3, data.size / t / 1e6, tc * 1e3))
