"""Raw pinned host -> device copy bandwidth on this box (debug aid for the e2e figure)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fiducials_b200 import _lib
from fiducials_b200.node import Detector, default_params
lib = _lib.load()
det = Detector(default_params(), 0, 640, 480, 1)
n = 796 * 2**20
h = C.c_void_p(); d = C.c_void_p()
_lib.check(lib.fid_host_alloc(n, C.byref(h))); _lib.check(lib.fid_device_alloc(det.h, n, C.byref(d)))
C.memset(h, 1, n)
for i in range(4):
    t0 = time.perf_counter(); _lib.check(lib.fid_memcpy_h2d(det.h, d, h, n)); t = time.perf_counter() - t0
    print("h2d %.1f MB in %.2f ms = %.1f GB/s" % (n / 1e6, t * 1e3, n / t / 1e9))
