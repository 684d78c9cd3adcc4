"""The C-ABI library loads on a GPU-less box, exports every symbol include/fiducials_b200.h declares,
and refuses to compute without a CUDA device (no CPU fallback).  CPU only."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    from fiducials_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        g.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    from fiducials_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "fiducials_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(fid_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_struct_layouts_match_header(lib):
    from fiducials_b200 import _lib

    assert C.sizeof(_lib.fid_transform) == 8 + 8 * (3 + 4 + 3 + 3)
    assert C.sizeof(_lib.fid_map_record) == 8 + 8 * 8
    assert C.sizeof(_lib.fid_map_entry) == 8 + 8 * 7
    p = _lib.fid_params()
    assert lib.fid_default_params(C.byref(p)) == 0
    # aruco_detect.cpp:690-727 defaults
    assert (p.dictionary, p.adaptiveThreshWinSizeMin, p.adaptiveThreshWinSizeMax, p.adaptiveThreshWinSizeStep) == (7, 3, 53, 4)
    assert (p.minMarkerPerimeterRate, p.polygonalApproxAccuracyRate, p.cornerRefinementMinAccuracy) == (0.1, 0.01, 0.01)
    assert (p.perspectiveRemovePixelPerCell, p.maxErroneousBitsInBorderRate, p.minMarkerDistanceRate) == (8, 0.04, 0.05)


def test_no_cpu_fallback(lib):
    import torch
    from fiducials_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = _lib.fid_params()
    lib.fid_default_params(C.byref(p))
    h = C.c_void_p()
    assert lib.fid_create(C.byref(p), 0, 640, 480, 1, C.byref(h)) == -2  # FID_ERR_NO_DEVICE
    mp = _lib.fid_map_params()
    lib.fid_map_default_params(C.byref(mp))
    m = C.c_void_p()
    assert lib.fid_map_create(C.byref(mp), 0, C.byref(m)) == -2
    j = C.c_void_p()
    assert lib.fid_jpeg_create(0, 640, 480, 4, 0, C.byref(j)) == -2  # the JPEG ingest has no host decoder to fall back to either


def test_argument_validation(lib):
    from fiducials_b200 import _lib

    p = _lib.fid_params()
    lib.fid_default_params(C.byref(p))
    h = C.c_void_p()
    assert lib.fid_create(None, 0, 640, 480, 1, C.byref(h)) == -1
    assert lib.fid_create(C.byref(p), 0, 4, 4, 1, C.byref(h)) == -1
    assert lib.fid_strerror(-4) == b"unsupported parameter or dictionary"
