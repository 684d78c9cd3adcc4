"""JPEG ingest, CPU: the host entropy decoder (fiducials_b200/csrc/jpeg_host.hpp) and the device arithmetic
(jpeg_math.cuh: inverse DCT, fancy upsampling, colour conversion), compiled for the host by tests/hostsim, against
cv2.imdecode -- the decoder compressed_image_transport puts in front of the reference's imageCallback
(aruco_detect.cpp:332,348; launch default transport `compressed`, aruco_detect.launch:6,28).  Bit-exact."""
import ctypes as C

import cv2
import numpy as np
import pytest

import hostsim_util as hs


def hs_decode(buf, max_w=4096, max_h=4096):
    lib = hs.load()
    data = np.frombuffer(bytes(buf), np.uint8)
    out = np.zeros((max_h, max_w, 3), np.uint8)
    w, h, nv = C.c_int(0), C.c_int(0), C.c_longlong(0)
    flat = np.zeros(max_w * max_h * 3, np.uint8)
    rc = lib.hs_jpeg_decode(data.ctypes.data_as(C.c_void_p), C.c_longlong(len(data)), flat.ctypes.data_as(C.c_void_p), max_w, max_h, C.byref(w), C.byref(h), C.byref(nv))
    if rc != 0:
        return rc, None, 0
    return 0, flat[: w.value * h.value * 3].reshape(h.value, w.value, 3).copy(), nv.value


def scene(h, w, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([128 + 100 * np.sin(xx / 17.0 + seed), 128 + 100 * np.cos(yy / 11.0), 64 + (xx + yy) % 160], -1)
    img += rng.normal(0, 12, img.shape)
    img[h // 4 : h // 2, w // 4 : w // 2] = rng.integers(0, 2, (h // 2 - h // 4, w // 2 - w // 4, 1)) * 255  # hard edges, saturating colours
    return np.clip(img, 0, 255).astype(np.uint8)


SAMPLING = {"444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, "420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420}


@pytest.mark.parametrize("sampling", ["420", "422", "444"])
@pytest.mark.parametrize("quality", [30, 75, 95, 100])
@pytest.mark.parametrize("shape", [(64, 64), (97, 131), (240, 321), (17, 9)])
def test_colour_matches_imdecode(sampling, quality, shape):
    img = scene(shape[0], shape[1], quality + shape[1])
    ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, quality, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SAMPLING[sampling]])
    assert ok
    ref = cv2.imdecode(buf, cv2.IMREAD_COLOR)
    rc, got, nv = hs_decode(buf)
    assert rc == 0
    assert got.shape == ref.shape
    assert np.array_equal(got, ref), "max diff %d at %d px" % (np.abs(got.astype(int) - ref).max(), (got != ref).any(-1).sum())


@pytest.mark.parametrize("rst", [1, 7])
def test_restart_intervals_and_grey(rst):
    img = scene(120, 200, rst)
    ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 85, cv2.IMWRITE_JPEG_RST_INTERVAL, rst])
    assert ok
    rc, got, _ = hs_decode(buf)
    assert rc == 0 and np.array_equal(got, cv2.imdecode(buf, cv2.IMREAD_COLOR))
    grey = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    ok, buf = cv2.imencode(".jpg", grey, [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_RST_INTERVAL, rst])
    rc, got, _ = hs_decode(buf)
    assert rc == 0 and np.array_equal(got, cv2.imdecode(buf, cv2.IMREAD_COLOR))


def test_optimised_huffman_tables():
    img = scene(200, 300, 5)
    ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 92, cv2.IMWRITE_JPEG_OPTIMIZE, 1])
    rc, got, _ = hs_decode(buf)
    assert rc == 0 and np.array_equal(got, cv2.imdecode(buf, cv2.IMREAD_COLOR))


def test_unsupported_and_broken_streams_are_rejected():
    img = scene(64, 64, 1)
    ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    assert hs_decode(buf)[0] == -2  # JPEG_UNSUPPORTED
    ok, buf = cv2.imencode(".jpg", img)
    assert hs_decode(bytes(buf)[:200])[0] < 0
    assert hs_decode(b"not a jpeg at all")[0] < 0


def test_sparse_coefficients_are_smaller_than_the_frame():
    from fiducials_b200 import synth

    frames, _ = synth.make_config_stream("C2", 1, seed=0)[:2]
    ok, buf = cv2.imencode(".jpg", frames[0], [cv2.IMWRITE_JPEG_QUALITY, 90])
    rc, got, nv = hs_decode(buf)
    assert rc == 0 and np.array_equal(got, cv2.imdecode(buf, cv2.IMREAD_COLOR))
    nblk = (1920 // 8) * (1088 // 8) * 3 // 2
    sparse_bytes = nblk * 12 + nv * 2
    assert sparse_bytes < frames[0].nbytes / 3, (sparse_bytes, frames[0].nbytes)


def test_corrupted_streams_never_crash_the_entropy_decoder():
    """Robustness of the host half (it parses bytes that arrive over the network in the reference's deployment): random byte
    flips, truncations and garbage tables must end in a status code or a decoded image, never in an out-of-bounds access."""
    rng = np.random.default_rng(7)
    base = []
    for q, samp in [(40, "420"), (90, "444"), (75, "422")]:
        ok, buf = cv2.imencode(".jpg", scene(48, 80, q), [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SAMPLING[samp], cv2.IMWRITE_JPEG_RST_INTERVAL, 2])
        base.append(np.frombuffer(bytes(buf), np.uint8))
    n_ok = 0
    for it in range(1500):
        b = base[it % 3].copy()
        mode = it % 4
        if mode == 0:  # flips anywhere (headers included)
            for _ in range(int(rng.integers(1, 6))):
                b[int(rng.integers(2, b.size))] = int(rng.integers(0, 256))
        elif mode == 1:  # flips in the entropy-coded part only
            for _ in range(int(rng.integers(1, 20))):
                b[int(rng.integers(b.size // 2, b.size))] = int(rng.integers(0, 256))
        elif mode == 2:  # truncation
            b = b[: int(rng.integers(4, b.size))]
        else:  # a run of 0xFF (markers in the middle of the scan)
            p = int(rng.integers(2, b.size - 8))
            b[p : p + int(rng.integers(1, 8))] = 0xFF
        rc, img, _ = hs_decode(b, max_w=128, max_h=128)
        if rc == 0:
            n_ok += 1
            assert img.shape[2] == 3 and img.shape[0] <= 128 and img.shape[1] <= 128
    assert n_ok > 0  # damage in the scan usually still decodes (to a damaged picture), like libjpeg
