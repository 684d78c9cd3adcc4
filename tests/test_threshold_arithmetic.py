"""The integer test k_threshold evaluates per pixel and scale (fiducials_b200/csrc/kernels_threshold.cuh):
    S >= g*k^2 + ((2C-1)*k^2 + 1)/2      with S = box sum over the k x k window (replicate border),
                                         g = the pixel, C = floor(adaptiveThreshConstant)
is exactly cv2.adaptiveThreshold(MEAN_C, THRESH_BINARY_INV, k, constant) -- the call detectMarkers makes per scale
(aruco_detect/src/aruco_detect.cpp:350; window sizes and constant from :690-693) -- for integer, fractional and
negative constants.  CPU only: this pins the arithmetic the GPU parity tests then check bit for bit on the device."""
import cv2
import numpy as np
import pytest


@pytest.mark.parametrize("k", [3, 7, 23, 51])
@pytest.mark.parametrize("const", [7.0, 7.5, 3.999, 0.2, 0.0, -2.5])
def test_integer_threshold_test_equals_cv2(k, const):
    rng = np.random.default_rng(k * 100 + int(const * 10) + 7)
    g = cv2.GaussianBlur(rng.integers(0, 256, (97, 131), dtype=np.uint8), (0, 0), 1.5)
    g[:8, :8] = 255  # saturated corners: the replicate border matters there
    g[-8:, -8:] = 0
    ref = cv2.adaptiveThreshold(g, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY_INV, k, const)
    S = cv2.boxFilter(g.astype(np.float64), -1, (k, k), normalize=False, borderType=cv2.BORDER_REPLICATE).round().astype(np.int64)
    C = int(np.floor(const))
    rhs = g.astype(np.int64) * k * k + ((2 * C - 1) * k * k + 1) // 2  # (2C-1)k^2 is odd: the division is exact
    assert np.array_equal((S >= rhs).astype(np.uint8) * 255, ref)


def test_gray_fixed_point_and_encodings():
    """k_gray: (3735 B + 19235 G + 9798 R + 16384) >> 15 is cv2's BGR2GRAY (8-bit, 15-bit fixed point); an rgb8 frame is
    the same pixels with R and B swapped by cv_bridge::toCvCopy(msg, BGR8) (aruco_detect.cpp:348), a mono8 frame is
    replicated to (g, g, g) whose gray value is g itself -- the identities fid_set_input_encoding relies on."""
    rng = np.random.default_rng(3)
    bgr = rng.integers(0, 256, (64, 80, 3), dtype=np.uint8)
    b, g, r = (bgr[..., i].astype(np.int64) for i in range(3))
    mine = ((3735 * b + 19235 * g + 9798 * r + 16384) >> 15).astype(np.uint8)
    assert np.array_equal(mine, cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY))
    rgb = np.ascontiguousarray(bgr[..., ::-1])
    assert np.array_equal(cv2.cvtColor(cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR), cv2.COLOR_BGR2GRAY), mine)
    mono = rng.integers(0, 256, (64, 80), dtype=np.uint8)
    assert np.array_equal(cv2.cvtColor(cv2.cvtColor(mono, cv2.COLOR_GRAY2BGR), cv2.COLOR_BGR2GRAY), mono)
    assert all(((3735 + 19235 + 9798) * v + 16384) >> 15 == v for v in range(256))
