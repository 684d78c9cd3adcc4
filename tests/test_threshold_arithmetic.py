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
