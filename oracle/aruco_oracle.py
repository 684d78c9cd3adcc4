"""CPU oracle for the aruco_detect per-frame path (detect + pose).  TEST INFRASTRUCTURE ONLY.

What the reference executes per frame:

  * ``cv::aruco::detectMarkers(image, dictionary, corners, ids, detectorParams)``
    -- aruco_detect/src/aruco_detect.cpp:350, with the parameters the node sets
    at aruco_detect.cpp:690-727;
  * per marker ``cv::solvePnP`` (default SOLVEPNP_ITERATIVE) -- :247 inside
    ``estimatePoseSingleMarkers`` :223-255, object points from
    ``getSingleMarkerObjectPoints`` :151-161;
  * ``getReprojectionError`` :203-221 (``cv::projectPoints`` :210),
    ``calcFiducialArea`` :179-200, ``dist`` :164-175;
  * pose packing in ``poseEstimateCallback`` :447-495 (axis/angle -> quaternion via
    tf2::Quaternion::setRotation, object_error).

OpenCV is a third-party, un-vendored dependency of the reference (CI pins
4.2 / 3.3; this image has the 4.13 wheel).  The oracle therefore *calls* cv2 for
the three OpenCV entry points and *restates* the node's glue arithmetic.  One
4.13-only parameter (``relativeCornerRefinmentWinSize``) is set to 100 so the
sub-pixel window is always ``cornerRefinementWinSize`` as in the reference's
OpenCV (SURVEY.md fact 3 / probe P4) -- with that, the golden
``aruco_transforms.bag`` is reproduced to <= 1e-6.

The stage-wise functions (``threshold_planes``, ``find_contours`` ...) expose the
cv2 primitives the detector is built from, for stage-by-stage parity tests of the
CUDA kernels.
"""
from __future__ import annotations

import math

import cv2
import numpy as np

# ---------------------------------------------------------------------------------------
# Parameters the reference runs with -- aruco_detect.cpp:690-727 (rosparam defaults).
# ---------------------------------------------------------------------------------------
REFERENCE_PARAMS = dict(
    adaptiveThreshConstant=7.0,  # :690
    adaptiveThreshWinSizeMax=53,  # :691
    adaptiveThreshWinSizeMin=3,  # :692
    adaptiveThreshWinSizeStep=4,  # :693
    cornerRefinementMaxIterations=30,  # :694
    cornerRefinementMinAccuracy=0.01,  # :695
    cornerRefinementWinSize=5,  # :696
    errorCorrectionRate=0.6,  # :716
    minCornerDistanceRate=0.05,  # :717
    markerBorderBits=1,  # :718
    maxErroneousBitsInBorderRate=0.04,  # :719
    minDistanceToBorder=3,  # :720
    minMarkerDistanceRate=0.05,  # :721
    minMarkerPerimeterRate=0.1,  # :722
    maxMarkerPerimeterRate=4.0,  # :723
    minOtsuStdDev=5.0,  # :724
    perspectiveRemoveIgnoredMarginPerCell=0.13,  # :725
    perspectiveRemovePixelPerCell=8,  # :726
    polygonalApproxAccuracyRate=0.01,  # :727
)
# OpenCV >= 4.7 only; not settable by the reference.  100 => window always 5.
ORACLE_ONLY_PARAMS = dict(relativeCornerRefinmentWinSize=100.0, minGroupDistance=0.21)


def reference_detector_params(**overrides) -> "cv2.aruco.DetectorParameters":
    p = cv2.aruco.DetectorParameters()
    for k, v in REFERENCE_PARAMS.items():
        setattr(p, k, v)
    p.cornerRefinementMethod = cv2.aruco.CORNER_REFINE_SUBPIX  # aruco_detect.cpp:700-711 default "SUBPIX"
    p.relativeCornerRefinmentWinSize = ORACLE_ONLY_PARAMS["relativeCornerRefinmentWinSize"]
    for k, v in overrides.items():
        setattr(p, k, v)
    return p


_detectors: dict = {}


def _detector(dict_id: int, key=None, **overrides):
    k = (dict_id, key if key is not None else tuple(sorted(overrides.items())))
    if k not in _detectors:
        d = cv2.aruco.getPredefinedDictionary(dict_id)
        _detectors[k] = cv2.aruco.ArucoDetector(d, reference_detector_params(**overrides))
    return _detectors[k]


def detect(bgr: np.ndarray, dict_id: int, **overrides):
    """imageCallback's detect call (aruco_detect.cpp:350).

    Returns (ids int32[n], corners float32[n,4,2]) in OpenCV's output order.
    """
    corners, ids, _rej = _detector(dict_id, **overrides).detectMarkers(bgr)
    if ids is None or len(ids) == 0:
        return np.zeros((0,), np.int32), np.zeros((0, 4, 2), np.float32)
    return ids.reshape(-1).astype(np.int32), np.stack([c.reshape(4, 2) for c in corners]).astype(np.float32)


def single_marker_object_points(marker_length: float) -> np.ndarray:
    """getSingleMarkerObjectPoints, aruco_detect.cpp:151-161 (float32, TL,TR,BR,BL)."""
    h = np.float32(marker_length) / np.float32(2.0)
    return np.array([[-h, h, 0], [h, h, 0], [h, -h, 0], [-h, -h, 0]], np.float32)


def _dist(p1, p2) -> float:
    """dist(), aruco_detect.cpp:164-175 (float32 points widened to double)."""
    dx = float(p1[0]) - float(p2[0])
    dy = float(p1[1]) - float(p2[1])
    return math.sqrt(dx * dx + dy * dy)


def fiducial_area(pts) -> float:
    """calcFiducialArea, aruco_detect.cpp:179-200 (Heron on two triangles)."""
    p0, p1, p2, p3 = pts
    a1, b1, c1 = _dist(p0, p1), _dist(p0, p3), _dist(p1, p3)
    a2, b2, c2 = _dist(p1, p2), _dist(p2, p3), c1
    s1 = (a1 + b1 + c1) / 2.0
    s2 = (a2 + b2 + c2) / 2.0
    a1 = math.sqrt(s1 * (s1 - a1) * (s1 - b1) * (s1 - c1))
    a2 = math.sqrt(s2 * (s2 - a2) * (s2 - b2) * (s2 - c2))
    return a1 + a2


def reprojection_error(obj, img_pts, K, D, rvec, tvec) -> float:
    """getReprojectionError, aruco_detect.cpp:203-221: mean *squared* pixel error; the
    projected points are vector<Point2f>, i.e. rounded to float32 before the subtraction."""
    proj, _ = cv2.projectPoints(obj, rvec, tvec, K, D)
    proj = proj.reshape(-1, 2).astype(np.float32)
    total = 0.0
    for i in range(len(obj)):
        e = _dist(img_pts[i], proj[i])
        total += e * e
    return total / float(len(obj))


def estimate_pose(ids, corners, K, D, fiducial_len: float, fiducial_lens: dict | None = None):
    """estimatePoseSingleMarkers, aruco_detect.cpp:223-255.  ``fiducial_len`` is cast to float
    at the call site (:425) and per-id overrides are doubles narrowed by the float argument of
    getSingleMarkerObjectPoints (:151)."""
    K = np.asarray(K, np.float64).reshape(3, 3)
    D = np.asarray(D, np.float64).reshape(-1)
    n = len(ids)
    rvecs = np.zeros((n, 3))
    tvecs = np.zeros((n, 3))
    err = np.zeros(n)
    for i in range(n):
        size = float(np.float32(fiducial_len))
        if fiducial_lens and int(ids[i]) in fiducial_lens:
            size = float(fiducial_lens[int(ids[i])])
        obj = single_marker_object_points(size)
        ok, rv, tv = cv2.solvePnP(obj, corners[i].astype(np.float32), K, D)
        rvecs[i] = rv.reshape(3)
        tvecs[i] = tv.reshape(3)
        err[i] = reprojection_error(obj, corners[i], K, D, rv, tv)
    return rvecs, tvecs, err


def pose_fields(ids, corners, rvecs, tvecs, reproj_err, fiducial_len: float):
    """Per-marker FiducialTransform fields, poseEstimateCallback aruco_detect.cpp:447-495.

    Returns list of dict(fiducial_id, translation(3), rotation xyzw(4), image_error,
    object_error, fiducial_area)."""
    out = []
    for i in range(len(ids)):
        rv = rvecs[i]
        angle = math.sqrt(float(rv[0]) ** 2 + float(rv[1]) ** 2 + float(rv[2]) ** 2)  # cv::norm :447
        axis = rv / angle  # :448
        # tf2::Quaternion::setRotation(axis, angle): d = |axis|; s = sin(angle/2)/d
        d = math.sqrt(float(axis[0]) ** 2 + float(axis[1]) ** 2 + float(axis[2]) ** 2)
        s = math.sin(angle * 0.5) / d
        q = (axis[0] * s, axis[1] * s, axis[2] * s, math.cos(angle * 0.5))
        tnorm = math.sqrt(float(tvecs[i][0]) ** 2 + float(tvecs[i][1]) ** 2 + float(tvecs[i][2]) ** 2)
        object_error = (reproj_err[i] / _dist(corners[i][0], corners[i][2])) * (tnorm / fiducial_len)  # :493-495
        out.append(
            dict(
                fiducial_id=int(ids[i]),
                translation=np.array(tvecs[i], np.float64),
                rotation=np.array(q, np.float64),
                image_error=float(reproj_err[i]),
                object_error=float(object_error),
                fiducial_area=fiducial_area(corners[i]),
            )
        )
    return out


def detect_and_pose(bgr, dict_id, K, D, fiducial_len, fiducial_lens=None, **overrides):
    """One full frame of the reference path: imageCallback + poseEstimateCallback."""
    ids, corners = detect(bgr, dict_id, **overrides)
    rvecs, tvecs, err = estimate_pose(ids, corners, K, D, fiducial_len, fiducial_lens)
    return ids, corners, rvecs, tvecs, pose_fields(ids, corners, rvecs, tvecs, err, fiducial_len)


# ---------------------------------------------------------------------------------------
# Stage-wise oracles (cv2 primitives in the order detectMarkers applies them)
# ---------------------------------------------------------------------------------------
def gray(bgr: np.ndarray) -> np.ndarray:
    return cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY) if bgr.ndim == 3 else bgr


def window_sizes(p=REFERENCE_PARAMS):
    n = (p["adaptiveThreshWinSizeMax"] - p["adaptiveThreshWinSizeMin"]) // p["adaptiveThreshWinSizeStep"] + 1
    out = []
    for i in range(n):
        w = p["adaptiveThreshWinSizeMin"] + i * p["adaptiveThreshWinSizeStep"]
        out.append(w + 1 if w % 2 == 0 else w)
    return out


def threshold_planes(g: np.ndarray, p=REFERENCE_PARAMS) -> np.ndarray:
    """13 x H x W uint8 {0,1}: cv2.adaptiveThreshold(MEAN_C, BINARY_INV, k, 7) per scale."""
    return np.stack(
        [
            cv2.adaptiveThreshold(g, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY_INV, k, p["adaptiveThreshConstant"]) >> 7
            for k in window_sizes(p)
        ]
    )


def find_contours(plane: np.ndarray):
    """cv2.findContours(RETR_LIST, CHAIN_APPROX_NONE) -> list of int32[n,2] (x,y), cv2 list order."""
    cs, _ = cv2.findContours(plane, cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)
    return [c.reshape(-1, 2) for c in cs]


def quad_candidates(g: np.ndarray, p=REFERENCE_PARAMS):
    """_findMarkerContours over all scales *before* the border rule: list of (scale,
    int32[4,2] quad in approxPolyDP order, contour length n), concatenated scale-major, contour-list
    order."""
    H, W = g.shape
    out = []
    min_per = int(p["minMarkerPerimeterRate"] * max(W, H))
    max_per = int(p["maxMarkerPerimeterRate"] * max(W, H))
    planes = threshold_planes(g, p)
    for s in range(planes.shape[0]):
        for c in find_contours(planes[s]):
            n = len(c)
            if n < min_per or n > max_per:
                continue
            ap = cv2.approxPolyDP(c.reshape(-1, 1, 2), n * p["polygonalApproxAccuracyRate"], True)
            if len(ap) != 4 or not cv2.isContourConvex(ap):
                continue
            pts = ap.reshape(4, 2).astype(np.float64)
            min_d = float(max(W, H)) ** 2
            for j in range(4):
                d = float(((pts[j] - pts[(j + 1) % 4]) ** 2).sum())
                min_d = min(min_d, d)
            if min_d < (n * p["minCornerDistanceRate"]) ** 2:
                continue
            # NB: no minDistanceToBorder test here.  Black-box probing of cv2 4.13 shows that quads
            # touching the border still take part in the grouping step and are discarded only
            # afterwards, with their whole group (tests/test_oracle_golden.py::test_border_rule_probe).
            out.append((s, ap.reshape(4, 2).astype(np.int32), n))
    return out
