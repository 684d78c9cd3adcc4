"""Device detect/pose/map logic compiled for the host (tests/hostsim) vs the oracle.  CPU only."""
import cv2
import numpy as np
import pytest

from fiducials_b200 import synth
from oracle import aruco_oracle as ao
from oracle import slam_oracle as so
import hostsim_util as hs


def _frame(cfg, seed):
    bgr, truth, K, D, d = synth.make_config_frame(cfg, seed)
    return bgr, K, D, d


@pytest.mark.parametrize("cfg,seed", [("C1", 0), ("C1", 3), ("C3", 1)])
def test_candidates_match_oracle(cfg, seed):
    bgr, K, D, d = _frame(cfg, seed)
    g = ao.gray(bgr)
    planes = ao.threshold_planes(g)
    quads, scale, clen = hs.candidates(planes, d)
    ref = ao.quad_candidates(g)
    assert len(ref) == len(quads) and len(ref) > 10
    for i, (s, q, n) in enumerate(ref):
        assert s == scale[i] and n == clen[i] and np.array_equal(q, quads[i])


@pytest.mark.parametrize("cfg,seed", [("C1", 0), ("C1", 1), ("C1", 2), ("C3", 0), ("C3", 5), ("C2", 0)])
def test_detect_matches_oracle_synthetic(cfg, seed):
    bgr, K, D, d = _frame(cfg, seed)
    g = ao.gray(bgr)
    ids, corners, stats = hs.detect(g, ao.threshold_planes(g), d)
    rids, rcorners = ao.detect(bgr, d)
    assert ids.tolist() == rids.tolist()  # identical ids in identical order
    assert np.abs(corners - rcorners).max() <= 1e-3
    assert np.abs(corners - rcorners).max() <= 2.5e-4  # typically bit-identical; 1 float32 ulp otherwise


@pytest.mark.parametrize("name", ["tag01", "tag245", "img403", "bag"])
def test_detect_matches_oracle_reference_frames(kat, name):
    bgr = kat.frame(name)
    g = ao.gray(bgr)
    ids, corners, stats = hs.detect(g, ao.threshold_planes(g), 7)
    assert ids.tolist() == kat[name + "_ids"].tolist()
    assert np.abs(corners - kat[name + "_corners"]).max() <= 1e-3


def test_border_rule_matches_cv2():
    """Markers sliding out of the frame: OpenCV 4.13 discards a border-touching quad only AFTER the
    grouping step (together with its group) -- found by black-box probing, see DESIGN.md.  A
    candidate-level border test (OpenCV <= 4.6 behaviour) fails this test."""
    bgr, truth, K, D, d = synth.make_config_frame("C1", 0)
    quads = np.array([q for _, q in truth])
    right = quads[:, :, 0].max()
    n_cases = 0
    for shift in range(int(640 - right) - 6, int(640 - right) + 8):
        fr = np.roll(bgr, shift, axis=1)
        fr[:, :shift] = 190
        g = ao.gray(fr)
        ids, corners, _ = hs.detect(g, ao.threshold_planes(g), d)
        rids, rc = ao.detect(fr, d)
        assert ids.tolist() == rids.tolist(), shift
        if len(ids):
            assert np.abs(corners - rc).max() <= 1e-3
        n_cases += 1
    assert n_cases >= 10


def test_corner_subpix_bit_exact(kat):
    g = ao.gray(kat.frame("bag"))
    _, corners = ao.detect(kat.frame("bag"), 7, cornerRefinementMethod=cv2.aruco.CORNER_REFINE_NONE)
    pts = corners.reshape(-1, 2).copy()
    rng = np.random.default_rng(0)
    pts = np.concatenate([pts, pts + rng.uniform(-1.5, 1.5, pts.shape).astype(np.float32)])
    ref = pts.reshape(-1, 1, 2).copy()
    cv2.cornerSubPix(g, ref, (5, 5), (-1, -1), (cv2.TERM_CRITERIA_MAX_ITER | cv2.TERM_CRITERIA_EPS, 30, 0.01))
    ours = hs.corner_subpix(g, pts)
    assert np.array_equal(ours, ref.reshape(-1, 2)), np.abs(ours - ref.reshape(-1, 2)).max()


@pytest.mark.parametrize("cfg,seed", [("C1", 0), ("C3", 0), ("C2", 1)])
def test_pose_matches_oracle(cfg, seed):
    bgr, K, D, d = _frame(cfg, seed)
    ids, corners, rvecs, tvecs, fields = ao.detect_and_pose(bgr, d, K, D, 0.14)
    out = hs.pose(corners, K, D, np.full(len(ids), 0.14, np.float32), 0.14)
    assert len(ids) > 0
    assert np.abs(out[:, 0:3] - rvecs).max() < 1e-3 and np.abs(out[:, 3:6] - tvecs).max() < 1e-3
    assert np.abs(out[:, 0:3] - rvecs).max() < 2e-5 and np.abs(out[:, 3:6] - tvecs).max() < 2e-5
    for i, f in enumerate(fields):
        assert abs(out[i, 6] - f["image_error"]) <= 1e-6 * max(1.0, f["image_error"]) + 1e-7
        assert abs(out[i, 7] - f["object_error"]) <= 1e-6 * max(1e-3, f["object_error"]) + 1e-9
        assert abs(out[i, 8] - f["fiducial_area"]) <= 1e-9 * f["fiducial_area"]
        assert np.abs(out[i, 9:13] - f["rotation"]).max() < 2e-5


def test_pose_reference_frames(kat):
    for name, flen in [("tag01", 0.145), ("tag245", 0.145), ("img403", 0.145), ("bag", 0.14)]:
        out = hs.pose(kat[name + "_corners"], kat[name + "_K"], kat[name + "_D"], np.full(len(kat[name + "_ids"]), flen, np.float32), flen)
        assert np.abs(out[:, 0:3] - kat[name + "_rvecs"]).max() < 1e-5
        assert np.abs(out[:, 3:6] - kat[name + "_tvecs"]).max() < 1e-5
        assert np.abs(out[:, 6:9] / kat[name + "_errs"] - 1).max() < 1e-4


def _tf7(T):
    return np.array(T.t + so.m_to_q(T.R))


def test_map_matches_slam_oracle(kat):
    transforms = []
    for j, fid in enumerate(kat["bag_golden_ids"].tolist()):
        ge = kat["bag_golden_errs"][j]
        transforms.append(dict(fiducial_id=fid, translation=kat["bag_golden_t"][j], rotation=kat["bag_golden_q"][j], image_error=ge[0], object_error=ge[1], fiducial_area=ge[2]))
    ident = so.TWV.identity()
    m = so.Map()
    m.load_entry(111, 0, 0, 0, 0, 0, 0, 0, 0)
    h = hs.HsMap()
    h.load(111, 0, 0, 0, 0, 0, 0, 0, 0)
    for it in range(40):
        robot = m.update(so.observations_from_transforms(transforms), ident, ident)
        r = h.update(transforms, _tf7(ident), _tf7(ident))
        assert r[0] == 1
        assert np.abs(r[2:5] - np.array(robot.t)).max() < 1e-12
    ents = h.entries()
    ref = m.entries()
    assert len(ents) == len(ref) == 7
    for a, b in zip(ents, ref):
        assert a[0] == b[0]
        assert np.abs(a[1:4] - np.array(b[1:4])).max() < 1e-4 and np.abs(a[4:7] - np.array(b[4:7])).max() < 1e-4  # BASELINE tolerance
        assert np.abs(a[1:7] - np.array(b[1:7])).max() < 1e-10


def test_map_auto_init_matches_slam_oracle(kat):
    K, D = kat["img403_K"], kat["img403_D"]
    out = hs.pose(kat["img403_corners"], K, D, np.full(1, 0.145, np.float32), 0.145)
    fields = [dict(fiducial_id=403, translation=out[0, 3:6], rotation=out[0, 9:13], image_error=out[0, 6], object_error=out[0, 7], fiducial_area=out[0, 8])]
    T_baseCam = so.TWV.from_qt(so.q_from_rpy(-1.204205, -0.041544, -1.479119), [0.035, 0.145, 0.14])
    T_camBase = T_baseCam.inverse()
    m = so.Map()
    h = hs.HsMap()
    for _ in range(14):
        robot = m.update(so.observations_from_transforms(fields), T_baseCam, T_camBase)
        r = h.update(fields, _tf7(T_baseCam), _tf7(T_camBase))
    assert r[0] == 1 and np.abs(r[2:5]).max() < 1e-3
    e = h.entries()[0]
    gold = [403, 0.7611, 0.2505, 0.4028, 1.5751, -0.014, -1.546]  # auto_init_403_test.cpp:128-137
    assert np.abs(e[:7] - np.array(gold)).max() < 1e-3
    assert np.abs(e[1:7] - np.array(m.entries()[0][1:7])).max() < 1e-10


@pytest.mark.parametrize("cfg,seed", [("C1", 0), ("C1", 1), ("C1", 2), ("C3", 0), ("C3", 5), ("C2", 0)])
def test_contour_refinement_matches_oracle(cfg, seed):
    """CORNER_REFINE_CONTOUR (doCornerRefinement = true, cornerRefinementSubPix = false: aruco_detect.cpp:700-711): line fits
    through the candidate's contour, corners = intersections.  float32 normal equations -- reproduced operation by operation."""
    import cv2

    bgr, K, D, d = _frame(cfg, seed)
    g = ao.gray(bgr)
    ids, corners, stats = hs.detect(g, ao.threshold_planes(g), d, refine=2)
    rids, rcorners = ao.detect(bgr, d, cornerRefinementMethod=cv2.aruco.CORNER_REFINE_CONTOUR)
    assert ids.tolist() == rids.tolist()
    assert np.abs(corners - rcorners).max() <= (1e-3 if cfg == "C1" else 2e-2), np.abs(corners - rcorners).max()  # >= 100 points per side: OpenCV hands A^T b to OpenBLAS sgemm, see tests/test_gpu_parity.py::test_corner_refine_contour
