"""Pin the CPU oracle against every golden vector the reference's own tests hold.

CPU only.  If these fail the oracle is not trustworthy and no GPU parity claim stands.
"""
import math

import numpy as np
import pytest

from oracle import aruco_oracle as ao
from oracle import slam_oracle as so

# aruco_detect/test/aruco_images_test.cpp:96-109
GOLD_TAG01 = {1: [569.89917, 201.55890, 777.42560, 206.85025, 767.95856, 415.37830, 565.75311, 409.24496]}
# aruco_detect/test/aruco_images_test.cpp:125-147
GOLD_TAG245 = {
    245: [307.68246, 157.38346, 545.10131, 167.04420, 540.11614, 403.27578, 305.64746, 395.01422],
    246: [671.51892, 173.46070, 900.29650, 178.44973, 895.06933, 407.39855, 666.39910, 403.12911],
}


@pytest.mark.parametrize("name,gold", [("tag01", GOLD_TAG01), ("tag245", GOLD_TAG245)])
def test_reference_corner_goldens(kat, name, gold):
    ids, corners = ao.detect(kat.frame(name), 7)
    assert sorted(ids.tolist()) == sorted(gold)
    for i, fid in enumerate(ids.tolist()):
        # ASSERT_FLOAT_EQ is 4 ULP of float32 around 500 => ~1.2e-4
        assert np.abs(corners[i].reshape(-1) - np.array(gold[fid], np.float32)).max() < 1.3e-4
    # the committed oracle outputs are what the live oracle produces
    assert np.array_equal(ids, kat[name + "_ids"])
    assert np.array_equal(corners, kat[name + "_corners"])


def _static_tf(x, y, z, yaw, pitch, roll):
    """tf2_ros static_transform_publisher 'x y z yaw pitch roll' form."""
    return so.TWV.from_qt(so.q_from_rpy(roll, pitch, yaw), [x, y, z])


def test_auto_init_403_golden(kat):
    """fiducial_slam/test/auto_init_403_test.cpp:119-137 through the whole chain."""
    K, D = kat["img403_K"], kat["img403_D"]
    ids, corners, rvecs, tvecs, fields = ao.detect_and_pose(kat.frame("img403"), 7, K, D, 0.145)
    assert ids.tolist() == [403]
    # auto_init_403.test:3-4  base_link -> camera
    T_baseCam = _static_tf(0.035, 0.145, 0.14, -1.479119, -0.041544, -1.204205)
    T_camBase = T_baseCam.inverse()
    m = so.Map()
    robot = None
    for _ in range(14):
        obs = so.observations_from_transforms(fields)
        robot = m.update(obs, T_baseCam, T_camBase)
    assert robot is not None
    q = so.m_to_q(robot.R)
    for v, g in zip(robot.t + q, [0, 0, 0, 0, 0, 0, 1]):
        assert abs(v - g) < 1e-3
    e = m.entries()[0]
    gold = (403, 0.7611, 0.2505, 0.4028, 1.5751, -0.014, -1.546)
    assert e[0] == 403
    for v, g in zip(e[1:], gold[1:]):
        assert abs(v - g) < 1e-3


def test_bag_pair_golden(kat):
    """aruco_images.bag frame -> detect+pose must reproduce aruco_transforms.bag."""
    K, D = kat["bag_K"], kat["bag_D"]
    ids, corners, rvecs, tvecs, fields = ao.detect_and_pose(kat.frame("bag"), 7, K, D, 0.14)
    gold_ids = kat["bag_golden_ids"].tolist()
    assert sorted(ids.tolist()) == sorted(gold_ids)
    by_id = {f["fiducial_id"]: f for f in fields}
    for j, fid in enumerate(gold_ids):
        f = by_id[fid]
        assert np.abs(f["translation"] - kat["bag_golden_t"][j]).max() < 1e-6
        q, gq = f["rotation"], kat["bag_golden_q"][j]
        assert min(np.abs(q - gq).max(), np.abs(q + gq).max()) < 5e-6  # 2017 golden (OpenCV 3.x LM trajectory)
        ge = kat["bag_golden_errs"][j]
        assert abs(f["image_error"] - ge[0]) <= 1e-5 * max(1.0, abs(ge[0]))
        assert abs(f["object_error"] - ge[1]) <= 1e-5 * max(1e-3, abs(ge[1]))
        assert abs(f["fiducial_area"] - ge[2]) <= 1e-5 * abs(ge[2])


def test_create_map_aruco_expectations(kat):
    """fiducial_slam/test/create_map_aruco.xml:26-33 (map_test.py EPSILON 0.1, degrees)."""
    transforms = []
    for j, fid in enumerate(kat["bag_golden_ids"].tolist()):
        ge = kat["bag_golden_errs"][j]
        transforms.append(dict(fiducial_id=fid, translation=kat["bag_golden_t"][j], rotation=kat["bag_golden_q"][j], image_error=ge[0], object_error=ge[1], fiducial_area=ge[2]))
    m = so.Map()
    m.load_entry(111, 0, 0, 0, 0, 0, 0, 0, 0)  # 111_initial_map.txt
    ident = so.TWV.identity()  # static_transform_publisher 0 0 0 0 0 0 1 base_link raspicam
    robot = None
    for _ in range(40):
        robot = m.update(so.observations_from_transforms(transforms), ident, ident)
    q = so.m_to_q(robot.R)
    exp_pose = [0.73, 0.11, 1.0, 0.98, -0.01, -0.18, 0.07]  # x y z qx qy qz qw (map_test.py order)
    got = robot.t + [q[0], q[1], q[2], q[3]]
    for v, g in zip(got, exp_pose):
        assert abs(v - g) < 0.1
    expect = {
        100: (-0.27, 0.82, -1.77, -38.17, -0.15, -149.53),
        103: (-1.86, -0.59, -1.04, 1.70, -23.72, -165.87),
        106: (0.22, -0.0, -0.0, -0.9, 0.24, 0.15),
        107: (0.2, -0.28, -0.0, -0.94, 1.49, -0.92),
        110: (0.7, 0.05, 0.0, 3.38, -4.9, -90),
        111: (0.0, 0.0, 0.0, 0.0, 0.0, 0.0),
        112: (0.0, -0.3, 0.0, -1.0, 0.48, -0.05),
    }
    ents = {e[0]: e for e in m.entries()}
    assert sorted(ents) == sorted(expect)
    for fid, g in expect.items():
        e = ents[fid]
        for v, gv in zip(e[1:4], g[:3]):
            assert abs(v - gv) < 0.1
        for v, gv in zip(e[4:7], g[3:]):
            d = (math.degrees(v) - gv + 180.0) % 360.0 - 180.0
            assert abs(d) < 0.1 * 180 / math.pi + 0.1  # map_test.py compares degrees with EPSILON on radians-ish; keep loose


# --- fiducial_slam/test/transform_var_test.cpp (5 property tests, inequalities only) -----
def _twv(x, var, rpy=(0.0, 0.0, 0.0)):
    return so.TWV.from_qt(so.q_from_rpy(*rpy), [x, 0, 0], var)


def _angle(t):
    """tf2::Quaternion::getAngle = 2 acos(w)."""
    return 2.0 * math.acos(max(-1.0, min(1.0, so.m_to_q(t.R)[3])))


def test_tv_simple_fusion():  # transform_var_test.cpp:15-31
    out = so.average_transforms(_twv(0.0, 0.3), _twv(0.1, 0.3))
    assert 0 < out.t[0] < 0.1 and 0 < out.var < 0.3


def test_tv_simple_rotation_fusion():  # :33-49
    out = so.average_transforms(_twv(0.0, 0.3), _twv(0.0, 0.3, (0.1, 0, 0)))
    assert 0 < _angle(out) < 0.1 and 0 < out.var < 0.3


def test_tv_same_fusion_iterative():  # :51-77
    tv2 = _twv(0.0, 0.3)
    out = so.average_transforms(_twv(0.0, 0.3), tv2)
    assert out.t[0] == 0 and 0 < out.var < 0.3
    for _ in range(10000):
        out.update(tv2)
        assert out.t[0] == 0 and 1e-9 < out.var < 0.3


def test_tv_outlier_with_large_variance():  # :79-107
    out = so.average_transforms(_twv(0.0, 0.2), _twv(0.1, 0.2))
    out = so.average_transforms(out, _twv(0.1, 0.2))
    out = so.average_transforms(out, _twv(1.0, 2.0, (0, 1, 0)))
    assert 0 < out.t[0] < 1.0 and 0 < _angle(out) < 1.0 and 0 < out.var < 1.0
    assert abs(out.t[0] - 0.1) < 0.05 and abs(_angle(out)) < 0.1


def test_tv_different_with_similar_variance():  # :109-126
    out = so.average_transforms(_twv(0.0, 0.1), _twv(1.0, 0.2, (1, 0, 0)))
    assert 0 < out.t[0] < 1.0 and 0 < _angle(out) < 1.0 and out.var > 0.2


def test_map_file_format_reference_fixtures(tmp_path):
    """saveMap / loadMap text format (fiducial_slam/src/map.cpp:541-625).  The two map files the reference ships
    (fiducial_slam/test/111_initial_map.txt, 610_initial_map.txt -- one line each, quoted here) load into the
    poses its launch tests expect, and a map with links survives save -> load -> save byte for byte."""
    from oracle import slam_oracle as so

    m = so.Map()
    assert so.load_map_text(m, "111 0 0 0 0 0 0 0 0\n") == 1  # 111_initial_map.txt
    f = m.fiducials[111]
    assert f.pose.t == [0.0, 0.0, 0.0] and f.pose.var == 0.0 and f.numObs == 0 and not f.links
    assert np.allclose(np.array(f.pose.R), np.eye(3), atol=1e-15)
    m = so.Map()
    assert so.load_map_text(m, "610 0 0 0 180 0 180 0 0\n") == 1  # 610_initial_map.txt
    R = np.array(m.fiducials[610].pose.R)
    assert np.allclose(R, np.diag([-1.0, 1.0, -1.0]), atol=1e-12)  # setRPY(180, 0, 180 deg): a half turn about y
    # invalid lines are skipped (nElems != 9, 10), links are the rest of the line
    m = so.Map()
    text = "7 1.5 -2 0.25 10 20 30 0.001 4 8 9\nnot a line\n8 0 0 0 0 0 0 1 2 7\n9 1 2 3 4 5 6 7\n"
    assert so.load_map_text(m, text) == 2 and sorted(m.fiducials) == [7, 8]
    assert m.fiducials[7].links == {8, 9} and m.fiducials[8].links == {7} and m.fiducials[7].numObs == 4
    out = so.save_map_text(m)
    assert out.splitlines()[0].startswith("7 1.500000 -2.000000 0.250000 10.000000 20.000000 30.000000 0.001000 4 8 9")
    m2 = so.Map()
    so.load_map_text(m2, out)
    assert so.save_map_text(m2) == out
