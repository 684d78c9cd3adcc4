"""Host-side message packing of the published robot pose (map.cpp:337-379): python mirror (fiducials_b200/node.py) against the
oracle restatement (oracle/slam_oracle.py) -- covariance_diagonal override (:110-125,341-345) and the x / y / yaw squash (:369-379)."""
import math

import numpy as np

from fiducials_b200.node import pose_tf, robot_pose_covariance
from oracle import slam_oracle as so


def test_covariance_override_rules():
    assert robot_pose_covariance(0.25) == so.pose_covariance(0.25)
    assert robot_pose_covariance(0.25)[0] == 0.25 and robot_pose_covariance(0.25)[35] == 0.25 and robot_pose_covariance(0.25)[1] == 0.0
    cd = [1, 2, 3, 4, 5, 6]
    assert robot_pose_covariance(0.25, cd) == so.pose_covariance(0.25, cd)
    assert [robot_pose_covariance(0.25, cd)[i * 7] for i in range(6)] == [1, 2, 3, 4, 5, 6]
    for bad in ([1, 2, 3], [1, 2, 0, 4, 5, 6]):  # wrong length / a zero entry: the parameter is ignored (map.cpp:112-124)
        assert robot_pose_covariance(0.25, bad) == so.pose_covariance(0.25, bad) == so.pose_covariance(0.25)


def test_pose_tf_squash_and_odom():
    rng = np.random.default_rng(2)
    for k in range(200):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        t = rng.uniform(-3, 3, 3)
        base = so.TWV.from_qt(q, t, 0.1)
        qo = so.q_from_rpy(0.0, 0.0, rng.uniform(-math.pi, math.pi)) if k % 3 else rng.normal(size=4)
        qo = np.array(qo) / np.linalg.norm(qo)
        to = rng.uniform(-2, 2, 3)
        odom7 = [*to, *qo]
        for with_odom in (False, True):
            for six in (False, True):
                exp = so.published_pose_tf(base, so.TWV.from_qt(qo, to) if with_odom else None, six)
                gt, gq = pose_tf(t, q, odom7 if with_odom else None, six)
                assert np.abs(gt - np.array(exp.t)).max() < 1e-12
                eq = np.array(so.m_to_q(exp.R))
                assert min(np.abs(gq - eq).max(), np.abs(gq + eq).max()) < 1e-12
                if not six:
                    assert gt[2] == 0.0 and abs(gq[0]) < 1e-15 and abs(gq[1]) < 1e-15  # rotation about z only
