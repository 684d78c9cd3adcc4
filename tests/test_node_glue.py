"""The C++ node glue (fiducials_b200/csrc/node_glue.hpp): builds and links on the GPU-less box; on a
GPU box it runs a frame through imageCallback / poseEstimateCallback / transformCallback and the
printed messages are compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "hostsim", "node_glue_main")


def _build():
    import __graft_entry__ as g
    from fiducials_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        g.build()
    libdir = os.path.join(ROOT, "fiducials_b200")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", EXE, os.path.join(ROOT, "tests", "node_glue_main.cpp"), "-L" + libdir, "-lfiducials_b200",
                           "-Wl,-rpath," + libdir, "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-lcudart"])


def test_node_glue_builds_and_fails_loudly_without_gpu(tmp_path):
    import torch

    _build()
    if torch.cuda.is_available():
        pytest.skip("GPU present; see the gpu test")
    raw = tmp_path / "f.bgr"
    raw.write_bytes(bytes(64 * 64 * 3))
    r = subprocess.run([EXE, str(raw), "64", "64", "6", "0.14"], capture_output=True, text=True)
    assert r.returncode == 1 and "no usable CUDA device" in r.stderr


@pytest.mark.gpu
def test_node_glue_messages_match_oracle(tmp_path):
    from fiducials_b200 import synth
    from oracle import aruco_oracle as ao

    _build()
    bgr, truth, K, D, d = synth.make_config_frame("C1", 0)
    raw = tmp_path / "f.bgr"
    raw.write_bytes(bgr.tobytes())
    args = [EXE, str(raw), "640", "480", str(d), "0.14"] + [repr(float(v)) for v in (K[0, 0], K[1, 1], K[0, 2], K[1, 2])] + [repr(float(v)) for v in D[:5]]
    r = subprocess.run(args, capture_output=True, text=True, check=True)
    ids, corners, rv, tv, fields = ao.detect_and_pose(bgr, d, K, D, 0.14)
    V = [l.split() for l in r.stdout.splitlines() if l.startswith("V ")]
    T = [l.split() for l in r.stdout.splitlines() if l.startswith("T ")]
    M = [l.split() for l in r.stdout.splitlines() if l.startswith("M ")]
    assert [int(v[1]) for v in V] == ids.tolist()
    for i, v in enumerate(V):
        assert np.abs(np.array(v[2:], float) - corners[i].reshape(-1)).max() < 1e-3
    for i, t in enumerate(T):
        vals = np.array(t[2:], float)
        assert np.abs(vals[:3] - fields[i]["translation"]).max() < 1e-3 and np.abs(vals[3:7] - fields[i]["rotation"]).max() < 1e-3
    assert sorted(int(m[1]) for m in M) == sorted(ids.tolist())
    # saveMap -> loadMap -> saveMap through the reference's text format: same ids, numObs and links, values to %lf precision
    path = [l.split()[1] for l in r.stdout.splitlines() if l.startswith("F ")][0]
    a = [l.split() for l in open(path).read().splitlines()]
    b = [l.split() for l in open(path + "2").read().splitlines()]
    assert len(a) == len(ids) and [int(l[0]) for l in a] == sorted(ids.tolist())
    for la, lb in zip(a, b):
        assert la[0] == lb[0] and la[8:] == lb[8:] and len(la) == 9 + len(ids) - 1  # every marker of the frame links to the others
        assert np.abs(np.array(la[1:8], float) - np.array(lb[1:8], float)).max() <= 2e-6
    # published pose (map.cpp:337-379): covariance (scalar variance / covariance_diagonal / ignored when a value is 0), 2-D squash
    from oracle import slam_oracle as so

    Cl = [np.array(l.split()[1:], float) for l in r.stdout.splitlines() if l.startswith("C ")]
    assert np.allclose(Cl[0], [0.125, 0.125, 0.125]) and np.allclose(Cl[1], [1, 2, 6]) and np.allclose(Cl[2], [0.125, 0.125, 0.125])
    Pl = [np.array(l.split()[1:], float) for l in r.stdout.splitlines() if l.startswith("P ")]
    base = so.TWV.from_qt([0.18257418583505536, 0.3651483716701107, 0.5477225575051661, 0.7302967433402214], [1.25, -0.5, 0.3], 0.125)
    odom = so.TWV.from_qt([0, 0, 0.3826834323650898, 0.9238795325112867], [0.4, 0.2, 0.0])
    for got, exp in ((Pl[0], so.published_pose_tf(base, odom, False)), (Pl[1], so.published_pose_tf(base, None, True))):
        assert np.abs(got[:3] - np.array(exp.t)).max() < 1e-12
        q = np.array(so.m_to_q(exp.R))
        assert min(np.abs(got[3:] - q).max(), np.abs(got[3:] + q).max()) < 1e-12
