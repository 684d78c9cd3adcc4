"""Device contour logic (fiducials_b200/csrc/contour_walk.cuh, approx_quad.cuh) compiled for the
host and checked against the cv2 primitives detectMarkers is built from.
CPU only; the same functions run inside the CUDA kernels (tests/test_gpu_*.py check those)."""
import cv2
import numpy as np
import pytest

from fiducials_b200 import synth
from oracle import aruco_oracle as ao
import hostsim_util as hs


def _same(a, b):
    return len(a) == len(b) and all(x.shape == y.shape and np.array_equal(x, y) for x, y in zip(a, b))


MODES = (0, 1, 2)  # one-directional walk; the GPU's rounds + segments; the same with tiny passes/checkpoints


@pytest.mark.parametrize("density", [0.2, 0.4, 0.5, 0.6, 0.8])
def test_find_contours_noise(density):
    rng = np.random.default_rng(int(density * 100))
    for shape in [(37, 53), (64, 64), (120, 161)]:
        plane = (rng.random(shape) < density).astype(np.uint8)
        ref = ao.find_contours(plane)
        for mode in MODES:
            ours, _ = hs.find_contours(plane, mode=mode)
            assert _same(ours, ref)


def test_find_contours_blobs_and_edges():
    rng = np.random.default_rng(7)
    img = np.zeros((200, 260), np.uint8)
    for _ in range(40):
        c = (int(rng.integers(0, 260)), int(rng.integers(0, 200)))
        cv2.circle(img, c, int(rng.integers(2, 30)), 1, int(rng.choice([-1, 1, 2, 3])))
    for _ in range(20):
        p0 = (int(rng.integers(-10, 270)), int(rng.integers(-10, 210)))
        p1 = (int(rng.integers(-10, 270)), int(rng.integers(-10, 210)))
        cv2.line(img, p0, p1, int(rng.integers(0, 2)), int(rng.integers(1, 4)))
    img[0, :] = 1  # touches the frame
    img[:, -1] = 1
    for mode in MODES:
        ours, _ = hs.find_contours(img, mode=mode)
        assert _same(ours, ao.find_contours(img))


def test_find_contours_degenerate():
    for plane in [np.zeros((5, 7), np.uint8), np.ones((5, 7), np.uint8), np.eye(9, dtype=np.uint8), np.ones((1, 40), np.uint8), np.ones((33, 1), np.uint8)]:
        for mode in MODES:
            ours, _ = hs.find_contours(plane, mode=mode)
            assert _same(ours, ao.find_contours(plane))


@pytest.mark.parametrize("cfg,seed", [("C1", 0), ("C1", 1), ("C3", 2)])
def test_find_contours_threshold_planes(cfg, seed):
    bgr, *_ = synth.make_config_frame(cfg, seed)
    g = ao.gray(bgr)
    planes = ao.threshold_planes(g)
    for s in (0, 3, 12):
        ref = ao.find_contours(planes[s])
        for mode in MODES:
            ours, nstarts = hs.find_contours(planes[s], mode=mode)
            assert _same(ours, ref)


def _bidir_mismatches(plane, max_len, uni_steps, chunk, fast):
    import ctypes as C
    lib = hs.load()
    lib.hs_walk_bidir_check.restype = C.c_int
    lib.hs_walk_bidir_check.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    plane = np.ascontiguousarray(plane, np.uint8)
    out = np.zeros(2, np.int64)
    bad = lib.hs_walk_bidir_check(plane.ctypes.data, plane.shape[1], plane.shape[0], max_len, uni_steps, chunk, fast, out.ctypes.data)
    return bad, int(out[0]), int(out[1])


def test_bidirectional_walk_equals_one_directional():
    """walk_resume_bidir / walk_bidir_fast (rounds >= 1 of k_walk) decide every start crack exactly like the one-directional
    walk that is pinned to cv2.findContours above: same verdict, same contour length."""
    rng = np.random.default_rng(11)
    planes = [(rng.random((90, 131)) < d).astype(np.uint8) for d in (0.3, 0.5, 0.62, 0.8)]
    img = np.zeros((200, 260), np.uint8)
    for _ in range(40):
        cv2.circle(img, (int(rng.integers(0, 260)), int(rng.integers(0, 200))), int(rng.integers(2, 30)), 1, int(rng.choice([-1, 1, 2, 3])))
    img[0, :] = 1
    img[:, -1] = 1
    planes += [img, np.eye(9, dtype=np.uint8), np.ones((5, 7), np.uint8), np.ones((1, 40), np.uint8), np.ones((33, 1), np.uint8)]
    bgr, *_ = synth.make_config_frame("C1", 3)
    tp = ao.threshold_planes(ao.gray(bgr))
    planes += [tp[1], tp[12]]
    total = 0
    for plane in planes:
        for max_len, uni, chunk in [(1 << 20, 0, 2), (1 << 20, 8, 64), (1 << 20, 3, 16), (40, 8, 6), (41, 0, 1 << 20)]:
            for fast in (0, 1):  # the readable loops and the instruction-count-optimised ones the kernels run
                bad, n_starts, n_canon = _bidir_mismatches(plane, max_len, uni, chunk, fast)
                assert bad == 0
                total += n_canon
    assert total > 1000


def test_length_filter_matches(kat):
    g = ao.gray(kat.frame("tag01"))
    plane = ao.threshold_planes(g)[5]
    lo, hi = int(0.1 * 1280), int(4.0 * 1280)
    ours, _ = hs.find_contours(plane, lo, hi)
    ref = [c for c in ao.find_contours(plane) if lo <= len(c) <= hi]
    assert _same(ours, ref)


def test_approx_poly_vs_cv2(kat):
    """approxPolyDP + 4-gon decision on every long contour of real threshold planes."""
    total = quads = 0
    for name in ("tag245", "bag"):
        g = ao.gray(kat.frame(name))
        planes = ao.threshold_planes(g)
        for s in range(0, 13, 3):
            for c in ao.find_contours(planes[s]):
                n = len(c)
                if n < 128:
                    continue
                ref = cv2.approxPolyDP(c.reshape(-1, 1, 2), n * 0.01, True).reshape(-1, 2)
                k, ours = hs.approx_poly(c, n * 0.01)
                total += 1
                if len(ref) == 4:
                    quads += 1
                    assert k == 4 and np.array_equal(ours, ref)
                elif k == 4:
                    # the restatement may differ on self-touching contours, but a wrong
                    # "4" would create a spurious candidate -- track it.
                    pytest.fail("ours says quad, cv2 says %d vertices" % len(ref))
    assert quads > 20 and total > 200


def test_is_convex_vs_cv2():
    rng = np.random.default_rng(3)
    for _ in range(3000):
        q = rng.integers(0, 40, (4, 2)).astype(np.int32)
        assert hs.is_convex(q) == bool(cv2.isContourConvex(q.reshape(4, 1, 2)))


def test_approx_poly_fuzz_spurred_marker_outlines():
    """VERDICT r1 weak-10 / SURVEY 7.3-4: approxPolyDP on SELF-TOUCHING outlines.  Marker-like quadrilaterals (random pose, 40..400 px)
    are rasterised with 1-pixel spurs, notches and pinches on their borders, every border cv2.findContours returns is run through
    cv2.approxPolyDP and through the device function (approx_quad.cuh, compiled for the host): the 4-gon decision must agree in
    BOTH directions and the four vertices must be identical -- a wrong "4" is a spurious candidate, a missed one a lost marker."""
    rng = np.random.default_rng(17)
    total = quads = touching = 0
    for trial in range(260):
        W = H = 480
        img = np.zeros((H, W), np.uint8)
        side = rng.uniform(40, 400)
        c = np.array([W / 2, H / 2]) + rng.uniform(-20, 20, 2)
        ang = rng.uniform(0, 2 * np.pi)
        base = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], np.float64) * side / 2
        base += rng.uniform(-0.12, 0.12, (4, 2)) * side
        R = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
        quad = (base @ R.T + c).astype(np.int32)
        cv2.fillPoly(img, [quad.reshape(-1, 1, 2)], 255)
        inner = ((base * rng.uniform(0.5, 0.8)) @ R.T + c).astype(np.int32)
        if trial % 3 == 0:
            cv2.fillPoly(img, [inner.reshape(-1, 1, 2)], 0)  # a ring: outer + hole border
        # spurs (1-px lines sticking out / cut in), notches and single-pixel bridges along the border
        border = cv2.findContours(img.copy(), cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)[0]
        for b in border:
            pts = b.reshape(-1, 2)
            for _ in range(int(rng.integers(2, 14))):
                x, y = pts[rng.integers(0, len(pts))]
                dx, dy = rng.integers(-1, 2, 2)
                ln = int(rng.integers(1, 4))
                val = 255 if rng.random() < 0.5 else 0
                for t in range(1, ln + 1):
                    xx, yy = x + dx * t, y + dy * t
                    if 1 <= xx < W - 1 and 1 <= yy < H - 1:
                        img[yy, xx] = val
        plane = (img > 0).astype(np.uint8)
        for cnt in ao.find_contours(plane):
            n = len(cnt)
            if n < 40:
                continue
            uniq = len({(int(p[0]), int(p[1])) for p in cnt})
            touching += uniq < n
            ref = cv2.approxPolyDP(cnt.reshape(-1, 1, 2), n * 0.01, True).reshape(-1, 2)
            k, ours = hs.approx_poly(cnt, n * 0.01)
            total += 1
            if len(ref) == 4:
                quads += 1
                assert k == 4 and np.array_equal(ours, ref), (trial, n, ref.tolist(), k)
            else:
                assert k != 4, (trial, n, len(ref))
    assert total > 300 and quads > 100 and touching > 100, (total, quads, touching)


# ---- table stage of the start pruning (start_prune_table.h) ------------------------------------------------------------------
def test_start_prune_table_is_what_the_generator_proves():
    """The committed table equals a fresh exhaustive enumeration (tools/gen_prune_table.py: every completion of the pixels outside
    the 3 x 5 neighbourhood, walked with the real walk code)."""
    import ctypes as C
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import gen_prune_table

    words, counts = gen_prune_table.generate()
    committed = np.zeros((2, 1024), np.uint32)
    hs.load().hs_committed_prune_table(committed.ctypes.data_as(C.c_void_p))
    assert np.array_equal(words, committed)
    assert counts[0] > 6000 and counts[1] > 5000


@pytest.mark.parametrize("kind", ["noise", "blobs", "threshold"])
def test_table_pruned_starts_give_identical_contours(kind):
    """With the table stage on, every contour (points, order, list order) still equals cv2.findContours -- only starts that can never
    be canonical disappear -- and a third of the starts are gone."""
    import cv2

    rng = np.random.default_rng(5)
    if kind == "noise":
        planes = [(rng.random((97, 131)) < p).astype(np.uint8) for p in (0.3, 0.5, 0.7)]
    elif kind == "blobs":
        planes = []
        for s in range(3):
            img = np.zeros((160, 200), np.uint8)
            for _ in range(60):
                cv2.circle(img, (int(rng.integers(0, 200)), int(rng.integers(0, 160))), int(rng.integers(1, 14)), 1, -1 if rng.random() < 0.7 else 1)
            planes.append(img ^ (rng.random(img.shape) < 0.02).astype(np.uint8))
    else:
        from fiducials_b200 import synth
        from oracle import aruco_oracle as ao

        g = ao.gray(synth.make_config_frame("C3", 1)[0])
        tp = ao.threshold_planes(g)
        planes = [(tp[s] > 0).astype(np.uint8) for s in (0, 5, 12)]
    lib = hs.load()
    try:
        for pl in planes:
            lib.hs_set_start_prune(0)
            _, n0 = hs.find_contours(pl)
            lib.hs_set_start_prune(1)
            ref, _ = cv2.findContours(pl, cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)
            for mode in (0, 1, 2):  # one-directional walk, the GPU's round structure, tiny passes / checkpoints
                got, n1 = hs.find_contours(pl, mode=mode)
                assert len(got) == len(ref)
                for a, b in zip(got, ref):
                    assert np.array_equal(a, b.reshape(-1, 2))
            assert n1 < 0.8 * n0, (n0, n1)
    finally:
        lib.hs_set_start_prune(0)
