"""Oracle for the batch SE(3) Gauss-Newton map refinement (the north-star's "batched SE(3) Gauss-Newton").

TEST INFRASTRUCTURE ONLY -- never imported by the product (fiducials_b200/).  PARITY UNPINNED: the reference has no such solver
(its map is the sequential scalar-variance fold of map.cpp:152-320); the graph it does keep is the co-visibility link set
(Fiducial::links, map.cpp:217-222).  This file states the optimisation the CUDA path (fiducials_b200/csrc/fid_map_refine.cu)
must reproduce, with dense normal equations in numpy, and the quality metric the reference ships for maps:
the residual of a plane fit through the fiducial positions (fiducial_slam/scripts/fit_plane.py).

Problem.  Unknowns: the map poses X_i = (R_i, t_i) of the fiducials; entries with variance 0 are fixed (the reference pins its origin
fiducial the same way, map.cpp:477-483).  Measurements: every message (one camera frame) that observes fiducials a and b (a before b
in the message, both in the map) yields the relative pose Z_ab = T_camFid_a^-1 * T_camFid_b with weight w = 1 / (object_error_a +
object_error_b + 1e-9).  Residual of an edge (6-vector):
    e_R = Log(Z_R^T R_a^T R_b)            e_t = R_a^T (t_b - t_a) - Z_t
cost = sum_edges w (|e_R|^2 + lambda_t |e_t|^2).  Gauss-Newton with the local parametrisation R <- R Exp(dtheta), t <- t + dt,
Levenberg damping mu on the diagonal, a fixed number of iterations.
"""
from __future__ import annotations

import math

import numpy as np


def hat(v):
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def so3_exp(w):
    th = float(np.linalg.norm(w))
    K = hat(w)
    if th < 1e-8:
        return np.eye(3) + K + 0.5 * K @ K
    return np.eye(3) + (math.sin(th) / th) * K + ((1.0 - math.cos(th)) / (th * th)) * K @ K


def so3_log(R):
    c = 0.5 * (R[0, 0] + R[1, 1] + R[2, 2] - 1.0)
    c = min(1.0, max(-1.0, c))
    th = math.acos(c)
    v = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    if th < 1e-8:
        return 0.5 * v
    return (th / (2.0 * math.sin(th))) * v


def jr_inv(phi):
    th = float(np.linalg.norm(phi))
    K = hat(phi)
    if th < 1e-8:
        return np.eye(3) + 0.5 * K + (1.0 / 12.0) * K @ K
    c = 1.0 / (th * th) - (1.0 + math.cos(th)) / (2.0 * th * math.sin(th))
    return np.eye(3) + 0.5 * K + c * K @ K


def q_to_R(q):
    x, y, z, w = [float(v) for v in q]
    s = 2.0 / (x * x + y * y + z * z + w * w)
    return np.array([[1 - s * (y * y + z * z), s * (x * y - w * z), s * (x * z + w * y)], [s * (x * y + w * z), 1 - s * (x * x + z * z), s * (y * z - w * x)],
                     [s * (x * z - w * y), s * (y * z + w * x), 1 - s * (x * x + y * y)]])


def build_edges(ids_in_map, messages):
    """[(ia, ib, Z_R, Z_t, w)] with ia / ib indices into ids_in_map; messages = lists of FiducialTransform-like dicts."""
    index = {int(f): i for i, f in enumerate(ids_in_map)}
    edges = []
    for msg in messages:
        obs = [(index[int(o["fiducial_id"])], q_to_R(o["rotation"]), np.array(o["translation"], float), float(o["object_error"])) for o in msg if int(o["fiducial_id"]) in index]
        for a in range(len(obs)):
            for b in range(a + 1, len(obs)):
                ia, Ra, ta, ea = obs[a]
                ib, Rb, tb, eb = obs[b]
                if ia == ib:
                    continue
                edges.append((ia, ib, Ra.T @ Rb, Ra.T @ (tb - ta), 1.0 / (ea + eb + 1e-9)))
    return edges


def edge_terms(Ra, ta, Rb, tb, ZR, Zt):
    eR = so3_log(ZR.T @ Ra.T @ Rb)
    p = Ra.T @ (tb - ta)
    et = p - Zt
    Ji = jr_inv(eR)
    A = np.zeros((6, 6))
    B = np.zeros((6, 6))
    A[:3, :3] = -Ji @ (Rb.T @ Ra)
    A[3:, :3] = hat(p)
    A[3:, 3:] = -Ra.T
    B[:3, :3] = Ji
    B[3:, 3:] = Ra.T
    return np.concatenate([eR, et]), A, B


def cost(R, t, edges, lambda_t=1.0):
    c = 0.0
    for ia, ib, ZR, Zt, w in edges:
        e, _, _ = edge_terms(R[ia], t[ia], R[ib], t[ib], ZR, Zt)
        c += w * (e[:3] @ e[:3] + lambda_t * (e[3:] @ e[3:]))
    return c


def refine(R, t, fixed, edges, iterations=10, damping=1e-6, lambda_t=1.0):
    """Gauss-Newton on copies of R (list of 3x3), t (list of 3-vectors); fixed = bool per node.  Returns R, t, [cost per iteration]."""
    R = [r.copy() for r in R]
    t = [x.copy() for x in t]
    n = len(R)
    W6 = np.array([1, 1, 1, lambda_t, lambda_t, lambda_t], float)
    costs = [cost(R, t, edges, lambda_t)]
    for _ in range(iterations):
        H = np.zeros((6 * n, 6 * n))
        g = np.zeros(6 * n)
        for ia, ib, ZR, Zt, w in edges:
            e, A, B = edge_terms(R[ia], t[ia], R[ib], t[ib], ZR, Zt)
            Wd = w * W6
            sa, sb = slice(6 * ia, 6 * ia + 6), slice(6 * ib, 6 * ib + 6)
            H[sa, sa] += A.T @ (Wd[:, None] * A)
            H[sb, sb] += B.T @ (Wd[:, None] * B)
            H[sa, sb] += A.T @ (Wd[:, None] * B)
            H[sb, sa] += B.T @ (Wd[:, None] * A)
            g[sa] += A.T @ (Wd * e)
            g[sb] += B.T @ (Wd * e)
        free = np.array([not f for f in fixed for _ in range(6)])
        Hf = H[np.ix_(free, free)] + damping * np.eye(int(free.sum()))
        d = np.zeros(6 * n)
        d[free] = np.linalg.solve(Hf, -g[free])
        for i in range(n):
            if fixed[i]:
                continue
            R[i] = R[i] @ so3_exp(d[6 * i : 6 * i + 3])
            t[i] = t[i] + d[6 * i + 3 : 6 * i + 6]
        costs.append(cost(R, t, edges, lambda_t))
    return R, t, costs


def plane_fit_residual(points):
    """fiducial_slam/scripts/fit_plane.py:69-80 with standard_fit.py: orthogonal plane through the centroid (normal = singular vector
    of the smallest singular value of the centred positions); `residual` = 2-norm of the point-plane distances."""
    P = np.asarray(points, float)
    C = P.mean(axis=0)
    _, _, Vt = np.linalg.svd(P - C)
    N = Vt[-1]
    return float(np.linalg.norm((P - C) @ N))
