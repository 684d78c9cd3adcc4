// Batch SE(3) Gauss-Newton refinement of a map instance (the north-star's "batched SE(3) Gauss-Newton").
//
// NEW -- the reference has no such solver: its map is the sequential scalar-variance fold of map.cpp:152-320; what it keeps of the
// pose graph is the co-visibility link set (Fiducial::links, map.cpp:217-222).  Parity is therefore UNPINNED; the statement of the
// problem and the checker are oracle/refine_oracle.py (dense normal equations in numpy), the quality metric is the reference's own
// fiducial_slam/scripts/fit_plane.py.
//
//   unknowns      map poses X_i = (R_i, t_i); entries with variance 0 stay fixed (the reference pins its origin fiducial the same way,
//                 map.cpp:477-483)
//   measurements  every message that observes fiducials a, b (a before b, both in the map): Z_ab = T_camFid_a^-1 T_camFid_b,
//                 weight w = 1 / (object_error_a + object_error_b + 1e-9)
//   residual      e_R = Log(Z_R^T R_a^T R_b),  e_t = R_a^T (t_b - t_a) - Z_t          cost = sum w (|e_R|^2 + lambda_t |e_t|^2)
//   step          Gauss-Newton, R <- R Exp(dtheta), t <- t + dt, Levenberg damping mu
//
// The normal equations are never formed: J^T W J is applied as u = W J p (one thread per edge, 6x6 blocks A, B kept from the
// linearisation) followed by a gather J^T u per node over its incident edges (a warp per node, no atomics) inside a block-Jacobi
// preconditioned conjugate gradient.  The whole refinement -- every Gauss-Newton step and every CG iteration -- is ONE cooperative
// kernel with grid-wide barriers between the phases: no host round trips, no launch gaps (a first version with three launches per
// CG iteration and double atomics spent 105 ms on C5's graph).  At C5's size (500 poses, 45 k edges,
// 432 flops per edge and product) the 6x6 block products are far below anything a tensor-core tile could use; the work
// is latency bound, so the design goal is "no host round trips", not flops.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <unordered_map>
#include <vector>

#include "fid_map_internal.h"

using namespace fid;
namespace cg = cooperative_groups;

namespace {

struct Edge {
    int32_t a, b;     // map slots
    double ZR[9];     // row major
    double Zt[3];
    double w;
};
struct EdgeLin {       // linearisation of one edge
    double e[6];
    double A[36], B[36];  // d e / d x_a, d e / d x_b, row major 6x6, x = [dtheta, dt]
};

__device__ __forceinline__ void mat3_mul(const double* X, const double* Y, double* Z) {  // Z = X Y
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Z[i * 3 + j] = X[i * 3] * Y[j] + X[i * 3 + 1] * Y[3 + j] + X[i * 3 + 2] * Y[6 + j];
}
__device__ __forceinline__ void mat3_tmul(const double* X, const double* Y, double* Z) {  // Z = X^T Y
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Z[i * 3 + j] = X[i] * Y[j] + X[3 + i] * Y[3 + j] + X[6 + i] * Y[6 + j];
}
__device__ __forceinline__ void hat3(const double* v, double* K) {
    K[0] = 0; K[1] = -v[2]; K[2] = v[1];
    K[3] = v[2]; K[4] = 0; K[5] = -v[0];
    K[6] = -v[1]; K[7] = v[0]; K[8] = 0;
}
__device__ void so3_exp(const double* w, double* R) {
    const double th = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    double K[9], K2[9];
    hat3(w, K);
    mat3_mul(K, K, K2);
    double a, b;
    if (th < 1e-8) {
        a = 1.0;
        b = 0.5;
    } else {
        a = sin(th) / th;
        b = (1.0 - cos(th)) / (th * th);
    }
    for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0 ? 1.0 : 0.0) + a * K[i] + b * K2[i];
}
__device__ void so3_log(const double* R, double* phi) {
    double c = 0.5 * (R[0] + R[4] + R[8] - 1.0);
    c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
    const double th = acos(c);
    const double v[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
    const double s = th < 1e-8 ? 0.5 : th / (2.0 * sin(th));
    for (int i = 0; i < 3; i++) phi[i] = s * v[i];
}
__device__ void jr_inv(const double* phi, double* J) {
    const double th = sqrt(phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2]);
    double K[9], K2[9];
    hat3(phi, K);
    mat3_mul(K, K, K2);
    const double c = th < 1e-8 ? 1.0 / 12.0 : 1.0 / (th * th) - (1.0 + cos(th)) / (2.0 * th * sin(th));
    for (int i = 0; i < 9; i++) J[i] = (i % 4 == 0 ? 1.0 : 0.0) + 0.5 * K[i] + c * K2[i];
}

// residual and Jacobian blocks of one edge (oracle/refine_oracle.py::edge_terms)
__device__ void edge_terms(const MapEntry& Xa, const MapEntry& Xb, const Edge& ed, EdgeLin* out) {
    const double *Ra = Xa.pose.R, *Rb = Xb.pose.R;
    double RaTRb[9], E[9], eR[3];
    mat3_tmul(Ra, Rb, RaTRb);
    mat3_tmul(ed.ZR, RaTRb, E);
    so3_log(E, eR);
    const double d[3] = {Xb.pose.t[0] - Xa.pose.t[0], Xb.pose.t[1] - Xa.pose.t[1], Xb.pose.t[2] - Xa.pose.t[2]};
    double p[3];
    for (int i = 0; i < 3; i++) p[i] = Ra[i] * d[0] + Ra[3 + i] * d[1] + Ra[6 + i] * d[2];  // Ra^T d
    for (int i = 0; i < 3; i++) {
        out->e[i] = eR[i];
        out->e[3 + i] = p[i] - ed.Zt[i];
    }
    double Ji[9], RbTRa[9], M[9], P[9];
    jr_inv(eR, Ji);
    mat3_tmul(Rb, Ra, RbTRa);
    mat3_mul(Ji, RbTRa, M);
    hat3(p, P);
    for (int i = 0; i < 36; i++) out->A[i] = out->B[i] = 0.0;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            out->A[i * 6 + j] = -M[i * 3 + j];
            out->A[(3 + i) * 6 + j] = P[i * 3 + j];
            out->A[(3 + i) * 6 + 3 + j] = -Ra[j * 3 + i];  // -Ra^T
            out->B[i * 6 + j] = Ji[i * 3 + j];
            out->B[(3 + i) * 6 + 3 + j] = Ra[j * 3 + i];
        }
}

struct RefineBufs {
    const Edge* edges;
    EdgeLin* lin;
    double* u;                 // per edge: W (A p_a + B p_b), 6 doubles
    const int32_t* node_off;   // CSR of the incident edges of every node: adj[node_off[n] .. node_off[n+1])
    const uint32_t* adj;       // edge index | (1u << 31 if the node is the edge's b side)
    int n_edges, n_nodes;      // n_nodes = fiducials of the instance
    MapEntry* entries;
    double *g, *x, *r, *z, *p, *q, *Minv;  // 6 n vectors; Minv = inverted block diagonal, 36 per node
    double* cost;              // [gn iteration]  (+1: final)
    double *pq, *rz, *rr;      // [gn iteration][pcg iteration + 1] accumulators (zeroed by the host)
    int max_iterations, pcg_iterations;
    double lambda_t, damping, pcg_tol2;
};

__device__ __forceinline__ double wk(const RefineBufs& b, double w, int k) { return k < 3 ? w : w * b.lambda_t; }
__device__ __forceinline__ double warp_sum(double v) {
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// One cooperative kernel runs the whole refinement: Gauss-Newton steps of {linearise (per edge) -> gradient + block diagonal
// (gather per node over its incident edges, no atomics) -> preconditioned conjugate gradient (per edge: u = W J p; per node:
// q = J^T u) -> apply}, the phases separated by grid-wide barriers.  A warp owns a node, its lanes stride over the node's edges.
__global__ void __launch_bounds__(256) k_gn_solve(const RefineBufs b) {
    cg::grid_group grid = cg::this_grid();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    const int slots = b.pcg_iterations + 1;
    for (int it = 0; it <= b.max_iterations; it++) {
        // ---- linearise (the extra pass it == max_iterations only evaluates the final cost) ----
        double c = 0.0;
        for (int i = tid; i < b.n_edges; i += nthreads) {
            const Edge ed = b.edges[i];
            EdgeLin L;
            edge_terms(b.entries[ed.a], b.entries[ed.b], ed, &L);
            if (it < b.max_iterations) b.lin[i] = L;
            for (int k = 0; k < 6; k++) c += wk(b, ed.w, k) * L.e[k] * L.e[k];
        }
        c = warp_sum(c);
        if (lane == 0 && c != 0.0) atomicAdd(&b.cost[it], c);
        if (it == b.max_iterations) break;
        grid.sync();
        double* pq = b.pq + (size_t)it * slots;
        double* rz = b.rz + (size_t)it * slots;
        double* rr = b.rr + (size_t)it * slots;
        // ---- per node: g = J^T W e, H_nn = sum J_n^T W J_n, M^-1 = (H_nn + mu I)^-1, PCG start ----
        for (int n = warp; n < b.n_nodes; n += nwarps) {
            double H[36], g[6];
            for (int i = 0; i < 36; i++) H[i] = 0.0;
            for (int i = 0; i < 6; i++) g[i] = 0.0;
            for (int k = b.node_off[n] + lane; k < b.node_off[n + 1]; k += 32) {
                const uint32_t av = b.adj[k];
                const int ei = (int)(av & 0x7fffffffu);
                const EdgeLin& L = b.lin[ei];
                const double* J = (av >> 31) ? L.B : L.A;
                const double w = b.edges[ei].w;
                for (int col = 0; col < 6; col++) {
                    double gs = 0;
                    for (int kk = 0; kk < 6; kk++) gs += J[kk * 6 + col] * wk(b, w, kk) * L.e[kk];
                    g[col] += gs;
                    for (int row = 0; row <= col; row++) {
                        double hs = 0;
                        for (int kk = 0; kk < 6; kk++) hs += J[kk * 6 + row] * wk(b, w, kk) * J[kk * 6 + col];
                        H[row * 6 + col] += hs;
                    }
                }
            }
            for (int i = 0; i < 6; i++) g[i] = warp_sum(g[i]);
            for (int col = 0; col < 6; col++)
                for (int row = 0; row <= col; row++) {
                    const double v = warp_sum(H[row * 6 + col]);
                    H[row * 6 + col] = v;
                    H[col * 6 + row] = v;
                }
            if (lane == 0) {
                const bool fixed = b.entries[n].pose.var == 0.0;
                double Lm[36], inv[36];
                bool ok = !fixed;
                if (ok) {
                    for (int i = 0; i < 6; i++) H[i * 6 + i] += b.damping;
                    for (int i = 0; i < 6 && ok; i++)
                        for (int j = 0; j <= i; j++) {
                            double s = H[i * 6 + j];
                            for (int k = 0; k < j; k++) s -= Lm[i * 6 + k] * Lm[j * 6 + k];
                            if (i == j) {
                                if (s <= 0.0) {
                                    ok = false;
                                    break;
                                }
                                Lm[i * 6 + i] = sqrt(s);
                            } else {
                                Lm[i * 6 + j] = s / Lm[j * 6 + j];
                            }
                        }
                }
                if (ok) {
                    for (int col = 0; col < 6; col++) {  // solve L L^T y = e_col
                        double y[6];
                        for (int i = 0; i < 6; i++) {
                            double s = i == col ? 1.0 : 0.0;
                            for (int k = 0; k < i; k++) s -= Lm[i * 6 + k] * y[k];
                            y[i] = s / Lm[i * 6 + i];
                        }
                        for (int i = 5; i >= 0; i--) {
                            double s = y[i];
                            for (int k = i + 1; k < 6; k++) s -= Lm[k * 6 + i] * y[k];
                            y[i] = s / Lm[i * 6 + i];
                        }
                        for (int i = 0; i < 6; i++) inv[i * 6 + col] = y[i];
                    }
                } else {
                    for (int i = 0; i < 36; i++) inv[i] = 0.0;
                }
                double a_rz = 0.0, a_rr = 0.0, r6[6];
                for (int i = 0; i < 6; i++) {
                    b.x[6 * n + i] = 0.0;
                    r6[i] = ok ? -g[i] : 0.0;
                    b.r[6 * n + i] = r6[i];
                    a_rr += r6[i] * r6[i];
                }
                for (int i = 0; i < 6; i++) {
                    double s = 0;
                    for (int k = 0; k < 6; k++) s += inv[i * 6 + k] * r6[k];
                    b.z[6 * n + i] = s;
                    b.p[6 * n + i] = s;
                    a_rz += r6[i] * s;
                }
                for (int i = 0; i < 36; i++) b.Minv[36 * n + i] = inv[i];
                atomicAdd(&rz[0], a_rz);
                atomicAdd(&rr[0], a_rr);
            }
        }
        grid.sync();
        // ---- preconditioned conjugate gradient on (J^T W J + mu I) x = -g ----
        const double rr0 = rr[0];
        for (int k = 0; k < b.pcg_iterations; k++) {
            if (rr[k] <= b.pcg_tol2 * rr0) break;  // grid-uniform: every thread reads the same completed sums
            for (int i = tid; i < b.n_edges; i += nthreads) {
                const Edge& ed = b.edges[i];
                const EdgeLin& L = b.lin[i];
                for (int kk = 0; kk < 6; kk++) {
                    double s = 0;
                    for (int cc = 0; cc < 6; cc++) s += L.A[kk * 6 + cc] * b.p[6 * ed.a + cc] + L.B[kk * 6 + cc] * b.p[6 * ed.b + cc];
                    b.u[6 * (size_t)i + kk] = wk(b, ed.w, kk) * s;
                }
            }
            grid.sync();
            for (int n = warp; n < b.n_nodes; n += nwarps) {
                double y[6] = {0, 0, 0, 0, 0, 0};
                if (b.entries[n].pose.var != 0.0)
                    for (int kk = b.node_off[n] + lane; kk < b.node_off[n + 1]; kk += 32) {
                        const uint32_t av = b.adj[kk];
                        const int ei = (int)(av & 0x7fffffffu);
                        const double* J = (av >> 31) ? b.lin[ei].B : b.lin[ei].A;
                        const double* ue = b.u + 6 * (size_t)ei;
                        for (int cc = 0; cc < 6; cc++) {
                            double s = 0;
                            for (int r6 = 0; r6 < 6; r6++) s += J[r6 * 6 + cc] * ue[r6];
                            y[cc] += s;
                        }
                    }
                double dot = 0.0;
                for (int cc = 0; cc < 6; cc++) {
                    y[cc] = warp_sum(y[cc]);
                    if (lane == 0) {
                        const double qv = b.entries[n].pose.var != 0.0 ? y[cc] + b.damping * b.p[6 * n + cc] : 0.0;
                        b.q[6 * n + cc] = qv;
                        dot += b.p[6 * n + cc] * qv;
                    }
                }
                if (lane == 0 && dot != 0.0) atomicAdd(&pq[k], dot);
            }
            grid.sync();
            const double alpha = pq[k] > 0.0 ? rz[k] / pq[k] : 0.0;
            for (int n = tid; n < b.n_nodes; n += nthreads) {
                double r6[6], a_rz = 0.0, a_rr = 0.0;
                for (int i = 0; i < 6; i++) {
                    b.x[6 * n + i] += alpha * b.p[6 * n + i];
                    r6[i] = b.r[6 * n + i] - alpha * b.q[6 * n + i];
                    b.r[6 * n + i] = r6[i];
                    a_rr += r6[i] * r6[i];
                }
                const double* Mi = b.Minv + 36 * n;
                for (int i = 0; i < 6; i++) {
                    double s = 0;
                    for (int kk = 0; kk < 6; kk++) s += Mi[i * 6 + kk] * r6[kk];
                    b.z[6 * n + i] = s;
                    a_rz += r6[i] * s;
                }
                atomicAdd(&rz[k + 1], a_rz);
                atomicAdd(&rr[k + 1], a_rr);
            }
            grid.sync();
            const double beta = rz[k] > 0.0 ? rz[k + 1] / rz[k] : 0.0;
            for (int i = tid; i < 6 * b.n_nodes; i += nthreads) b.p[i] = b.z[i] + beta * b.p[i];
            grid.sync();
        }
        // ---- apply the step: R <- R Exp(dtheta), t <- t + dt ----
        for (int n = tid; n < b.n_nodes; n += nthreads) {
            if (b.entries[n].pose.var == 0.0) continue;
            double dR[9], Rn[9];
            so3_exp(b.x + 6 * n, dR);
            mat3_mul(b.entries[n].pose.R, dR, Rn);
            for (int i = 0; i < 9; i++) b.entries[n].pose.R[i] = Rn[i];
            for (int i = 0; i < 3; i++) b.entries[n].pose.t[i] += b.x[6 * n + 3 + i];
        }
        grid.sync();
    }
}

void host_q_to_R(const double q[4], double R[9]) { q_to_m(q, R); }

}  // namespace

extern "C" int fid_map_refine_default_params(fid_refine_params* p) {
    if (!p) return FID_ERR_INVALID_ARG;
    p->max_iterations = 8;
    p->pcg_iterations = 100;
    p->pcg_tolerance = 1e-10;
    p->damping = 1e-6;
    p->translation_weight = 1.0;
    return FID_OK;
}

extern "C" int fid_map_refine(fid_map* m, int instance, int n_msgs, const int32_t* offsets, const fid_transform* obs, const fid_refine_params* params, fid_refine_stats* stats) {
    if (!m || instance < 0 || instance >= m->p.n_instances || n_msgs < 1 || !offsets || !obs) return FID_ERR_INVALID_ARG;
    fid_refine_params P;
    fid_map_refine_default_params(&P);
    if (params) P = *params;
    if (P.max_iterations < 1 || P.max_iterations > 64 || P.pcg_iterations < 1 || !(P.damping >= 0) || !(P.translation_weight > 0)) return FID_ERR_INVALID_ARG;
    CK(cudaSetDevice(m->device));
    CK(cudaStreamSynchronize(m->stream));
    MapState st;
    CK(cudaMemcpy(&st, m->d_state + instance, sizeof(st), cudaMemcpyDeviceToHost));
    const int cap = m->p.max_fiducials, n = st.n;
    std::vector<MapEntry> ent(n);
    if (n) CK(cudaMemcpy(ent.data(), m->d_entries + (size_t)instance * cap, sizeof(MapEntry) * n, cudaMemcpyDeviceToHost));
    std::unordered_map<int, int> slot;
    int n_free = 0;
    for (int i = 0; i < n; i++) {
        slot[ent[i].id] = i;
        n_free += ent[i].pose.var != 0.0 ? 1 : 0;
    }
    // edges: all pairs of mapped fiducials within a message, in message order
    std::vector<Edge> edges;
    struct O {
        int s;
        double R[9], t[3], oe;
    };
    std::vector<O> cur;
    for (int k = 0; k < n_msgs; k++) {
        if (offsets[k + 1] < offsets[k]) return FID_ERR_INVALID_ARG;
        cur.clear();
        for (int i = offsets[k]; i < offsets[k + 1]; i++) {
            auto it = slot.find(obs[i].fiducial_id);
            if (it == slot.end()) continue;
            O o;
            o.s = it->second;
            host_q_to_R(obs[i].rotation, o.R);
            for (int c = 0; c < 3; c++) o.t[c] = obs[i].translation[c];
            o.oe = obs[i].object_error;
            cur.push_back(o);
        }
        for (size_t a = 0; a < cur.size(); a++)
            for (size_t b = a + 1; b < cur.size(); b++) {
                if (cur[a].s == cur[b].s) continue;
                Edge e;
                e.a = cur[a].s;
                e.b = cur[b].s;
                const double* Ra = cur[a].R;
                const double* Rb = cur[b].R;
                for (int i = 0; i < 3; i++)
                    for (int j = 0; j < 3; j++) e.ZR[i * 3 + j] = Ra[i] * Rb[j] + Ra[3 + i] * Rb[3 + j] + Ra[6 + i] * Rb[6 + j];  // Ra^T Rb
                const double d[3] = {cur[b].t[0] - cur[a].t[0], cur[b].t[1] - cur[a].t[1], cur[b].t[2] - cur[a].t[2]};
                for (int i = 0; i < 3; i++) e.Zt[i] = Ra[i] * d[0] + Ra[3 + i] * d[1] + Ra[6 + i] * d[2];
                e.w = 1.0 / (cur[a].oe + cur[b].oe + 1e-9);
                edges.push_back(e);
            }
    }
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->n_edges = (int)edges.size();
        stats->n_free = n_free;
    }
    if (edges.empty() || n_free == 0) return FID_OK;
    // incident-edge lists (CSR) so that J^T u is a gather per node instead of atomics per edge
    std::vector<int32_t> node_off((size_t)n + 1, 0);
    for (const Edge& e : edges) {
        node_off[e.a + 1]++;
        node_off[e.b + 1]++;
    }
    for (int i = 0; i < n; i++) node_off[i + 1] += node_off[i];
    std::vector<uint32_t> adj(2 * edges.size());
    {
        std::vector<int32_t> fill(node_off.begin(), node_off.end() - 1);
        for (size_t i = 0; i < edges.size(); i++) {
            adj[fill[edges[i].a]++] = (uint32_t)i;
            adj[fill[edges[i].b]++] = (uint32_t)i | 0x80000000u;
        }
    }
    // device buffers (freed on every exit path below)
    Edge* d_edges = nullptr;
    EdgeLin* d_lin = nullptr;
    double* d_vec = nullptr;
    int32_t* d_off = nullptr;
    uint32_t* d_adj = nullptr;
    const size_t nv = (size_t)6 * n;
    const size_t slots = (size_t)P.pcg_iterations + 1;
    const size_t n_scal = (size_t)P.max_iterations + 1 + 3 * (size_t)P.max_iterations * slots;
    const size_t vec_doubles = 6 * nv + 36 * (size_t)n + 6 * edges.size() + n_scal;  // g x r z p q, Minv, u, scalars
    int rc = FID_OK;
    if (cudaMalloc((void**)&d_edges, sizeof(Edge) * edges.size()) != cudaSuccess || cudaMalloc((void**)&d_lin, sizeof(EdgeLin) * edges.size()) != cudaSuccess ||
        cudaMalloc((void**)&d_vec, sizeof(double) * vec_doubles) != cudaSuccess || cudaMalloc((void**)&d_off, sizeof(int32_t) * node_off.size()) != cudaSuccess ||
        cudaMalloc((void**)&d_adj, sizeof(uint32_t) * adj.size()) != cudaSuccess) {
        cudaGetLastError();
        rc = FID_ERR_NO_MEMORY;
    }
    if (rc == FID_OK) {
        cudaStream_t s = m->stream;
        cudaMemcpyAsync(d_edges, edges.data(), sizeof(Edge) * edges.size(), cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(d_off, node_off.data(), sizeof(int32_t) * node_off.size(), cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(d_adj, adj.data(), sizeof(uint32_t) * adj.size(), cudaMemcpyHostToDevice, s);
        cudaMemsetAsync(d_vec, 0, sizeof(double) * vec_doubles, s);
        RefineBufs b{};
        b.edges = d_edges;
        b.lin = d_lin;
        b.node_off = d_off;
        b.adj = d_adj;
        b.n_edges = (int)edges.size();
        b.n_nodes = n;
        b.entries = m->d_entries + (size_t)instance * cap;
        b.g = d_vec;
        b.x = b.g + nv;
        b.r = b.x + nv;
        b.z = b.r + nv;
        b.p = b.z + nv;
        b.q = b.p + nv;
        b.Minv = b.q + nv;
        b.u = b.Minv + 36 * (size_t)n;
        b.cost = b.u + 6 * edges.size();
        b.pq = b.cost + P.max_iterations + 1;
        b.rz = b.pq + (size_t)P.max_iterations * slots;
        b.rr = b.rz + (size_t)P.max_iterations * slots;
        b.max_iterations = P.max_iterations;
        b.pcg_iterations = P.pcg_iterations;
        b.lambda_t = P.translation_weight;
        b.damping = P.damping;
        b.pcg_tol2 = P.pcg_tolerance * P.pcg_tolerance;
        // cooperative launch: every block must be resident (one block of 256 threads per SM at most)
        int dev_sms = 0, per_sm = 0;
        cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, m->device);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gn_solve, 256, 0);
        const int want = std::max(1, std::min({dev_sms * std::max(per_sm, 0), dev_sms, (b.n_edges + 255) / 256}));
        void* args[] = {(void*)&b};
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        cudaEventCreate(&ev0);
        cudaEventCreate(&ev1);
        cudaEventRecord(ev0, s);
        cudaError_t e = per_sm > 0 ? cudaLaunchCooperativeKernel((void*)k_gn_solve, dim3(want), dim3(256), args, 0, s) : cudaErrorLaunchOutOfResources;
        cudaEventRecord(ev1, s);
        std::vector<double> cost_log((size_t)P.max_iterations + 1, 0.0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(cost_log.data(), b.cost, sizeof(double) * cost_log.size(), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) {
            cudaGetLastError();
            fprintf(stderr, "[fiducials_b200] CUDA error %s in fid_map_refine\n", cudaGetErrorString(e));
            rc = FID_ERR_CUDA;
        } else if (stats) {
            stats->initial_cost = cost_log[0];
            stats->final_cost = cost_log[P.max_iterations];
            stats->iterations = P.max_iterations;
            stats->kernel_launches = 1;
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev0, ev1);
            stats->solve_ms = ms;
        }
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
    }
    if (d_off) cudaFree(d_off);
    if (d_adj) cudaFree(d_adj);
    if (d_edges) cudaFree(d_edges);
    if (d_lin) cudaFree(d_lin);
    if (d_vec) cudaFree(d_vec);
    return rc;
}
