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
// The normal equations are never formed: J^T W J is applied edge by edge (one thread per edge, 6x6 blocks A, B kept from the
// linearisation, double atomics into the 6 N vector) inside a block-Jacobi preconditioned conjugate gradient whose scalars stay on
// the device -- a GN step is a fixed sequence of launches with no host synchronisation.  At C5's size (500 poses, 45 k edges,
// 432 flops per edge and product) the 6x6 block products are far below anything a tensor-core tile could use; the work
// is latency bound, so the design goal is "no host round trips", not flops.
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <unordered_map>
#include <vector>

#include "fid_map_internal.h"

using namespace fid;

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
    int n_edges, n_nodes;      // n_nodes = fiducials of the instance
    MapEntry* entries;
    double *g, *x, *r, *z, *p, *q, *Hd;  // 6 n vectors; Hd = block diagonal, 36 per node
    double* scal;              // [0] cost  [1] rz  [2] pq  [3] rz_new  [4] r0 norm^2  [5] r norm^2
    double lambda_t, damping, pcg_tol2;
};

__device__ __forceinline__ double wk(const RefineBufs& b, double w, int k) { return k < 3 ? w : w * b.lambda_t; }

// linearise: residuals, Jacobian blocks, gradient g = J^T W e, block diagonal of J^T W J, cost
__global__ void k_gn_linearize(const RefineBufs b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double c = 0.0;
    if (i < b.n_edges) {
        const Edge ed = b.edges[i];
        EdgeLin L;
        edge_terms(b.entries[ed.a], b.entries[ed.b], ed, &L);
        b.lin[i] = L;
        for (int k = 0; k < 6; k++) c += wk(b, ed.w, k) * L.e[k] * L.e[k];
        for (int col = 0; col < 6; col++) {
            double ga = 0, gb = 0;
            for (int k = 0; k < 6; k++) {
                ga += L.A[k * 6 + col] * wk(b, ed.w, k) * L.e[k];
                gb += L.B[k * 6 + col] * wk(b, ed.w, k) * L.e[k];
            }
            atomicAdd(&b.g[6 * ed.a + col], ga);
            atomicAdd(&b.g[6 * ed.b + col], gb);
            for (int row = 0; row < 6; row++) {
                double ha = 0, hb = 0;
                for (int k = 0; k < 6; k++) {
                    ha += L.A[k * 6 + row] * wk(b, ed.w, k) * L.A[k * 6 + col];
                    hb += L.B[k * 6 + row] * wk(b, ed.w, k) * L.B[k * 6 + col];
                }
                atomicAdd(&b.Hd[36 * ed.a + row * 6 + col], ha);
                atomicAdd(&b.Hd[36 * ed.b + row * 6 + col], hb);
            }
        }
    }
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0 && c != 0.0) atomicAdd(&b.scal[0], c);
}

// per node: invert the damped 6x6 diagonal block (Cholesky); fixed nodes get a zero block.  Also PCG start: x = 0, r = -g (free), z = M^-1 r, p = z
__global__ void k_gn_precondition(const RefineBufs b) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= b.n_nodes) return;
    double* H = b.Hd + 36 * n;
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
    double rz = 0.0, rr = 0.0;
    for (int i = 0; i < 6; i++) {
        b.x[6 * n + i] = 0.0;
        b.r[6 * n + i] = ok ? -b.g[6 * n + i] : 0.0;
    }
    for (int i = 0; i < 6; i++) {
        double s = 0;
        for (int k = 0; k < 6; k++) s += inv[i * 6 + k] * b.r[6 * n + k];
        b.z[6 * n + i] = s;
        b.p[6 * n + i] = s;
        rz += b.r[6 * n + i] * s;
        rr += b.r[6 * n + i] * b.r[6 * n + i];
    }
    for (int i = 0; i < 36; i++) H[i] = inv[i];  // Hd now holds M^-1
    atomicAdd(&b.scal[1], rz);
    atomicAdd(&b.scal[4], rr);
    atomicAdd(&b.scal[5], rr);
}

// q = (J^T W J + mu I) p over the free nodes, edge by edge; pq = p . q
__global__ void k_gn_matvec(const RefineBufs b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n_edges) return;
    if (b.scal[5] <= b.pcg_tol2 * b.scal[4]) return;  // converged: the remaining iterations of the fixed launch sequence are no-ops
    const Edge& ed = b.edges[i];
    const EdgeLin& L = b.lin[i];
    double u[6];
    for (int k = 0; k < 6; k++) {
        double s = 0;
        for (int c = 0; c < 6; c++) s += L.A[k * 6 + c] * b.p[6 * ed.a + c] + L.B[k * 6 + c] * b.p[6 * ed.b + c];
        u[k] = wk(b, ed.w, k) * s;
    }
    for (int c = 0; c < 6; c++) {
        double ya = 0, yb = 0;
        for (int k = 0; k < 6; k++) {
            ya += L.A[k * 6 + c] * u[k];
            yb += L.B[k * 6 + c] * u[k];
        }
        atomicAdd(&b.q[6 * ed.a + c], ya);
        atomicAdd(&b.q[6 * ed.b + c], yb);
    }
}
// one block: finish q (damping, mask fixed nodes), pq
__global__ void k_pcg_dot(const RefineBufs b) {
    __shared__ double sh[32];
    if (b.scal[5] <= b.pcg_tol2 * b.scal[4]) return;
    double s = 0.0;
    for (int i = threadIdx.x; i < 6 * b.n_nodes; i += blockDim.x) {
        const bool fixed = b.entries[i / 6].pose.var == 0.0;
        const double q = fixed ? 0.0 : b.q[i] + b.damping * b.p[i];
        b.q[i] = q;
        s += b.p[i] * q;
    }
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += sh[w];
        b.scal[2] = t;
    }
}
// one block: x += alpha p, r -= alpha q, z = M^-1 r, rz_new, beta, p = z + beta p, clear q
__global__ void k_pcg_update(const RefineBufs b) {
    __shared__ double sh[2][32];
    if (b.scal[5] <= b.pcg_tol2 * b.scal[4]) return;
    const double alpha = b.scal[2] > 0.0 ? b.scal[1] / b.scal[2] : 0.0;
    double rz = 0.0, rr = 0.0;
    for (int n = threadIdx.x; n < b.n_nodes; n += blockDim.x) {
        double r6[6];
        for (int i = 0; i < 6; i++) {
            b.x[6 * n + i] += alpha * b.p[6 * n + i];
            r6[i] = b.r[6 * n + i] - alpha * b.q[6 * n + i];
            b.r[6 * n + i] = r6[i];
            b.q[6 * n + i] = 0.0;
            rr += r6[i] * r6[i];
        }
        const double* Mi = b.Hd + 36 * n;
        for (int i = 0; i < 6; i++) {
            double s = 0;
            for (int k = 0; k < 6; k++) s += Mi[i * 6 + k] * r6[k];
            b.z[6 * n + i] = s;
            rz += r6[i] * s;
        }
    }
    for (int d = 16; d > 0; d >>= 1) {
        rz += __shfl_xor_sync(0xffffffffu, rz, d);
        rr += __shfl_xor_sync(0xffffffffu, rr, d);
    }
    if ((threadIdx.x & 31) == 0) {
        sh[0][threadIdx.x >> 5] = rz;
        sh[1][threadIdx.x >> 5] = rr;
    }
    __syncthreads();
    __shared__ double s_beta;
    if (threadIdx.x == 0) {
        double t = 0, u = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) {
            t += sh[0][w];
            u += sh[1][w];
        }
        s_beta = b.scal[1] > 0.0 ? t / b.scal[1] : 0.0;
        b.scal[1] = t;
        b.scal[5] = u;
    }
    __syncthreads();
    const double beta = s_beta;
    for (int i = threadIdx.x; i < 6 * b.n_nodes; i += blockDim.x) b.p[i] = b.z[i] + beta * b.p[i];
}
// apply the step: R <- R Exp(dtheta), t <- t + dt; reset the accumulators of the next linearisation
__global__ void k_gn_apply(const RefineBufs b, double* cost_log, int iter) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n == 0) {
        cost_log[iter] = b.scal[0];
        for (int i = 0; i < 6; i++) b.scal[i] = 0.0;
    }
    if (n >= b.n_nodes) return;
    if (b.entries[n].pose.var != 0.0) {
        double dR[9], Rn[9];
        so3_exp(b.x + 6 * n, dR);
        mat3_mul(b.entries[n].pose.R, dR, Rn);
        for (int i = 0; i < 9; i++) b.entries[n].pose.R[i] = Rn[i];
        for (int i = 0; i < 3; i++) b.entries[n].pose.t[i] += b.x[6 * n + 3 + i];
    }
    for (int i = 0; i < 6; i++) b.g[6 * n + i] = 0.0;
    for (int i = 0; i < 36; i++) b.Hd[36 * n + i] = 0.0;
}
// cost only (after the last step)
__global__ void k_gn_cost(const RefineBufs b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double c = 0.0;
    if (i < b.n_edges) {
        const Edge ed = b.edges[i];
        EdgeLin L;
        edge_terms(b.entries[ed.a], b.entries[ed.b], ed, &L);
        for (int k = 0; k < 6; k++) c += wk(b, ed.w, k) * L.e[k] * L.e[k];
    }
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0 && c != 0.0) atomicAdd(&b.scal[0], c);
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
    // device buffers (freed on every exit path below)
    Edge* d_edges = nullptr;
    EdgeLin* d_lin = nullptr;
    double* d_vec = nullptr;
    const size_t nv = (size_t)6 * n;
    const size_t vec_doubles = 6 * nv + 36 * (size_t)n + 8 + 72;  // g x r z p q, Hd, scal, cost log
    int rc = FID_OK;
    if (cudaMalloc((void**)&d_edges, sizeof(Edge) * edges.size()) != cudaSuccess || cudaMalloc((void**)&d_lin, sizeof(EdgeLin) * edges.size()) != cudaSuccess ||
        cudaMalloc((void**)&d_vec, sizeof(double) * vec_doubles) != cudaSuccess) {
        cudaGetLastError();
        rc = FID_ERR_NO_MEMORY;
    }
    std::vector<double> cost_log(72, 0.0);
    if (rc == FID_OK) {
        cudaStream_t s = m->stream;
        cudaMemcpyAsync(d_edges, edges.data(), sizeof(Edge) * edges.size(), cudaMemcpyHostToDevice, s);
        cudaMemsetAsync(d_vec, 0, sizeof(double) * vec_doubles, s);
        RefineBufs b{};
        b.edges = d_edges;
        b.lin = d_lin;
        b.n_edges = (int)edges.size();
        b.n_nodes = n;
        b.entries = m->d_entries + (size_t)instance * cap;
        b.g = d_vec;
        b.x = b.g + nv;
        b.r = b.x + nv;
        b.z = b.r + nv;
        b.p = b.z + nv;
        b.q = b.p + nv;
        b.Hd = b.q + nv;
        b.scal = b.Hd + 36 * (size_t)n;
        double* d_cost_log = b.scal + 8;
        b.lambda_t = P.translation_weight;
        b.damping = P.damping;
        b.pcg_tol2 = P.pcg_tolerance * P.pcg_tolerance;
        const int eb = (b.n_edges + 127) / 128, nb = (n + 63) / 64;
        for (int it = 0; it < P.max_iterations; it++) {
            k_gn_linearize<<<eb, 128, 0, s>>>(b);
            k_gn_precondition<<<nb, 64, 0, s>>>(b);
            for (int k = 0; k < P.pcg_iterations; k++) {
                k_gn_matvec<<<eb, 128, 0, s>>>(b);
                k_pcg_dot<<<1, 512, 0, s>>>(b);
                k_pcg_update<<<1, 512, 0, s>>>(b);
            }
            k_gn_apply<<<nb, 64, 0, s>>>(b, d_cost_log, it);
        }
        k_gn_cost<<<eb, 128, 0, s>>>(b);
        cudaMemcpyAsync(cost_log.data(), d_cost_log, sizeof(double) * 64, cudaMemcpyDeviceToHost, s);
        double final_cost = 0;
        cudaMemcpyAsync(&final_cost, b.scal, sizeof(double), cudaMemcpyDeviceToHost, s);
        const cudaError_t e = cudaStreamSynchronize(s);
        if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) {
            fprintf(stderr, "[fiducials_b200] CUDA error %s in fid_map_refine\n", cudaGetErrorString(e));
            rc = FID_ERR_CUDA;
        } else if (stats) {
            stats->initial_cost = cost_log[0];
            stats->final_cost = final_cost;
            stats->iterations = P.max_iterations;
            stats->kernel_launches = P.max_iterations * (3 + 3 * P.pcg_iterations) + 1;
        }
    }
    if (d_edges) cudaFree(d_edges);
    if (d_lin) cudaFree(d_lin);
    if (d_vec) cudaFree(d_vec);
    return rc;
}
