// tests/hostsim -- CPU unit-test harness for the FID_HD device functions of fiducials_b200/csrc.
//
// TEST INFRASTRUCTURE ONLY.  This file compiles the *same headers* the CUDA kernels are built
// from with g++, and drives them with serial loops so that the per-element logic (border walk,
// polygon approximation, grouping, bit extraction, sub-pixel refinement, PnP, map update) can be
// checked against the cv2 oracle in the GPU-less authoring container.  It is a separate shared
// object (tests/hostsim/libfid_hostsim.so), it is not linked into libfiducials_b200.so, and the
// C-ABI has no way to reach it: the product fails loudly without a CUDA device.
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <vector>

#include "../../fiducials_b200/csrc/common.cuh"
#include "../../fiducials_b200/csrc/contour_walk.cuh"
#include "../../fiducials_b200/csrc/start_prune_table.h"
#include "../../fiducials_b200/csrc/approx_quad.cuh"
#include "../../fiducials_b200/csrc/quad_group.cuh"
#include "../../fiducials_b200/csrc/identify.cuh"
#include "../../fiducials_b200/csrc/subpix.cuh"
#include "../../fiducials_b200/csrc/contour_refine.cuh"
#include "../../fiducials_b200/csrc/pnp.cuh"
#include "../../fiducials_b200/csrc/slam.cuh"
#include "../../fiducials_b200/csrc/params_host.h"
#include "../../fiducials_b200/csrc/jpeg_host.hpp"
#include "../../fiducials_b200/csrc/jpeg_math.cuh"

using namespace fid;

// Walk representation exactly as k_threshold writes it (HaloView layout) + the step tables.
struct Start {
    int x, y, is_right;
};

struct HostPlane {
    std::vector<uint32_t> halo;
    int tpr = 0, tiles_y = 0, W = 0, H = 0;
    static std::vector<uint32_t>& lut(int which) {
        static std::vector<uint32_t> p, n;
        if (p.empty()) {
            p.resize(FID_LUT_SIZE);
            n.resize(FID_LUT_SIZE);
            build_step_tables(p.data(), n.data());
        }
        return which ? n : p;
    }
    WalkCtx ctx() const { return WalkCtx{HaloView{halo.data(), tpr}, lut(0).data(), lut(1).data()}; }
};

static void pack_plane(const uint8_t* plane, int W, int H, HostPlane& hp) {
    hp.W = W;
    hp.H = H;
    hp.tpr = halo_tiles_x(W);
    hp.tiles_y = (H + FID_HALO_T - 1) / FID_HALO_T;
    hp.halo.assign(halo_plane_words(W, H), 0u);
    for (int ty = 0; ty < hp.tiles_y; ty++)
        for (int tx = 0; tx < hp.tpr; tx++)
            for (int r = 0; r < 32; r++) {
                const int Y = FID_HALO_T * ty - 1 + r;
                if (Y < 0 || Y >= H) continue;
                uint32_t w = 0;
                for (int i = 0; i < 32; i++) {
                    const int X = FID_HALO_T * tx - 1 + i;
                    if (X >= 0 && X < W && plane[(size_t)Y * W + X]) w |= 1u << i;
                }
                hp.halo[((size_t)ty * hp.tpr + tx) * 32 + r] = w;
            }
}

static int g_start_prune = 0;  // hs_set_start_prune: apply the table stage (halo_prune_starts) as the device does with FID_START_PRUNE=1

static void find_starts(const HostPlane& hp, std::vector<Start>& starts) {
    for (int ty = 0; ty < hp.tiles_y; ty++)
        for (int r = 1; r <= FID_HALO_T; r++)
            for (int tx = 0; tx < hp.tpr; tx++) {
                const uint32_t* t = hp.halo.data() + ((size_t)ty * hp.tpr + tx) * 32 + r;
                uint32_t L = 0, R = 0;
                if (t[0]) halo_row_starts(t[-1], t[0], t[1], &L, &R);
                if (g_start_prune && (L | R)) halo_prune_starts(t[-1], t[0], t[1], &L, &R, &kStartPruneTable[0][0]);
                const int y = FID_HALO_T * ty - 1 + r;
                for (int i = 1; i <= FID_HALO_T; i++) {
                    if ((L >> i) & 1) starts.push_back({FID_HALO_T * tx - 1 + i, y, 0});
                    if ((R >> i) & 1) starts.push_back({FID_HALO_T * tx - 1 + i, y, 1});
                }
            }
}

struct HostWalk {  // thin adapter kept for the harness below
    const HostPlane* hp = nullptr;
    void build(const HostPlane& p) { hp = &p; }
    WalkCtx ctx() const { return hp->ctx(); }
};

extern "C" {

// Walk + emission exactly as the GPU does them: `uni` one-directional steps (round 0), then the
// bidirectional walk in passes of `pass` steps with checkpoints every `ck_step` steps, then one
// trace_segment per segment.  Returns WALK_* and fills pts (n points) for a canonical start.
static int walk_and_emit(const WalkCtx& ctx, const Start& s, int max_len, int uni, int pass, int ck_step, std::vector<uint32_t>& pts, int* n_out) {
    WalkState st;
    if (walk_init(ctx, s.x, s.y, s.is_right, &st) != WALK_CONTINUE) return WALK_ABORT;
    int r = WALK_CONTINUE;
    if (uni > 0) r = s.is_right ? walk_uni_fast<true>(ctx, s.x, s.y, max_len, uni, &st) : walk_uni_fast<false>(ctx, s.x, s.y, max_len, uni, &st);
    WalkState2 s2;
    if (s.is_right) walk_split<true>(s.x, s.y, st, &s2); else walk_split<false>(s.x, s.y, st, &s2);
    WalkCkpt ck;
    ck.count[0] = ck.count[1] = 0;
    int last_f = 0, last_b = 0;
    while (r == WALK_CONTINUE) {
        r = s.is_right ? walk_bidir_fast<true>(ctx, s.x, s.y, max_len, pass, &s2) : walk_bidir_fast<false>(ctx, s.x, s.y, max_len, pass, &s2);
        if (r == WALK_CONTINUE && ck_step > 0) walk_checkpoint(s2, &ck, &last_f, &last_b, ck_step);
    }
    if (r != WALK_CANONICAL) return r;
    const int n = s2.n;
    *n_out = n;
    pts.assign((size_t)n, 0xFFFFFFFFu);
    std::vector<SegRec> segs((size_t)segment_count(&ck));
    make_segments(ctx, s.x, s.y, s.is_right, n, s2.nf, &ck, 0u, 0u, [&](int k, const SegRec& sr) { segs[(size_t)k] = sr; });
    for (const SegRec& sr : segs) trace_segment(ctx, sr, pts.data());
    return r;
}

// All contours of a {0,!=0} plane with min_len <= n <= max_len, in cv2.findContours list order.
// out_pts: (x,y) int16 pairs, out_len: per-contour lengths.  Returns the number of contours, or
// -1 if a buffer is too small.  mode 0: one-directional walk + trace_forward; mode 1: the GPU's
// round structure (8 one-directional steps, bidirectional passes of 16, checkpoints, segments);
// mode 2: same with tiny passes/checkpoint spacing to exercise every code path on small planes.
int hs_find_contours_mode(const uint8_t* plane, int W, int H, int min_len, int max_len, int16_t* out_pts, int64_t max_pts, int32_t* out_len, int max_contours,
                          int64_t* n_starts_out, int mode) {
    HostPlane mask;
    pack_plane(plane, W, H, mask);
    std::vector<Start> starts;
    find_starts(mask, starts);
    HostWalk hw;
    hw.build(mask);
    if (n_starts_out) *n_starts_out = (int64_t)starts.size();
    struct Chain {
        int64_t key;
        int x, y, is_right, n;
        std::vector<uint32_t> pts;
    };
    std::vector<Chain> chains;
    for (const Start& s : starts) {
        int n = 0;
        std::vector<uint32_t> pts;
        int st;
        if (mode == 0)
            st = walk_start(hw.ctx(), s.x, s.y, s.is_right, max_len, &n);
        else
            st = walk_and_emit(hw.ctx(), s, max_len, mode == 1 ? 8 : 1, mode == 1 ? 16 : 2, mode == 1 ? FID_CKPT_STEP : 3, pts, &n);
        if (st == WALK_CANONICAL && n >= min_len && n <= max_len) chains.push_back({((int64_t)s.y * W + s.x) * 2 + s.is_right, s.x, s.y, s.is_right, n, std::move(pts)});
    }
    if (min_len <= 1) {  // isolated pixels are 1-point outer contours
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++)
                if (plane[(size_t)y * W + x] && (mask.ctx().plane.idx9(x, y) & ~0x10u) == 0u)
                    chains.push_back({((int64_t)y * W + x) * 2, x, y, 0, 1, std::vector<uint32_t>{(uint32_t)x | ((uint32_t)y << 16)}});
    }
    std::sort(chains.begin(), chains.end(), [](const Chain& a, const Chain& b) { return a.key > b.key; });  // reverse discovery order
    if ((int)chains.size() > max_contours) return -1;
    int64_t off = 0;
    for (size_t i = 0; i < chains.size(); i++) {
        const Chain& c = chains[i];
        if (off + c.n > max_pts) return -1;
        if (mode == 0 && c.n > 1) {
            std::vector<Pt16> tmp((size_t)c.n + 4);
            trace_forward(hw.ctx(), c.x, c.y, c.is_right, c.n, tmp.data());
            memcpy(reinterpret_cast<Pt16*>(out_pts) + off, tmp.data(), sizeof(Pt16) * c.n);
        } else {
            memcpy(reinterpret_cast<Pt16*>(out_pts) + off, c.pts.data(), sizeof(Pt16) * c.n);
        }
        out_len[i] = c.n;
        off += c.n;
    }
    return (int)chains.size();
}

int hs_find_contours(const uint8_t* plane, int W, int H, int min_len, int max_len, int16_t* out_pts, int64_t max_pts, int32_t* out_len, int max_contours,
                     int64_t* n_starts_out) {
    return hs_find_contours_mode(plane, W, H, min_len, max_len, out_pts, max_pts, out_len, max_contours, n_starts_out, 1);
}

// Walk statistics of one plane: out[0] starts, [1] total reverse-walk steps, [2] walks > 64 steps,
// [3] walks > 1024 steps, [4] walks that hit max_len, [5] canonical walks, [6] steps spent in
// canonical walks, [7] steps spent in too-long walks.
void hs_walk_stats(const uint8_t* plane, int W, int H, int max_len, int64_t* out) {
    HostPlane mask;
    pack_plane(plane, W, H, mask);
    std::vector<Start> starts;
    find_starts(mask, starts);
    HostWalk hw;
    hw.build(mask);
    for (int i = 0; i < 8; i++) out[i] = 0;
    out[0] = (int64_t)starts.size();
    for (const Start& st : starts) {
        int n = 0, steps = 0;
        const int result = walk_start(hw.ctx(), st.x, st.y, st.is_right, max_len, &n, &steps);
        out[1] += steps;
        if (steps > 64) out[2]++;
        if (steps > 1024) out[3]++;
        if (result == WALK_TOO_LONG) {
            out[4]++;
            out[7] += steps;
        }
        if (result == WALK_CANONICAL) {
            out[5]++;
            out[6] += steps;
        }
    }
}

// Detailed walk log of one plane: for every start with more than min_steps steps:
// rows of (x, y, is_right, result, steps).  Returns the number of rows.
int hs_walk_log(const uint8_t* plane, int W, int H, int max_len, int min_steps, int32_t* out, int max_rows) {
    HostPlane mask;
    pack_plane(plane, W, H, mask);
    std::vector<Start> starts;
    find_starts(mask, starts);
    HostWalk hw;
    hw.build(mask);
    int rows = 0;
    for (const Start& st : starts) {
        int n = 0, steps = 0;
        const int result = walk_start(hw.ctx(), st.x, st.y, st.is_right, max_len, &n, &steps);
        if (steps > min_steps && rows < max_rows) {
            int32_t* o = out + 5 * rows++;
            o[0] = st.x; o[1] = st.y; o[2] = st.is_right; o[3] = result; o[4] = steps;
        }
    }
    return rows;
}

// Simulation of the round-based walk with optional kill marks: a walk that passes the crack of
// another (raster-larger) start clears that start's alive bit; dead starts are dropped when their
// next round begins.  out[r] = steps taken in round r, out[8+r] = walks entering round r.
void hs_walk_sim(const uint8_t* plane, int W, int H, int max_len, const int* budgets, int n_rounds, int use_kills, int64_t* out) {
    HostPlane mask;
    pack_plane(plane, W, H, mask);
    std::vector<Start> starts;
    find_starts(mask, starts);
    HostWalk hw;
    hw.build(mask);
    std::vector<uint8_t> aliveL((size_t)W * H, 0), aliveR((size_t)W * H, 0);
    for (const Start& s : starts) (s.is_right ? aliveR : aliveL)[(size_t)s.y * W + s.x] = 1;
    struct Live { Start s; WalkState st; };
    std::vector<Live> cur, nxt;
    for (int i = 0; i < 16; i++) out[i] = 0;
    const WalkCtx ctx = hw.ctx();
    for (const Start& s : starts) {
        Live l;
        l.s = s;
        if (walk_init(ctx, s.x, s.y, s.is_right, &l.st) == WALK_CONTINUE) cur.push_back(l);
    }
    for (int r = 0; r < n_rounds; r++) {
        out[8 + r] = (int64_t)cur.size();
        nxt.clear();
        for (Live& l : cur) {
            if (use_kills && !(l.s.is_right ? aliveR : aliveL)[(size_t)l.s.y * W + l.s.x]) continue;
            const int before = l.st.n;
            const int x0 = l.s.x, y0 = l.s.y, isr = l.s.is_right;
            auto visit = [&](int x, int y, bool exL, bool exR) {
                if (!use_kills) return;
                if (exL && !(x == x0 && y == y0 && !isr)) aliveL[(size_t)y * W + x] = 0;
                if (exR && !(x == x0 && y == y0 && isr)) aliveR[(size_t)y * W + x] = 0;
            };
            int res;
            if (l.s.is_right)
                res = walk_resume_dir<true>(ctx, x0, y0, max_len, budgets[r], &l.st, visit);
            else
                res = walk_resume_dir<false>(ctx, x0, y0, max_len, budgets[r], &l.st, visit);
            out[r] += l.st.n - before;
            if (res == WALK_CONTINUE) nxt.push_back(l);
        }
        cur.swap(nxt);
    }
}

// Round-based walk with bidirectional walkers from round `bidir_from` on (contour_walk.cuh,
// walk_resume_bidir).  out[r] = steps in round r, out[8+r] = walks entering round r, out[16] = canonical
// walks found, out[17] = sum of their lengths, out[18] = too-long walks.
void hs_walk_sim2(const uint8_t* plane, int W, int H, int max_len, const int* budgets, int n_rounds, int bidir_from, int64_t* out) {
    HostPlane mask;
    pack_plane(plane, W, H, mask);
    std::vector<Start> starts;
    find_starts(mask, starts);
    HostWalk hw;
    hw.build(mask);
    struct Live { Start s; WalkState st; WalkState2 s2; bool split; };
    std::vector<Live> cur, nxt;
    for (int i = 0; i < 24; i++) out[i] = 0;
    const WalkCtx ctx = hw.ctx();
    for (const Start& s : starts) {
        Live l;
        l.s = s;
        l.split = false;
        if (walk_init(ctx, s.x, s.y, s.is_right, &l.st) == WALK_CONTINUE) cur.push_back(l);
    }
    for (int r = 0; r < n_rounds; r++) {
        out[8 + r] = (int64_t)cur.size();
        nxt.clear();
        for (Live& l : cur) {
            const int x0 = l.s.x, y0 = l.s.y;
            int res, before, after;
            if (r >= bidir_from) {
                if (!l.split) {
                    if (l.s.is_right) walk_split<true>(x0, y0, l.st, &l.s2); else walk_split<false>(x0, y0, l.st, &l.s2);
                    l.split = true;
                }
                before = l.s2.n;
                res = l.s.is_right ? walk_resume_bidir<true>(ctx, x0, y0, max_len, budgets[r], &l.s2) : walk_resume_bidir<false>(ctx, x0, y0, max_len, budgets[r], &l.s2);
                after = l.s2.n;
            } else {
                before = l.st.n;
                res = l.s.is_right ? walk_resume_dir<true>(ctx, x0, y0, max_len, budgets[r], &l.st) : walk_resume_dir<false>(ctx, x0, y0, max_len, budgets[r], &l.st);
                after = l.st.n;
            }
            out[r] += after - before;
            if (res == WALK_CONTINUE) nxt.push_back(l);
            if (res == WALK_CANONICAL) { out[16]++; out[17] += after; }
            if (res == WALK_TOO_LONG) out[18]++;
        }
        cur.swap(nxt);
    }
}

// Every start crack of the plane walked one-directionally and bidirectionally (after `uni_steps` steps of
// the one-directional walk, as the GPU does between round 0 and round 1): number of starts whose
// (result, contour length) differ.  out[0] = starts, out[1] = canonical.
int hs_walk_bidir_check(const uint8_t* plane, int W, int H, int max_len, int uni_steps, int chunk, int fast, int64_t* out) {
    HostPlane mask;
    pack_plane(plane, W, H, mask);
    std::vector<Start> starts;
    find_starts(mask, starts);
    HostWalk hw;
    hw.build(mask);
    const WalkCtx ctx = hw.ctx();
    int bad = 0;
    out[0] = (int64_t)starts.size();
    out[1] = 0;
    for (const Start& s : starts) {
        int n1 = 0;
        const int r1 = walk_start(ctx, s.x, s.y, s.is_right, max_len, &n1);
        WalkState st;
        int r2 = walk_init(ctx, s.x, s.y, s.is_right, &st), n2 = 0;
        if (r2 == WALK_CONTINUE) {
            if (uni_steps <= 0) r2 = WALK_CONTINUE;
            else if (!fast) r2 = walk_resume(ctx, s.x, s.y, s.is_right, max_len, uni_steps, &st);
            else r2 = s.is_right ? walk_uni_fast<true>(ctx, s.x, s.y, max_len, uni_steps, &st) : walk_uni_fast<false>(ctx, s.x, s.y, max_len, uni_steps, &st);
            n2 = st.n;
            if (r2 == WALK_CONTINUE) {
                WalkState2 s2;
                if (s.is_right) walk_split<true>(s.x, s.y, st, &s2); else walk_split<false>(s.x, s.y, st, &s2);
                do {
                    if (!fast) r2 = s.is_right ? walk_resume_bidir<true>(ctx, s.x, s.y, max_len, chunk, &s2) : walk_resume_bidir<false>(ctx, s.x, s.y, max_len, chunk, &s2);
                    else r2 = s.is_right ? walk_bidir_fast<true>(ctx, s.x, s.y, max_len, chunk, &s2) : walk_bidir_fast<false>(ctx, s.x, s.y, max_len, chunk, &s2);
                } while (r2 == WALK_CONTINUE);
                n2 = s2.n;
            }
        }
        if (r1 == WALK_CANONICAL && n1 <= max_len) out[1]++;
        // ABORT and TOO_LONG both mean "no contour": which one is hit first depends on the direction walked
        // (and a lap that closes a step or a pass after max_len is "no contour" as well: k_walk drops it)
        const bool c1 = r1 == WALK_CANONICAL && n1 <= max_len, c2 = r2 == WALK_CANONICAL && n2 <= max_len;
        if (c1 != c2 || (c1 && n1 != n2)) bad++;
    }
    return bad;
}

// approxPolyDP (closed) of one contour; returns vertex count (-1 = more than 8 before clean-up).
int hs_approx_poly(const int16_t* pts, int n, double eps, int16_t* out) {
    SerialReducer red;
    return approx_poly_closed(red, reinterpret_cast<const Pt16*>(pts), n, eps, reinterpret_cast<Pt16*>(out));
}

int hs_is_convex(const int16_t* pts, int n) { return is_convex_int(reinterpret_cast<const Pt16*>(pts), n) ? 1 : 0; }


// Raw quad candidates of all scales from precomputed threshold planes (n_scales x H x W, {0,!=0}),
// in OpenCV's concatenation order.  quads: n x 8 int32 (x0,y0,..), scale[n], clen[n].
static void raw_candidates(const uint8_t* planes, int W, int H, const DevParams& P, std::vector<RawQuad>& out, std::vector<Pt16>* keep_pts = nullptr) {
    const int mx = W > H ? W : H;
    const int min_len = (int)(P.min_perimeter_rate * mx), max_len = (int)(P.max_perimeter_rate * mx);
    for (int s = 0; s < P.n_scales; s++) {
        HostPlane mask;
        pack_plane(planes + (size_t)s * W * H, W, H, mask);
        std::vector<Start> starts;
        find_starts(mask, starts);
        HostWalk hw;
        hw.build(mask);
        std::vector<RawQuad> found;
        std::vector<Pt16> pts;
        for (const Start& st : starts) {
            int n = 0;
            if (walk_start(hw.ctx(), st.x, st.y, st.is_right, max_len, &n) != WALK_CANONICAL) continue;
            if (n < min_len || n > max_len) continue;
            pts.resize((size_t)n + 4);
            trace_forward(hw.ctx(), st.x, st.y, st.is_right, n, pts.data());
            Pt16 q[FID_APPROX_MAX_V];
            SerialReducer red;
            if (approx_poly_closed(red, pts.data(), n, (double)n * P.poly_accuracy_rate, q) != 4) continue;
            if (!quad_passes_filters(q, n, W, H, P.min_corner_dist_rate)) continue;
            RawQuad r;
            for (int k = 0; k < 4; k++) {
                r.x[k] = q[k].x;
                r.y[k] = q[k].y;
            }
            r.n_contour = n;
            r.pts_off = 0;
            if (keep_pts) {
                r.pts_off = (uint32_t)keep_pts->size();
                keep_pts->insert(keep_pts->end(), pts.begin(), pts.begin() + n);
            }
            r.order_hi = (uint32_t)s;
            r.order_lo = 0xFFFFFFFFu - (uint32_t)(((uint32_t)st.y * (uint32_t)W + (uint32_t)st.x) * 2u + (uint32_t)st.is_right);
            found.push_back(r);
        }
        std::sort(found.begin(), found.end(), [](const RawQuad& a, const RawQuad& b) { return a.order_lo < b.order_lo; });
        out.insert(out.end(), found.begin(), found.end());
    }
}

int hs_candidates(const uint8_t* planes, int W, int H, int dict_id, int32_t* quads, int32_t* scale, int32_t* clen, int max_out) {
    fid_params fp;
    default_params(&fp);
    fp.dictionary = dict_id;
    DevParams P;
    if (make_dev_params(fp, &P) != FID_OK) return -2;
    std::vector<RawQuad> raw;
    raw_candidates(planes, W, H, P, raw);
    if ((int)raw.size() > max_out) return -1;
    for (size_t i = 0; i < raw.size(); i++) {
        for (int k = 0; k < 4; k++) {
            quads[i * 8 + 2 * k] = raw[i].x[k];
            quads[i * 8 + 2 * k + 1] = raw[i].y[k];
        }
        scale[i] = (int)raw[i].order_hi;
        clen[i] = raw[i].n_contour;
    }
    return (int)raw.size();
}

// Full detect from gray + threshold planes: ids/corners in OpenCV order.  refine: apply cornerSubPix.
int hs_detect(const uint8_t* gray, const uint8_t* planes, int W, int H, int dict_id, int refine, int32_t* ids, float* corners, int max_out, int32_t* stats) {
    fid_params fp;
    default_params(&fp);
    fp.dictionary = dict_id;
    DevParams P;
    if (make_dev_params(fp, &P) != FID_OK) return -2;
    std::vector<RawQuad> raw;
    std::vector<Pt16> contour_pts;
    raw_candidates(planes, W, H, P, raw, &contour_pts);
    const int n = (int)raw.size();
    // stable sort by descending float perimeter
    std::vector<QuadF> q(n);
    std::vector<float> per(n);
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) {
        q[i] = quad_clockwise(raw[i]);
        per[i] = quad_perimeter(q[i]);
        order[i] = i;
    }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return per[a] > per[b]; });
    std::vector<QuadF> sq(n);
    std::vector<float> sper(n);
    for (int i = 0; i < n; i++) {
        sq[i] = q[order[i]];
        sper[i] = per[order[i]];
    }
    std::vector<uint8_t> selected(n);
    std::vector<int> gid(n), gmem(2 * (size_t)n + 2), nxt(n), ghead(n), gtail(n), ccount(n), cidx(n), coff(n + 1);
    std::vector<uint32_t> grouped_bits((size_t)(n + 31) / 32 + 1);
    const float rate = (float)P.min_marker_dist_rate;
    struct CloseWordHost {
        const std::vector<QuadF>* sq;
        const std::vector<float>* sper;
        int n;
        float rate;
        uint32_t operator()(int i, int w) const {
            uint32_t bits = 0;
            for (int b = 0; b < 32; b++) {
                const int j = 32 * w + b;
                if (j > i && j < n && quad_avg_distance((*sq)[i], (*sq)[j]) < (*sper)[j] * rate) bits |= 1u << b;
            }
            return bits;
        }
        bool row_any(int) const { return true; }
    } close_word{&sq, &sper, n, rate};
    group_candidates(SerialLanes(), n, sq.data(), P.marker_size, P.marker_border_bits, (float)P.min_group_dist, close_word, selected.data(), gid.data(), gmem.data(), nxt.data(),
                     ghead.data(), gtail.data(), ccount.data(), cidx.data(), coff.data(), grouped_bits.data());
    std::vector<unsigned long long> dict;
    pack_dictionary(P, &dict);
    SerialLanes L;
    std::vector<uint8_t> img(64 * 64);
    int hist[256];
    int n_out = 0, n_sel = 0;
    float mask[121];
    std::vector<float> patch(13 * 13);
    for (int i = 0; i < n; i++) {
        if (!selected[i]) continue;
        if (quad_near_border(sq[i], W, H, P.min_dist_to_border)) continue;
        n_sel++;
        QuadF use = sq[i];
        int use_sorted = i;
        IdentifyResult r = identify_candidate(L, GrayPlane{gray, (size_t)W}, W, H, use, P, dict.data(), img.data(), hist);
        if (r.id < 0) {
            for (int k = 0; k < ccount[i]; k++) {
                const QuadF& alt = sq[cidx[coff[i] + k]];
                r = identify_candidate(L, GrayPlane{gray, (size_t)W}, W, H, alt, P, dict.data(), img.data(), hist);
                if (r.id >= 0) {
                    use = alt;
                    use_sorted = cidx[coff[i] + k];
                    break;
                }
            }
        }
        if (r.id < 0) continue;
        if (n_out >= max_out) return -1;
        // correctCornerPosition: std::rotate(begin, begin + 4 - rotation, end)
        float cx[4], cy[4];
        for (int k = 0; k < 4; k++) {
            cx[k] = use.x[(k + 4 - r.rotation) & 3];
            cy[k] = use.y[(k + 4 - r.rotation) & 3];
        }
        if (refine == 2) {  // CORNER_REFINE_CONTOUR
            const RawQuad& rw = raw[order[use_sorted]];
            refine_candidate_lines_serial(contour_pts.data() + rw.pts_off, rw.n_contour, cx, cy);
        } else if (refine && P.corner_refine) {
            QuadF rq;
            for (int k = 0; k < 4; k++) {
                rq.x[k] = cx[k];
                rq.y[k] = cy[k];
            }
            const float module = quad_module_size(rq, P.marker_size, P.marker_border_bits);
            int win = (int)nearbyintf((float)P.rel_refine_win * module);
            win = win < 1 ? 1 : win;
            win = win < P.refine_win ? win : P.refine_win;
            subpix_mask(win, mask);
            for (int k = 0; k < 4; k++)
                corner_subpix(GrayPlane{gray, (size_t)W}, W, H, &cx[k], &cy[k], win, mask, P.refine_max_iter, P.refine_min_acc * P.refine_min_acc, patch.data());
        }
        ids[n_out] = r.id;
        for (int k = 0; k < 4; k++) {
            corners[n_out * 8 + 2 * k] = cx[k];
            corners[n_out * 8 + 2 * k + 1] = cy[k];
        }
        n_out++;
    }
    if (stats) {
        stats[0] = n;
        stats[1] = n_sel;
    }
    return n_out;
}

// cornerSubPix of n points (x,y interleaved, in place).
void hs_corner_subpix(const uint8_t* gray, int W, int H, float* pts, int n, int win, int max_iters, double eps) {
    float mask[121];
    float patch[13 * 13];
    subpix_mask(win, mask);
    for (int i = 0; i < n; i++) corner_subpix(GrayPlane{gray, (size_t)W}, W, H, &pts[2 * i], &pts[2 * i + 1], win, mask, max_iters, eps * eps, patch);
}

// Pose of n markers; out: n x 19 doubles (rvec3 tvec3 image_error object_error area quat4 iters pad..)
void hs_pose(int n, const float* corners, const double* K, const double* D, const float* lens, double default_len, double* out) {
    Camera cam = {K[0], K[4], K[2], K[5], D[0], D[1], D[2], D[3], D[4]};
    for (int i = 0; i < n; i++) {
        PoseOut po;
        solve_marker_pose(corners + 8 * i, cam, lens[i], default_len, &po);
        double* o = out + 16 * i;
        for (int k = 0; k < 3; k++) {
            o[k] = po.rvec[k];
            o[3 + k] = po.tvec[k];
        }
        o[6] = po.image_error;
        o[7] = po.object_error;
        o[8] = po.area;
        for (int k = 0; k < 4; k++) o[9 + k] = po.quat[k];
        o[13] = po.lm_iters;
    }
}

// --- map ---------------------------------------------------------------------------------------
struct HsMap {
    MapState st;
    std::vector<MapEntry> e;
    std::vector<uint32_t> links;
};

void* hs_map_create(int capacity, int read_only) {
    HsMap* m = new HsMap();
    memset(&m->st, 0, sizeof(m->st));
    m->st.capacity = capacity;
    m->st.origin_fid = -1;
    m->st.fiducial_to_add = -1;
    m->st.read_only = read_only;
    m->e.resize(capacity);
    m->links.assign((size_t)capacity * ((capacity + 31) / 32), 0u);
    return m;
}
void hs_map_destroy(void* h) { delete (HsMap*)h; }
// obs: n x (id as double, t3, q4, object_error, area) = 10 doubles; tf: 7 doubles (t3 q4) or null
int hs_map_update(void* h, int n, const double* obs, const double* baseCam, const double* camBase, double* robot /* valid,n,t3,q4,var = 10 */) {
    HsMap* m = (HsMap*)h;
    std::vector<Obs> o(n);
    for (int i = 0; i < n; i++) {
        o[i].id = (int)obs[10 * i];
        for (int k = 0; k < 3; k++) o[i].t[k] = obs[10 * i + 1 + k];
        for (int k = 0; k < 4; k++) o[i].q[k] = obs[10 * i + 4 + k];
        o[i].object_error = obs[10 * i + 8];
        o[i].area = obs[10 * i + 9];
    }
    Twv bc, cb;
    if (baseCam) {
        q_to_m(baseCam + 3, bc.R);
        bc.t[0] = baseCam[0];
        bc.t[1] = baseCam[1];
        bc.t[2] = baseCam[2];
        bc.var = 0;
    }
    if (camBase) {
        q_to_m(camBase + 3, cb.R);
        cb.t[0] = camBase[0];
        cb.t[1] = camBase[1];
        cb.t[2] = camBase[2];
        cb.var = 0;
    }
    RobotPose rp;
    std::vector<double> var_scratch((size_t)n + 1);
    std::vector<int> slot_scratch((size_t)n + 1);
    map_update(m->st, m->e.data(), m->links.data(), o.data(), n, baseCam ? &bc : nullptr, camBase ? &cb : nullptr, 1e9, 0, 0.01, &rp, nullptr, var_scratch.data(), slot_scratch.data());
    robot[0] = rp.valid;
    robot[1] = rp.n_estimates;
    for (int k = 0; k < 3; k++) robot[2 + k] = rp.t[k];
    for (int k = 0; k < 4; k++) robot[5 + k] = rp.q[k];
    robot[9] = rp.var;
    return m->st.n;
}
void hs_map_load(void* h, int id, double x, double y, double z, double r_deg, double p_deg, double y_deg, double var, int num_obs) {
    HsMap* m = (HsMap*)h;
    const double d2r = 3.14159265358979323846 / 180.0;
    const double hr = r_deg * d2r * 0.5, hp = p_deg * d2r * 0.5, hy = y_deg * d2r * 0.5;
    const double cy = cos(hy), sy = sin(hy), cp = cos(hp), sp = sin(hp), cr = cos(hr), sr = sin(hr);
    const double q[4] = {sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy};
    MapEntry& e = m->e[m->st.n++];
    e.id = id;
    e.num_obs = num_obs;
    q_to_m(q, e.pose.R);
    e.pose.t[0] = x;
    e.pose.t[1] = y;
    e.pose.t[2] = z;
    e.pose.var = var;
}
// entries: n x (id, x,y,z, rx,ry,rz, var, numObs) = 9 doubles, ascending id
int hs_map_entries(void* h, double* out) {
    HsMap* m = (HsMap*)h;
    std::vector<int> idx(m->st.n);
    for (int i = 0; i < m->st.n; i++) idx[i] = i;
    std::sort(idx.begin(), idx.end(), [&](int a, int b) { return m->e[a].id < m->e[b].id; });
    for (int k = 0; k < m->st.n; k++) {
        const MapEntry& e = m->e[idx[k]];
        double r, p, y;
        get_rpy(e.pose.R, &r, &p, &y);
        double* o = out + 9 * k;
        o[0] = e.id;
        o[1] = e.pose.t[0];
        o[2] = e.pose.t[1];
        o[3] = e.pose.t[2];
        o[4] = r;
        o[5] = p;
        o[6] = y;
        o[7] = e.pose.var;
        o[8] = e.num_obs;
    }
    return m->st.n;
}


// JPEG ingest: host entropy decoder (jpeg_host.hpp) + the device arithmetic (jpeg_math.cuh) in serial loops.
// Returns 0 and fills bgr[H][W][3] / dims, or the decoder's negative status.
int hs_jpeg_decode(const uint8_t* data, long long size, uint8_t* bgr, int max_w, int max_h, int* out_w, int* out_h, long long* n_vals) {
    using namespace fidjpeg;
    FrameInfo fi;
    const size_t max_blk = (size_t)((max_w + 15) / 8 + 2) * ((max_h + 15) / 8 + 2) * 3;
    std::vector<uint64_t> mask(max_blk);
    std::vector<uint32_t> off(max_blk);
    std::vector<int16_t> vals(max_blk * 64);
    size_t nv = 0;
    const int rc = decode_image(data, (size_t)size, &fi, mask.data(), off.data(), vals.data(), vals.size(), max_blk, &nv);
    if (rc != JPEG_OK) return rc;
    if (fi.W > max_w || fi.H > max_h) return JPEG_CAPACITY;
    *out_w = fi.W;
    *out_h = fi.H;
    *n_vals = (long long)nv;
    std::vector<std::vector<uint8_t>> plane(fi.ncomp);
    for (int c = 0; c < fi.ncomp; c++) {
        const int pitch = fi.bw[c] * 8;
        plane[c].assign((size_t)pitch * fi.bh[c] * 8, 0);
        for (int by = 0; by < fi.bh[c]; by++)
            for (int bx = 0; bx < fi.bw[c]; bx++) {
                const int b = fi.blk_base[c] + by * fi.bw[c] + bx;
                int coef[64];
                for (int i = 0; i < 64; i++) coef[i] = 0;
                uint64_t m = mask[b];
                const int16_t* v = vals.data() + off[b];
                for (int k = 0; k < 64; k++)
                    if ((m >> k) & 1) coef[kZigzag[k]] = (int)(*v++) * (int)fi.q[c][k];
                uint8_t out[64];
                jpeg_idct_block(coef, out);
                for (int r = 0; r < 8; r++) memcpy(&plane[c][(size_t)(by * 8 + r) * pitch + bx * 8], out + r * 8, 8);
            }
    }
    const int mode = fi.ncomp == 1 ? 0 : (fi.hmax == 1 ? 0 : (fi.vmax == 1 ? 1 : 2));
    for (int y = 0; y < fi.H; y++)
        for (int x = 0; x < fi.W; x++) {
            uint8_t* o = bgr + ((size_t)y * fi.W + x) * 3;
            const int Y = plane[0][(size_t)y * fi.bw[0] * 8 + x];
            if (fi.ncomp == 1) {
                o[0] = o[1] = o[2] = (uint8_t)Y;
            } else {
                const int cb = jpeg_chroma_at(plane[1].data(), fi.bw[1] * 8, fi.cw[1], fi.ch[1], mode, x, y);
                const int cr = jpeg_chroma_at(plane[2].data(), fi.bw[2] * 8, fi.cw[2], fi.ch[2], mode, x, y);
                jpeg_ycc_to_bgr(Y, cb, cr, o);
            }
        }
    return 0;
}


// entropy stage alone, `reps` times (tools/time_jpeg_entropy.py): returns the number of values of the last pass or a negative status
long long hs_jpeg_entropy(const uint8_t* data, long long size, int reps) {
    using namespace fidjpeg;
    FrameInfo fi;
    const size_t max_blk = (size_t)(4096 / 8 + 2) * (4096 / 8 + 2) * 3;
    static std::vector<uint64_t> mask(max_blk);
    static std::vector<uint32_t> off(max_blk);
    static std::vector<int16_t> vals(max_blk * 64);
    size_t nv = 0;
    for (int r = 0; r < reps; r++) {
        const int rc = decode_image(data, (size_t)size, &fi, mask.data(), off.data(), vals.data(), vals.size(), max_blk, &nv);
        if (rc != JPEG_OK) return rc;
    }
    return (long long)nv;
}


void hs_set_start_prune(int on) { g_start_prune = on; }
void hs_committed_prune_table(uint32_t* out) { memcpy(out, kStartPruneTable, sizeof(kStartPruneTable)); }

// ---- start-prune table (tools/gen_prune_table.py) ------------------------------------------------------------
// For every 3-row x `cols`-column neighbourhood of a start crack: does the walk of that start abort within K steps for EVERY
// completion of the pixels outside the neighbourhood?  Pixels are assigned lazily: the walk only reads the 3x3 around the
// pixels it visits, so only unknown pixels inside those read sets are branched on.
namespace {
struct PruneSim {
    static const int N = 40, CX = 20, CY = 20;
    int8_t val[N][N];  // -1 unknown, 0, 1
    int K, is_right;
    long nodes = 0;
    bool all_abort() {
        nodes++;
        std::vector<uint8_t> img((size_t)N * N);
        for (int y = 0; y < N; y++)
            for (int x = 0; x < N; x++) img[(size_t)y * N + x] = val[y][x] == 1 ? 1 : 0;
        HostPlane hp;
        pack_plane(img.data(), N, N, hp);
        const WalkCtx c = hp.ctx();
        int vx[16], vy[16], nv = 0;
        vx[nv] = CX;
        vy[nv++] = CY;
        WalkState st;
        int r = walk_init(c, CX, CY, is_right, &st);
        int steps = 0;
        while (r == WALK_CONTINUE && steps < K) {
            r = is_right ? walk_uni_fast<true>(c, CX, CY, 1 << 20, 1, &st) : walk_uni_fast<false>(c, CX, CY, 1 << 20, 1, &st);
            steps++;
            vx[nv] = st.x;
            vy[nv++] = st.y;
        }
        // unknown pixels inside the read sets of this run
        int ux[64], uy[64], nu = 0;
        for (int k = 0; k < nv; k++)
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    const int x = vx[k] + dx, y = vy[k] + dy;
                    if (val[y][x] != -1) continue;
                    bool seen = false;
                    for (int q = 0; q < nu; q++) seen = seen || (ux[q] == x && uy[q] == y);
                    if (!seen && nu < 64) {
                        ux[nu] = x;
                        uy[nu++] = y;
                    }
                }
        if (nu == 0) return r == WALK_ABORT;
        if (nu > 16) return false;  // give up on patterns that fan out this far
        for (uint32_t a = 0; a < (1u << nu); a++) {
            for (int q = 0; q < nu; q++) val[uy[q]][ux[q]] = (int8_t)((a >> q) & 1);
            const bool ok = all_abort();
            if (!ok) {
                for (int q = 0; q < nu; q++) val[uy[q]][ux[q]] = -1;
                return false;
            }
        }
        for (int q = 0; q < nu; q++) val[uy[q]][ux[q]] = -1;
        return true;
    }
};
}  // namespace

// out[pattern] = 1 when a start with this neighbourhood can be dropped.  pattern bit (row * cols + col), row 0 = the row above the
// start, col cols/2 = the start's column.  Returns the number of prunable patterns.
long hs_prune_table(int cols, int K, int is_right, uint8_t* out) {
    const int half = cols / 2, nbits = 3 * cols;
    long n_prunable = 0;
    PruneSim sim;
    sim.K = K;
    sim.is_right = is_right;
    for (uint32_t pat = 0; pat < (1u << nbits); pat++) {
        out[pat] = 0;
        const int centre = (int)((pat >> (cols + half)) & 1u);
        const int side = (int)((pat >> (cols + half + (is_right ? 1 : -1))) & 1u);
        if (!centre || side) continue;  // not a left / right crack
        for (int y = 0; y < PruneSim::N; y++)
            for (int x = 0; x < PruneSim::N; x++) sim.val[y][x] = -1;
        for (int r = 0; r < 3; r++)
            for (int cc = 0; cc < cols; cc++) sim.val[PruneSim::CY - 1 + r][PruneSim::CX - half + cc] = (int8_t)((pat >> (r * cols + cc)) & 1u);
        if (sim.all_abort()) {
            out[pat] = 1;
            n_prunable++;
        }
    }
    return n_prunable;
}

// how many of the starts halo_row_starts leaves on this plane would the table drop (cols as above)?  out[0] = starts, out[1] = dropped
void hs_prune_gain(const uint8_t* plane, int W, int H, int cols, const uint8_t* tabL, const uint8_t* tabR, int64_t* out) {
    HostPlane hp;
    pack_plane(plane, W, H, hp);
    const int half = cols / 2;
    const uint32_t cm = (1u << cols) - 1u;
    out[0] = out[1] = 0;
    for (int ty = 0; ty < hp.tiles_y; ty++)
        for (int r = 1; r <= FID_HALO_T; r++)
            for (int tx = 0; tx < hp.tpr; tx++) {
                const uint32_t* t = hp.halo.data() + ((size_t)ty * hp.tpr + tx) * 32 + r;
                uint32_t L = 0, R = 0;
                if (t[0]) halo_row_starts(t[-1], t[0], t[1], &L, &R);
                for (int i = 1; i <= FID_HALO_T; i++)
                    for (int side = 0; side < 2; side++) {
                        if (!(((side ? R : L) >> i) & 1)) continue;
                        out[0]++;
                        if (i - half < 0 || i + half > 31) continue;
                        const uint32_t pat = ((t[-1] >> (i - half)) & cm) | (((t[0] >> (i - half)) & cm) << cols) | (((t[1] >> (i - half)) & cm) << (2 * cols));
                        if ((side ? tabR : tabL)[pat]) out[1]++;
                    }
            }
}


// Experiment only: the same proof for a neighbourhood with `above` rows above the start and one below, 5 columns.
// out has 2^((above+2)*5) entries; pattern bit (row * 5 + col), row 0 = the topmost row.  Returns the number of prunable patterns.
long hs_prune_table_rows(int above, int K, int is_right, uint8_t* out) {
    const int cols = 5, half = 2, rows = above + 2, nbits = rows * cols;
    long n_prunable = 0;
    PruneSim sim;
    sim.K = K;
    sim.is_right = is_right;
    for (uint32_t pat = 0; pat < (1u << nbits); pat++) {
        out[pat] = 0;
        const int centre = (int)((pat >> (above * cols + half)) & 1u);
        const int side = (int)((pat >> (above * cols + half + (is_right ? 1 : -1))) & 1u);
        if (!centre || side) continue;
        for (int y = 0; y < PruneSim::N; y++)
            for (int x = 0; x < PruneSim::N; x++) sim.val[y][x] = -1;
        for (int r = 0; r < rows; r++)
            for (int cc = 0; cc < cols; cc++) sim.val[PruneSim::CY - above + r][PruneSim::CX - half + cc] = (int8_t)((pat >> (r * cols + cc)) & 1u);
        if (sim.all_abort()) {
            out[pat] = 1;
            n_prunable++;
        }
    }
    return n_prunable;
}

void hs_prune_gain_rows(const uint8_t* plane, int W, int H, int above, const uint8_t* tabL, const uint8_t* tabR, int64_t* out) {
    HostPlane hp;
    pack_plane(plane, W, H, hp);
    out[0] = out[1] = 0;
    for (int ty = 0; ty < hp.tiles_y; ty++)
        for (int r = 1; r <= FID_HALO_T; r++)
            for (int tx = 0; tx < hp.tpr; tx++) {
                const uint32_t* t = hp.halo.data() + ((size_t)ty * hp.tpr + tx) * 32 + r;
                uint32_t L = 0, R = 0;
                if (t[0]) halo_row_starts(t[-1], t[0], t[1], &L, &R);
                for (int i = 1; i <= FID_HALO_T; i++)
                    for (int side = 0; side < 2; side++) {
                        if (!(((side ? R : L) >> i) & 1)) continue;
                        out[0]++;
                        if (i - 2 < 0 || i + 2 > 31 || r - above < 0) continue;
                        uint32_t pat = 0;
                        for (int k = 0; k < above + 2; k++) pat |= ((t[k - above] >> (i - 2)) & 31u) << (5 * k);
                        if ((side ? tabR : tabL)[pat]) out[1]++;
                    }
            }
}

}  // extern "C"
