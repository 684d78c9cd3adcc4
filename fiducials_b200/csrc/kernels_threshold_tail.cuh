// Tail of the threshold kernels (kernels_threshold.cuh, kernels_threshold_mma.cuh): a warp holds the 32x32-bit
// halo tile of every scale in registers (lane r = tile row r).
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"
#include "contour_walk.cuh"
#include "kernels_contour.cuh"  // StartRec, Counters

namespace fid {

// stores the 13 tiles (one 128-byte transaction each) and queues the start cracks of the border walk
template <int NS, bool PRUNE = false>
__device__ __forceinline__ void thr_store_tile_and_starts(const uint32_t (&acc)[NS], int n_scales, int f, int tx, int ty, int lane, uint32_t* halo, size_t halo_frame_stride,
                                                          size_t halo_scale_stride, int halo_tpr, int halo_tiles_y, StartRec* starts, Counters* counters, unsigned int max_starts,
                                                          const uint32_t* prune = nullptr /* start_prune_table.h on the device, or off */) {
    uint32_t* out = halo + (size_t)f * halo_frame_stride + ((size_t)ty * halo_tpr + tx) * 32 + lane;
    int cnt_l = 0, cnt_r = 0;
    const bool row_ok = lane >= 1 && lane <= FID_HALO_T;
    uint32_t Lm[NS], Rm[NS];  // start cracks of this lane's row, per scale (kept for the emission below)
#pragma unroll
    for (int s = 0; s < NS; s++) {
        Lm[s] = Rm[s] = 0;
        if (s >= n_scales) continue;
        out[(size_t)s * halo_scale_stride] = acc[s];
        const uint32_t up = __shfl_up_sync(0xffffffffu, acc[s], 1), dn = __shfl_down_sync(0xffffffffu, acc[s], 1);
        if (row_ok && acc[s]) halo_row_starts(up, acc[s], dn, &Lm[s], &Rm[s]);
        if (PRUNE && (Lm[s] | Rm[s])) halo_prune_starts(up, acc[s], dn, &Lm[s], &Rm[s], prune);  // opt-in second stage (FID_START_PRUNE=1)
        cnt_l += __popc(Lm[s]);
        cnt_r += __popc(Rm[s]);
    }
    // one queue reservation per warp and side (all scales of the tile): left cracks grow from the front of
    // the buffer, right cracks from the back, so that every warp of the walk kernels sees a single direction
    int incl_l = cnt_l, incl_r = cnt_r;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int tl = __shfl_up_sync(0xffffffffu, incl_l, d), tr = __shfl_up_sync(0xffffffffu, incl_r, d);
        if (lane >= d) {
            incl_l += tl;
            incl_r += tr;
        }
    }
    unsigned int base_l = 0, base_r = 0;
    if (lane == 31) {
        base_l = incl_l ? atomicAdd(&counters->n_starts[0], (unsigned int)incl_l) : 0u;
        base_r = incl_r ? atomicAdd(&counters->n_starts[1], (unsigned int)incl_r) : 0u;
    }
    unsigned int pos_l = __shfl_sync(0xffffffffu, base_l, 31) + (unsigned int)(incl_l - cnt_l);
    unsigned int pos_r = __shfl_sync(0xffffffffu, base_r, 31) + (unsigned int)(incl_r - cnt_r);
    const uint32_t tile_id = (uint32_t)((f * halo_tiles_y + ty) * halo_tpr + tx);
    const unsigned int cap = max_starts / 2;
    bool overflow = false;
#pragma unroll
    for (int s = 0; s < NS; s++) {
        uint32_t L = Lm[s], Rr = Rm[s];
        if (!(L | Rr)) continue;
        const uint32_t rec = start_rec_pack(tile_id, (uint32_t)s, (uint32_t)lane, 0u);
        while (L) {
            const int i = __ffs(L) - 1;
            L &= L - 1;
            if (pos_l < cap)
                starts[pos_l].v = rec | (uint32_t)i;
            else
                overflow = true;
            pos_l++;
        }
        while (Rr) {
            const int i = __ffs(Rr) - 1;
            Rr &= Rr - 1;
            if (pos_r < cap)
                starts[max_starts - 1 - pos_r].v = rec | (uint32_t)i;
            else
                overflow = true;
            pos_r++;
        }
    }
    if (overflow) atomicOr(&counters->overflow, 1u);
}


}  // namespace fid
