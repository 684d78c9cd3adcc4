// Hardware probe for the tensor-core threshold kernel (tools/, not part of the library):
//   1. tcgen05.mma kind::i8 (u8 x u8 -> s32), M=128, N=128, K=192 from no-swizzle K-major smem descriptors,
//      A given (a) fully materialised and (b) as an aliased Hankel table (SBO=128 B, LBO=256 B);
//      B in a "row-contiguous" layout (SBO=128 B, LBO=rows*16 B) read at a row offset
//   2. tcgen05.ld 32x32b.x16 at unaligned column offsets
//   3. TMEM read bandwidth with 1 / 4 / 8 warps, MMA issue rate for N=128 / 256
//   4. TMA (cp.async.bulk.tensor.3d) of a u32 view of BGR rows with a negative start coordinate
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_umma_bin tools/probe_umma.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__);            \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;                // no swizzle, base offset 0
}
__device__ __forceinline__ uint32_t make_idesc_u8(int M, int N) {
    // c_format S32 = 2 at [4,6); a/b format UINT8 = 0; K-major both; n_dim = N>>3 at [17,23); m_dim = M>>4 at [24,29)
    return (2u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* b) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]),
                   "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------
// test 1/2/3
// ---------------------------------------------------------------------------------------------------
#define KTOT 192
#define BROWS 160  // rows of the B tile in smem (row-contiguous layout: LBO = BROWS*16)
struct ProbeOut {
    int32_t d_full[128 * 128];    // A materialised, B rows 0..127
    int32_t d_hankel[128 * 128];  // A aliased Hankel, B rows 7..134 (row offset 7)
    int32_t d_unal[128 * 16];     // d_hankel columns 3..18 read with an unaligned tcgen05.ld
    long long clk[16];
};

__global__ void __launch_bounds__(256, 1) k_probe(const uint8_t* __restrict__ gB /* BROWS x KTOT */, int r, ProbeOut* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA_full = smem;                       // 128 rows x 192 B canonical: core(m8,kc) at (m8*12 + kc)*128
    uint8_t* sA_hank = sA_full + 128 * KTOT;       // 38 core matrices x 128 B (+ slack)
    uint8_t* sB = sA_hank + 40 * 128;              // row-contiguous: (n, i) at (i/16)*BROWS*16 + n*16 + i%16
    uint64_t* bar = (uint64_t*)(sB + BROWS * KTOT);
    uint32_t* tmem_slot = (uint32_t*)(bar + 4);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c0 = 154;  // band centre: A[m][i] = |i + m - c0| <= r
    for (int e = tid; e < 128 * KTOT; e += blockDim.x) {
        const int m = e / KTOT, i = e % KTOT;
        const int v = (abs(i + m - c0) <= r) ? 1 : 0;
        sA_full[((m >> 3) * 12 + (i >> 4)) * 128 + (m & 7) * 16 + (i & 15)] = (uint8_t)v;
    }
    for (int e = tid; e < 40 * 128; e += blockDim.x) {
        const int d = e >> 7, row = (e >> 4) & 7, col = e & 15;
        const int s = 8 * d + row + col;
        sA_hank[e] = (abs(s - c0) <= r) ? 1 : 0;
    }
    for (int e = tid; e < BROWS * KTOT; e += blockDim.x) {
        const int n = e / KTOT, i = e % KTOT;
        sB[(i >> 4) * BROWS * 16 + n * 16 + (i & 15)] = gB[e];
    }
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_init(&bar[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t idesc = make_idesc_u8(128, 128);
    if (tid == 0) {
        // D0 (cols 0..127): A full, B rows 0..
        for (int kk = 0; kk < KTOT / 32; kk++) {
            const uint64_t ad = make_desc(smem_u32(sA_full) + kk * 256, 128, 12 * 128);
            const uint64_t bd = make_desc(smem_u32(sB) + kk * 2 * BROWS * 16, BROWS * 16, 128);
            umma_i8(tmem + 0, ad, bd, idesc, kk > 0);
        }
        // D1 (cols 128..255): A Hankel alias, B rows 7..
        for (int kk = 0; kk < KTOT / 32; kk++) {
            const uint64_t ad = make_desc(smem_u32(sA_hank) + kk * 512, 256, 128);
            const uint64_t bd = make_desc(smem_u32(sB) + kk * 2 * BROWS * 16 + 7 * 16, BROWS * 16, 128);
            umma_i8(tmem + 128, ad, bd, idesc, kk > 0);
        }
        umma_commit(&bar[0]);
    }
    mbar_wait(&bar[0], 0);
    tc_fence_after();
    if (warp < 4) {
        uint32_t v[16];
        for (int c = 0; c < 128; c += 16) {
            tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
            tmem_ld_wait();
            for (int j = 0; j < 16; j++) out->d_full[(warp * 32 + lane) * 128 + c + j] = (int32_t)v[j];
            tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + 128 + c, v);
            tmem_ld_wait();
            for (int j = 0; j < 16; j++) out->d_hankel[(warp * 32 + lane) * 128 + c + j] = (int32_t)v[j];
        }
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + 128 + 3, v);  // unaligned column offset
        tmem_ld_wait();
        for (int j = 0; j < 16; j++) out->d_unal[(warp * 32 + lane) * 16 + j] = (int32_t)v[j];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // ---- timing: TMEM read bandwidth.  Each warp reads its lane quarter, 128 columns, REP times ----
    const int REP = 200;
    for (int cfg = 0; cfg < 3; cfg++) {
        const int nw = cfg == 0 ? 1 : (cfg == 1 ? 4 : 8);
        __syncthreads();
        const long long t0 = clock64();
        if (warp < nw) {
            uint32_t v[16], acc = 0;
            for (int it = 0; it < REP; it++)
                for (int c = 0; c < 128; c += 16) {
                    tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128 + c, v);
                    tmem_ld_wait();
                    acc += v[0] + v[7] + v[15];
                }
            if (acc == 0x12345678u) out->clk[15] = acc;
        }
        __syncthreads();
        const long long t1 = clock64();
        if (tid == 0) out->clk[cfg] = t1 - t0;  // bytes read = nw * REP * 128 cols * 32 lanes * 4 B
    }
    // same with two loads in flight before the wait
    for (int cfg = 0; cfg < 2; cfg++) {
        const int nw = cfg == 0 ? 4 : 8;
        __syncthreads();
        const long long t0 = clock64();
        if (warp < nw) {
            uint32_t v[16], w[16], acc = 0;
            for (int it = 0; it < REP; it++)
                for (int c = 0; c < 128; c += 32) {
                    tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128 + c, v);
                    tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128 + c + 16, w);
                    tmem_ld_wait();
                    acc += v[0] + v[7] + w[15];
                }
            if (acc == 0x12345678u) out->clk[15] = acc;
        }
        __syncthreads();
        const long long t1 = clock64();
        if (tid == 0) out->clk[3 + cfg] = t1 - t0;
    }
    // ---- timing: MMA issue rate, N = 128 and N = 256, 6 k-steps x 64 groups ----
    for (int cfg = 0; cfg < 2; cfg++) {
        __syncthreads();
        const long long t0 = clock64();
        if (tid == 0) {
            const uint32_t id2 = make_idesc_u8(128, cfg == 0 ? 128 : 256);
            for (int g = 0; g < 64; g++)
                for (int kk = 0; kk < KTOT / 32; kk++) {
                    const uint64_t ad = make_desc(smem_u32(sA_hank) + kk * 512, 256, 128);
                    const uint64_t bd = make_desc(smem_u32(sB) + kk * 2 * BROWS * 16, BROWS * 16, 128);  // N = 256 reads past 160 rows: garbage, timing only
                    umma_i8(tmem + (g & 1) * 256, ad, bd, id2, kk > 0);
                }
            umma_commit(&bar[1 + cfg]);
        }
        mbar_wait(&bar[1 + cfg], 0);
        const long long t1 = clock64();
        if (tid == 0) out->clk[5 + cfg] = t1 - t0;  // 384 MMAs
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// ---------------------------------------------------------------------------------------------------
// test 4: TMA of BGR rows viewed as u32
// ---------------------------------------------------------------------------------------------------
#define BOXW 140
#define BOXH 16
__global__ void k_tma(const __grid_constant__ CUtensorMap tmap, int c0, int c1, int c2, uint32_t* out) {
    __shared__ __align__(128) uint32_t box[BOXW * BOXH];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, BOXW * BOXH * 4);
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(box)),
                     "l"((uint64_t)&tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(&bar))
                     : "memory");
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < BOXW * BOXH; i += blockDim.x) out[i] = box[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    CK(cudaSetDevice(0));
    // ---------------- MMA probe ----------------
    std::vector<uint8_t> hB(BROWS * KTOT);
    srand(1);
    for (auto& b : hB) b = (uint8_t)(rand() & 255);
    uint8_t* dB;
    ProbeOut* dOut;
    CK(cudaMalloc(&dB, hB.size()));
    CK(cudaMalloc(&dOut, sizeof(ProbeOut)));
    CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dOut, 0xff, sizeof(ProbeOut)));
    const int r = 25;
    const size_t smem = 128 * KTOT + 40 * 128 + BROWS * KTOT + 64 + 1024 + 4096;
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_probe<<<1, 256, smem>>>(dB, r, dOut);
    CK(cudaDeviceSynchronize());
    std::vector<ProbeOut> ho(1);
    CK(cudaMemcpy(ho.data(), dOut, sizeof(ProbeOut), cudaMemcpyDeviceToHost));
    const ProbeOut& o = ho[0];
    long bad_full = 0, bad_hank = 0, bad_unal = 0;
    for (int m = 0; m < 128; m++)
        for (int n = 0; n < 128; n++) {
            int ref0 = 0, ref1 = 0;
            for (int i = 0; i < KTOT; i++)
                if (abs(i + m - 154) <= r) {
                    ref0 += hB[n * KTOT + i];
                    ref1 += hB[(n + 7) * KTOT + i];
                }
            if (o.d_full[m * 128 + n] != ref0) {
                if (bad_full < 5) printf("full mismatch m=%d n=%d got %d want %d\n", m, n, o.d_full[m * 128 + n], ref0);
                bad_full++;
            }
            if (o.d_hankel[m * 128 + n] != ref1) {
                if (bad_hank < 5) printf("hankel mismatch m=%d n=%d got %d want %d\n", m, n, o.d_hankel[m * 128 + n], ref1);
                bad_hank++;
            }
            if (n >= 3 && n < 19 && o.d_unal[m * 16 + n - 3] != ref1) {
                if (bad_unal < 5) printf("unaligned-ld mismatch m=%d n=%d got %d want %d\n", m, n, o.d_unal[m * 16 + n - 3], ref1);
                bad_unal++;
            }
        }
    printf("PROBE mma_full_A: %s (%ld bad)\n", bad_full ? "FAIL" : "ok", bad_full);
    printf("PROBE mma_hankel_A_rowoffset_B: %s (%ld bad)\n", bad_hank ? "FAIL" : "ok", bad_hank);
    printf("PROBE tmem_ld_unaligned: %s (%ld bad)\n", bad_unal ? "FAIL" : "ok", bad_unal);
    const double bytes1 = 200.0 * 128 * 32 * 4;
    printf("PROBE tmem_read  1 warp : %lld clk -> %.1f B/clk\n", o.clk[0], bytes1 * 1 / o.clk[0]);
    printf("PROBE tmem_read  4 warps: %lld clk -> %.1f B/clk\n", o.clk[1], bytes1 * 4 / o.clk[1]);
    printf("PROBE tmem_read  8 warps: %lld clk -> %.1f B/clk\n", o.clk[2], bytes1 * 8 / o.clk[2]);
    printf("PROBE tmem_read2 4 warps: %lld clk -> %.1f B/clk (2 loads in flight)\n", o.clk[3], bytes1 * 4 / o.clk[3]);
    printf("PROBE tmem_read2 8 warps: %lld clk -> %.1f B/clk (2 loads in flight)\n", o.clk[4], bytes1 * 8 / o.clk[4]);
    printf("PROBE mma N=128: %lld clk / 384 MMAs = %.1f clk each\n", o.clk[5], o.clk[5] / 384.0);
    printf("PROBE mma N=256: %lld clk / 384 MMAs = %.1f clk each\n", o.clk[6], o.clk[6] / 384.0);

    // ---------------- TMA probe ----------------
    const int W = 1920, H = 64, F = 2;
    std::vector<uint8_t> img((size_t)W * 3 * H * F);
    for (size_t i = 0; i < img.size(); i++) img[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t* dImg;
    uint32_t* dBox;
    CK(cudaMalloc(&dImg, img.size()));
    CK(cudaMalloc(&dBox, BOXW * BOXH * 4));
    CK(cudaMemcpy(dImg, img.data(), img.size(), cudaMemcpyHostToDevice));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
        printf("PROBE tma: cuTensorMapEncodeTiled not available\n");
        return 1;
    }
    CUtensorMap tmap;
    const cuuint64_t gdim[3] = {(cuuint64_t)(W * 3 / 4), (cuuint64_t)H, (cuuint64_t)F};
    const cuuint64_t gstr[2] = {(cuuint64_t)W * 3, (cuuint64_t)W * 3 * H};
    const cuuint32_t box[3] = {BOXW, BOXH, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult cr = ((EncodeTiledFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, dImg, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        printf("PROBE tma: encode failed %d\n", (int)cr);
        return 1;
    }
    const int cases[3][3] = {{-21, -10, 1}, {600, 20, 0}, {1400, 56, 1}};  // left/top OOB, interior, right/bottom OOB
    for (int t = 0; t < 3; t++) {
        const int c0 = cases[t][0], c1 = cases[t][1], c2 = cases[t][2];
        k_tma<<<1, 128>>>(tmap, c0, c1, c2, dBox);
        CK(cudaDeviceSynchronize());
        std::vector<uint32_t> hb(BOXW * BOXH);
        CK(cudaMemcpy(hb.data(), dBox, hb.size() * 4, cudaMemcpyDeviceToHost));
        long bad = 0;
        for (int y = 0; y < BOXH; y++)
            for (int x = 0; x < BOXW; x++) {
                const int gx = c0 + x, gy = c1 + y;
                uint32_t want = 0;
                if (gx >= 0 && gx < W * 3 / 4 && gy >= 0 && gy < H) memcpy(&want, &img[((size_t)c2 * H + gy) * W * 3 + (size_t)gx * 4], 4);
                if (hb[y * BOXW + x] != want) bad++;
            }
        printf("PROBE tma case %d (%d,%d,%d): %s (%ld bad)\n", t, c0, c1, c2, bad ? "FAIL" : "ok", bad);
    }
    return 0;
}
