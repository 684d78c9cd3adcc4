// TMA variants probe (tools/, not part of the library): one variant per process, see main().
// (first lines below are shared with probe_umma.cu)
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


template <int RANK, int BW, int BH>
__global__ void k_tma(const __grid_constant__ CUtensorMap tmap, const CUtensorMap* gmap, int c0, int c1, int c2, uint32_t* out) {
    __shared__ __align__(1024) uint32_t box[BW * BH];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint64_t desc = gmap ? (uint64_t)gmap : (uint64_t)&tmap;
        mbar_expect_tx(&bar, BW * BH * 4);
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(box)), "l"(desc), "r"(c0),
                         "r"(c1), "r"(c2), "r"(smem_u32(&bar))
                         : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(box)), "l"(desc), "r"(c0), "r"(c1),
                         "r"(smem_u32(&bar))
                         : "memory");
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = box[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int RANK, int BW, int BH>
int run(int variant, bool use_global_map, int c0, int c1, int c2, CUtensorMapL2promotion l2) {
    const int W = 1920, H = 64, F = 2;
    std::vector<uint8_t> img((size_t)W * 3 * H * F);
    for (size_t i = 0; i < img.size(); i++) img[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t* dImg;
    uint32_t* dBox;
    CK(cudaMalloc(&dImg, img.size()));
    CK(cudaMalloc(&dBox, BW * BH * 4));
    CK(cudaMemcpy(dImg, img.data(), img.size(), cudaMemcpyHostToDevice));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    const cuuint64_t gdim[3] = {(cuuint64_t)(W * 3 / 4), (cuuint64_t)(RANK == 3 ? H : H * F), (cuuint64_t)F};
    const cuuint64_t gstr[2] = {(cuuint64_t)W * 3, (cuuint64_t)W * 3 * H};
    const cuuint32_t box[3] = {BW, BH, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult cr = ((EncodeTiledFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, RANK, dImg, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2,
                                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        printf("TMA variant %d: encode failed %d\n", variant, (int)cr);
        return 1;
    }
    const uint64_t* raw = (const uint64_t*)&tmap;
    printf("TMA variant %d map:", variant);
    for (int i = 0; i < 16; i++) printf(" %016llx", (unsigned long long)raw[i]);
    printf("\n");
    CUtensorMap* dMap = nullptr;
    if (use_global_map) {
        CK(cudaMalloc(&dMap, sizeof(CUtensorMap)));
        CK(cudaMemcpy(dMap, &tmap, sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    }
    k_tma<RANK, BW, BH><<<1, 128>>>(tmap, dMap, c0, c1, c2, dBox);
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("TMA variant %d: kernel FAILED: %s\n", variant, cudaGetErrorString(e));
        return 1;
    }
    std::vector<uint32_t> hb(BW * BH);
    CK(cudaMemcpy(hb.data(), dBox, hb.size() * 4, cudaMemcpyDeviceToHost));
    long bad = 0;
    const int Hh = RANK == 3 ? H : H * F;
    for (int y = 0; y < BH; y++)
        for (int x = 0; x < BW; x++) {
            const int gx = c0 + x, gy = c1 + y;
            uint32_t want = 0;
            if (gx >= 0 && gx < W * 3 / 4 && gy >= 0 && gy < Hh) memcpy(&want, &img[((size_t)(RANK == 3 ? c2 : 0) * H + gy) * W * 3 + (size_t)gx * 4], 4);
            if (hb[y * BW + x] != want) bad++;
        }
    printf("TMA variant %d (rank %d box %dx%d coords %d,%d,%d globalmap %d): %s (%ld bad)\n", variant, RANK, BW, BH, c0, c1, c2, (int)use_global_map, bad ? "FAIL" : "ok", bad);
    return bad ? 1 : 0;
}

int main(int argc, char** argv) {
    const int v = argc > 1 ? atoi(argv[1]) : 0;
    CK(cudaSetDevice(0));
    switch (v) {
        case 0: return run<3, 140, 16>(v, false, 600, 20, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
        case 1: return run<3, 140, 16>(v, false, -21, -10, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
        case 2: return run<2, 140, 16>(v, false, 600, 20, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
        case 3: return run<3, 128, 16>(v, false, 600, 20, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
        case 4: return run<3, 140, 16>(v, true, 600, 20, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
        case 5: return run<3, 140, 8>(v, false, 600, 20, 0, CU_TENSOR_MAP_L2_PROMOTION_NONE);
        case 6: return run<2, 128, 8>(v, true, 600, 20, 0, CU_TENSOR_MAP_L2_PROMOTION_NONE);
        case 7: return run<3, 140, 16>(v, false, 1400, 56, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
        case 8: return run<2, 64, 8>(v, false, 0, 0, 0, CU_TENSOR_MAP_L2_PROMOTION_NONE);
    }
    return 0;
}
