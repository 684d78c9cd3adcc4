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



struct Pad128 { uint64_t v[16]; };

// mode bits: 1 = tensor map is the SECOND kernel parameter (after a 128-byte struct); 2 = destination high in a 174 KB dynamic
// shared memory block; 4 = 148 CTAs; 8 = small tensor (480 x 480 x 1) with coordinates (69, 34, 0)
template <bool SECOND>
__global__ void __launch_bounds__(448, 1) k_tma2(const Pad128 pad, const __grid_constant__ CUtensorMap tmapB, int hi_smem, int c0, int c1, int c2, uint32_t* out, const __grid_constant__ CUtensorMap tmapA) {
    extern __shared__ __align__(1024) uint8_t dyn[];
    uint32_t* box = reinterpret_cast<uint32_t*>(dyn + (hi_smem ? 111104 : 0));
    uint64_t* bar = reinterpret_cast<uint64_t*>(dyn + (hi_smem ? 173568 : 16384));
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint64_t desc = SECOND ? (uint64_t)&tmapB : (uint64_t)&tmapA;
        mbar_expect_tx(bar, 140 * 16 * 4);
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(box)), "l"(desc), "r"(c0), "r"(c1),
                     "r"(c2), "r"(smem_u32(bar))
                     : "memory");
    }
    mbar_wait(bar, 0);
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < 140 * 16; i += blockDim.x) out[i] = box[i];
    if (pad.v[3] == 0x1234567) out[0] = 1;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    CK(cudaSetDevice(0));
    const bool small = mode & 8;
    const int W = small ? 640 : 1920, H = small ? 480 : 64, F = small ? 1 : 2;
    const bool neg = mode & 16;  // 16-byte aligned but NEGATIVE start coordinates (zero fill expected)
    const int c0 = neg ? -24 : (small ? 69 : 600), c1 = neg ? -26 : (small ? 34 : 20), c2 = 0;
    std::vector<uint8_t> img((size_t)W * 3 * H * F);
    for (size_t i = 0; i < img.size(); i++) img[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t* dImg;
    uint32_t* dBox;
    CK(cudaMalloc(&dImg, img.size()));
    CK(cudaMalloc(&dBox, 140 * 16 * 4));
    CK(cudaMemcpy(dImg, img.data(), img.size(), cudaMemcpyHostToDevice));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    const cuuint64_t gdim[3] = {(cuuint64_t)(W * 3 / 4), (cuuint64_t)H, (cuuint64_t)F};
    const cuuint64_t gstr[2] = {(cuuint64_t)W * 3, (cuuint64_t)W * 3 * H};
    const cuuint32_t box[3] = {140, 16, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult cr = ((EncodeTiledFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, dImg, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        printf("TMA2 mode %d: encode failed %d\n", mode, (int)cr);
        return 1;
    }
    Pad128 pad;
    memset(&pad, 0, sizeof(pad));
    const size_t smem = (mode & 2) ? 174080 : 32768;
    CK(cudaFuncSetAttribute(k_tma2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 174080));
    CK(cudaFuncSetAttribute(k_tma2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 174080));
    const int grid = (mode & 4) ? 148 : 1;
    if (mode & 1)
        k_tma2<true><<<grid, 448, smem>>>(pad, tmap, (mode & 2) ? 1 : 0, c0, c1, c2, dBox, tmap);
    else
        k_tma2<false><<<grid, 448, smem>>>(pad, tmap, (mode & 2) ? 1 : 0, c0, c1, c2, dBox, tmap);
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("TMA2 mode %d: kernel FAILED: %s\n", mode, cudaGetErrorString(e));
        return 1;
    }
    std::vector<uint32_t> hb(140 * 16);
    CK(cudaMemcpy(hb.data(), dBox, hb.size() * 4, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int y = 0; y < 16; y++)
        for (int x = 0; x < 140; x++) {
            const int gx = c0 + x, gy = c1 + y;
            uint32_t want = 0;
            if (gx >= 0 && gx < W * 3 / 4 && gy >= 0 && gy < H) memcpy(&want, &img[((size_t)c2 * H + gy) * W * 3 + (size_t)gx * 4], 4);
            if (hb[y * 140 + x] != want) bad++;
        }
    printf("TMA2 mode %d (second-param %d, high-smem %d, 148-ctas %d, small %d): %s (%ld bad)\n", mode, mode & 1, (mode >> 1) & 1, (mode >> 2) & 1, (mode >> 3) & 1, bad ? "FAIL" : "ok", bad);
    return bad ? 1 : 0;
}
