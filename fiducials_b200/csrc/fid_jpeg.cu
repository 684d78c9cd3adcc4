// JPEG ingest: fid_jpeg_* of include/fiducials_b200.h.
//
// The reference's default image transport is `compressed` (aruco_detect/launch/aruco_detect.launch:6,28):
// compressed_image_transport decodes every frame with cv::imdecode on one host thread and hands the BGR8 image to
// imageCallback (aruco_detect.cpp:332,348).  Here the serial part of JPEG decoding -- the Huffman bit stream -- runs on host
// threads, one image per thread (jpeg_host.hpp), and produces the quantised coefficients in a sparse form that is typically
// 4-6x smaller than the decoded frame; that is what crosses PCIe.  The device does the rest (jpeg_math.cuh: dequantisation,
// integer inverse DCT, triangle-filter chroma upsampling, fixed-point YCbCr -> BGR) and writes BGR8 frames into the device buffer
// that fid_submit_batch(..., bgr_on_device = 1) consumes.  Bit-exact against cv2.imdecode (tests/test_gpu_jpeg.py).
#include <cuda_runtime.h>
#include <stdio.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

#include "../../include/fiducials_b200.h"
#include "jpeg_host.hpp"
#include "jpeg_math.cuh"

using namespace fid;

#define CKJ(call)                                                                                      \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            fprintf(stderr, "[fiducials_b200] CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return FID_ERR_CUDA;                                                                       \
        }                                                                                              \
    } while (0)

namespace {

struct JpegGeom {       // identical for every image of a batch
    int W, H, ncomp, mode;  // mode 0: 4:4:4 / grey, 1: h2v1, 2: h2v2
    int bw[3], bh[3], blk_base[3], nblk;
    int cw[3], ch[3];
    int plane_off[3];   // byte offset of every component plane inside a frame's plane block
    int plane_bytes;
};

struct IdctArgs {
    JpegGeom g;
    int n_frames;
    const uint64_t* mask;    // [n][blk_cap]
    const uint32_t* off;     // [n][blk_cap]
    const int16_t* vals;     // [n][val_cap]
    const uint16_t* quant;   // [n][3][64], zig-zag order
    const int32_t* status;   // [n]: frames that failed on the host are skipped
    size_t blk_cap, val_cap;
    uint8_t* planes;         // [n][plane_stride]
    size_t plane_stride;
};

__constant__ uint8_t c_zigzag[64];

// One thread per 8x8 block: gather the sparse coefficients (zig-zag order), dequantise, inverse DCT, store 8 rows of 8 bytes.
__global__ void __launch_bounds__(128) k_jpeg_idct(const IdctArgs a) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)a.n_frames * a.g.nblk) return;
    const int f = (int)(gid / a.g.nblk), b = (int)(gid - (long long)f * a.g.nblk);
    if (a.status[f] != FID_OK) return;
    const int c = (a.g.ncomp == 3 && b >= a.g.blk_base[2]) ? 2 : ((a.g.ncomp == 3 && b >= a.g.blk_base[1]) ? 1 : 0);
    const int lb = b - a.g.blk_base[c];
    const int by = lb / a.g.bw[c], bx = lb - by * a.g.bw[c];
    int coef[64];
#pragma unroll
    for (int i = 0; i < 64; i++) coef[i] = 0;
    uint64_t m = a.mask[(size_t)f * a.blk_cap + b];
    const int16_t* v = a.vals + (size_t)f * a.val_cap + a.off[(size_t)f * a.blk_cap + b];
    const uint16_t* q = a.quant + ((size_t)f * 3 + c) * 64;
    while (m) {
        const int k = __ffsll((long long)m) - 1;
        m &= m - 1;
        coef[c_zigzag[k]] = (int)(*v++) * (int)q[k];
    }
    uint8_t out[64];
    jpeg_idct_block(coef, out);
    const int pitch = a.g.bw[c] * 8;
    uint8_t* dst = a.planes + (size_t)f * a.plane_stride + a.g.plane_off[c] + (size_t)(by * 8) * pitch + bx * 8;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        uint2 w;
        w.x = out[r * 8] | (out[r * 8 + 1] << 8) | (out[r * 8 + 2] << 16) | ((uint32_t)out[r * 8 + 3] << 24);
        w.y = out[r * 8 + 4] | (out[r * 8 + 5] << 8) | (out[r * 8 + 6] << 16) | ((uint32_t)out[r * 8 + 7] << 24);
        *reinterpret_cast<uint2*>(dst + (size_t)r * pitch) = w;
    }
}

struct ColorArgs {
    JpegGeom g;
    int n_frames;
    const int32_t* status;
    const uint8_t* planes;
    size_t plane_stride;
    uint8_t* bgr;
    size_t row_stride, frame_stride;
    int aligned4;
};

// One thread per 4 horizontal pixels: chroma upsampling + colour conversion, 12 bytes of BGR out.
__global__ void __launch_bounds__(256) k_jpeg_color(const ColorArgs a) {
    const int quads = (a.g.W + 3) >> 2;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)a.n_frames * a.g.H * quads) return;
    const int qx = (int)(gid % quads);
    const long long t = gid / quads;
    const int y = (int)(t % a.g.H), f = (int)(t / a.g.H);
    if (a.status[f] != FID_OK) return;
    const uint8_t* P = a.planes + (size_t)f * a.plane_stride;
    const uint8_t* Y = P + a.g.plane_off[0] + (size_t)y * a.g.bw[0] * 8;
    uint8_t px[12];
    const int x0 = qx * 4;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int x = x0 + k < a.g.W ? x0 + k : a.g.W - 1;
        const int yy = Y[x];
        if (a.g.ncomp == 1) {
            px[3 * k] = px[3 * k + 1] = px[3 * k + 2] = (uint8_t)yy;
        } else {
            const int cb = jpeg_chroma_at(P + a.g.plane_off[1], a.g.bw[1] * 8, a.g.cw[1], a.g.ch[1], a.g.mode, x, y);
            const int cr = jpeg_chroma_at(P + a.g.plane_off[2], a.g.bw[2] * 8, a.g.cw[2], a.g.ch[2], a.g.mode, x, y);
            jpeg_ycc_to_bgr(yy, cb, cr, px + 3 * k);
        }
    }
    uint8_t* dst = a.bgr + (size_t)f * a.frame_stride + (size_t)y * a.row_stride + 3 * (size_t)x0;
    if (a.aligned4 && x0 + 3 < a.g.W) {
        uint32_t* d = reinterpret_cast<uint32_t*>(dst);
        d[0] = px[0] | (px[1] << 8) | (px[2] << 16) | ((uint32_t)px[3] << 24);
        d[1] = px[4] | (px[5] << 8) | (px[6] << 16) | ((uint32_t)px[7] << 24);
        d[2] = px[8] | (px[9] << 8) | (px[10] << 16) | ((uint32_t)px[11] << 24);
    } else {
        for (int k = 0; k < 4 && x0 + k < a.g.W; k++) {
            dst[3 * k] = px[3 * k];
            dst[3 * k + 1] = px[3 * k + 1];
            dst[3 * k + 2] = px[3 * k + 2];
        }
    }
}

}  // namespace

struct fid_jpeg {
    int device = 0, max_w = 0, max_h = 0, max_batch = 0, n_threads = 1;
    size_t blk_cap = 0, val_cap = 0, plane_stride = 0;
    cudaStream_t stream = nullptr;
    // two staging sets: the host decodes batch k+1 into one while the copies of batch k drain from the other
    struct Set {
        uint64_t* h_mask = nullptr;
        uint32_t* h_off = nullptr;
        int16_t* h_vals = nullptr;
        uint16_t* h_quant = nullptr;
        int32_t* h_status = nullptr;
        cudaEvent_t drained = nullptr;
    } set[2];
    int next_set = 0;
    uint64_t* d_mask = nullptr;
    uint32_t* d_off = nullptr;
    int16_t* d_vals = nullptr;
    uint16_t* d_quant = nullptr;
    int32_t* d_status = nullptr;
    uint8_t* d_planes = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double last_host_ms = 0, last_h2d_bytes = 0;
    bool timed = false;
};

extern "C" {

int fid_jpeg_destroy(fid_jpeg* j) {
    if (!j) return FID_OK;
    cudaSetDevice(j->device);
    if (j->stream) cudaStreamSynchronize(j->stream);
    for (auto& s : j->set) {
        if (s.h_mask) cudaFreeHost(s.h_mask);
        if (s.h_off) cudaFreeHost(s.h_off);
        if (s.h_vals) cudaFreeHost(s.h_vals);
        if (s.h_quant) cudaFreeHost(s.h_quant);
        if (s.h_status) cudaFreeHost(s.h_status);
        if (s.drained) cudaEventDestroy(s.drained);
    }
    if (j->d_mask) cudaFree(j->d_mask);
    if (j->d_off) cudaFree(j->d_off);
    if (j->d_vals) cudaFree(j->d_vals);
    if (j->d_quant) cudaFree(j->d_quant);
    if (j->d_status) cudaFree(j->d_status);
    if (j->d_planes) cudaFree(j->d_planes);
    if (j->ev0) cudaEventDestroy(j->ev0);
    if (j->ev1) cudaEventDestroy(j->ev1);
    if (j->stream) cudaStreamDestroy(j->stream);
    cudaGetLastError();
    delete j;
    return FID_OK;
}

int fid_jpeg_create(int device, int max_width, int max_height, int max_batch, int n_threads, fid_jpeg** out) {
    if (!out || max_width < 1 || max_height < 1 || max_batch < 1 || max_width > 65535 || max_height > 65535) return FID_ERR_INVALID_ARG;
    *out = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) {
        cudaGetLastError();
        return FID_ERR_NO_DEVICE;  // no CPU fallback
    }
    if (cudaSetDevice(device) != cudaSuccess) {
        cudaGetLastError();
        return FID_ERR_NO_DEVICE;
    }
    fid_jpeg* j = new fid_jpeg();
    j->device = device;
    j->max_w = max_width;
    j->max_h = max_height;
    j->max_batch = max_batch;
    const unsigned hw = std::thread::hardware_concurrency();
    j->n_threads = n_threads > 0 ? n_threads : (int)std::max(1u, std::min(hw ? hw : 1u, 64u));
    // worst case: 4:4:4, every coefficient non-zero
    const size_t bwm = (size_t)(max_width + 15) / 16 * 2, bhm = (size_t)(max_height + 15) / 16 * 2;
    j->blk_cap = bwm * bhm * 3;
    j->val_cap = j->blk_cap * 64;
    j->plane_stride = j->blk_cap * 64;
    bool ok = cudaStreamCreateWithFlags(&j->stream, cudaStreamNonBlocking) == cudaSuccess;
    const size_t n = (size_t)max_batch;
    for (auto& s : j->set) {
        ok = ok && cudaHostAlloc((void**)&s.h_mask, n * j->blk_cap * 8, cudaHostAllocDefault) == cudaSuccess;
        ok = ok && cudaHostAlloc((void**)&s.h_off, n * j->blk_cap * 4, cudaHostAllocDefault) == cudaSuccess;
        ok = ok && cudaHostAlloc((void**)&s.h_vals, n * j->val_cap * 2, cudaHostAllocDefault) == cudaSuccess;
        ok = ok && cudaHostAlloc((void**)&s.h_quant, n * 3 * 64 * 2, cudaHostAllocDefault) == cudaSuccess;
        ok = ok && cudaHostAlloc((void**)&s.h_status, n * 4, cudaHostAllocDefault) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&s.drained, cudaEventDisableTiming) == cudaSuccess;
    }
    ok = ok && cudaMalloc((void**)&j->d_mask, n * j->blk_cap * 8) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&j->d_off, n * j->blk_cap * 4) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&j->d_vals, n * j->val_cap * 2) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&j->d_quant, n * 3 * 64 * 2) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&j->d_status, n * 4) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&j->d_planes, n * j->plane_stride) == cudaSuccess;
    ok = ok && cudaEventCreate(&j->ev0) == cudaSuccess && cudaEventCreate(&j->ev1) == cudaSuccess;
    ok = ok && cudaMemcpyToSymbol(c_zigzag, fidjpeg::kZigzag, 64) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        fid_jpeg_destroy(j);
        return FID_ERR_NO_MEMORY;
    }
    *out = j;
    return FID_OK;
}

int fid_jpeg_decode_batch(fid_jpeg* j, int n, const uint8_t* const* data, const size_t* bytes, int width, int height, void* device_bgr, size_t row_stride, size_t frame_stride,
                          int32_t* status) {
    if (!j || n < 0 || (n > 0 && (!data || !bytes || !device_bgr))) return FID_ERR_INVALID_ARG;
    if (n == 0) return FID_OK;
    if (n > j->max_batch || width > j->max_w || height > j->max_h) return FID_ERR_CAPACITY;
    if (width < 1 || height < 1 || row_stride < (size_t)width * 3 || frame_stride < row_stride * (size_t)height) return FID_ERR_INVALID_ARG;
    CKJ(cudaSetDevice(j->device));
    fid_jpeg::Set& S = j->set[j->next_set];
    j->next_set ^= 1;
    CKJ(cudaEventSynchronize(S.drained));  // the copies of the batch that used this set two calls ago
    // ---- host: entropy decoding, one image per task ----
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<fidjpeg::FrameInfo> info((size_t)n);
    std::vector<size_t> nvals((size_t)n, 0);
    std::atomic<int> next{0};
    auto work = [&]() {
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= n) break;
            size_t nv = 0;
            const int rc = fidjpeg::decode_image(data[i], bytes[i], &info[i], S.h_mask + (size_t)i * j->blk_cap, S.h_off + (size_t)i * j->blk_cap, S.h_vals + (size_t)i * j->val_cap,
                                                 j->val_cap, j->blk_cap, &nv);
            int st = rc == fidjpeg::JPEG_OK ? FID_OK : (rc == fidjpeg::JPEG_UNSUPPORTED ? FID_ERR_UNSUPPORTED : (rc == fidjpeg::JPEG_CAPACITY ? FID_ERR_CAPACITY : FID_ERR_INVALID_ARG));
            if (st == FID_OK && (info[i].W != width || info[i].H != height)) st = FID_ERR_INVALID_ARG;
            S.h_status[i] = st;
            nvals[i] = st == FID_OK ? nv : 0;
            if (st == FID_OK)
                for (int c = 0; c < info[i].ncomp; c++) memcpy(S.h_quant + ((size_t)i * 3 + c) * 64, info[i].q[c], 128);
        }
    };
    const int nt = std::min(j->n_threads, n);
    if (nt <= 1) {
        work();
    } else {
        std::vector<std::thread> pool;
        pool.reserve((size_t)nt - 1);
        for (int t = 1; t < nt; t++) pool.emplace_back(work);
        work();
        for (auto& t : pool) t.join();
    }
    // geometry of the batch = geometry of its first good image; images that differ are rejected
    int first = -1;
    for (int i = 0; i < n && first < 0; i++)
        if (S.h_status[i] == FID_OK) first = i;
    JpegGeom g{};
    if (first >= 0) {
        const fidjpeg::FrameInfo& fi = info[first];
        g.W = fi.W;
        g.H = fi.H;
        g.ncomp = fi.ncomp;
        g.mode = fi.ncomp == 1 ? 0 : (fi.hmax == 1 ? 0 : (fi.vmax == 1 ? 1 : 2));
        g.nblk = fi.nblk;
        int po = 0;
        for (int c = 0; c < fi.ncomp; c++) {
            g.bw[c] = fi.bw[c];
            g.bh[c] = fi.bh[c];
            g.blk_base[c] = fi.blk_base[c];
            g.cw[c] = fi.cw[c];
            g.ch[c] = fi.ch[c];
            g.plane_off[c] = po;
            po += fi.bw[c] * 8 * fi.bh[c] * 8;
        }
        g.plane_bytes = po;
        for (int i = 0; i < n; i++) {
            if (S.h_status[i] != FID_OK) continue;
            const fidjpeg::FrameInfo& o = info[i];
            if (o.ncomp != fi.ncomp || o.hmax != fi.hmax || o.vmax != fi.vmax) {
                S.h_status[i] = FID_ERR_UNSUPPORTED;  // mixed sampling inside one batch
                nvals[i] = 0;
            }
        }
    }
    j->last_host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    int worst = FID_OK;
    for (int i = 0; i < n; i++) {
        if (status) status[i] = S.h_status[i];
        if (S.h_status[i] != FID_OK && worst == FID_OK) worst = S.h_status[i];
    }
    // ---- device: copies of the used parts, inverse DCT, upsampling + colour ----
    cudaStream_t st = j->stream;
    CKJ(cudaEventRecord(j->ev0, st));
    double h2d = 0;
    CKJ(cudaMemcpyAsync(j->d_status, S.h_status, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CKJ(cudaMemcpyAsync(j->d_quant, S.h_quant, (size_t)n * 3 * 64 * 2, cudaMemcpyHostToDevice, st));
    h2d += (double)n * (4 + 384);
    if (first >= 0) {
        for (int i = 0; i < n; i++) {
            if (S.h_status[i] != FID_OK) continue;
            CKJ(cudaMemcpyAsync(j->d_mask + (size_t)i * j->blk_cap, S.h_mask + (size_t)i * j->blk_cap, (size_t)g.nblk * 8, cudaMemcpyHostToDevice, st));
            CKJ(cudaMemcpyAsync(j->d_off + (size_t)i * j->blk_cap, S.h_off + (size_t)i * j->blk_cap, (size_t)g.nblk * 4, cudaMemcpyHostToDevice, st));
            if (nvals[i]) CKJ(cudaMemcpyAsync(j->d_vals + (size_t)i * j->val_cap, S.h_vals + (size_t)i * j->val_cap, nvals[i] * 2, cudaMemcpyHostToDevice, st));
            h2d += (double)g.nblk * 12 + (double)nvals[i] * 2;
        }
    }
    CKJ(cudaEventRecord(S.drained, st));
    if (first >= 0) {
        IdctArgs ia{};
        ia.g = g;
        ia.n_frames = n;
        ia.mask = j->d_mask;
        ia.off = j->d_off;
        ia.vals = j->d_vals;
        ia.quant = j->d_quant;
        ia.status = j->d_status;
        ia.blk_cap = j->blk_cap;
        ia.val_cap = j->val_cap;
        ia.planes = j->d_planes;
        ia.plane_stride = j->plane_stride;
        const long long nb = (long long)n * g.nblk;
        k_jpeg_idct<<<(unsigned int)((nb + 127) / 128), 128, 0, st>>>(ia);
        ColorArgs ca{};
        ca.g = g;
        ca.n_frames = n;
        ca.status = j->d_status;
        ca.planes = j->d_planes;
        ca.plane_stride = j->plane_stride;
        ca.bgr = static_cast<uint8_t*>(device_bgr);
        ca.row_stride = row_stride;
        ca.frame_stride = frame_stride;
        ca.aligned4 = (row_stride % 4 == 0) && (frame_stride % 4 == 0) && ((uintptr_t)device_bgr % 4 == 0);
        const long long nq = (long long)n * g.H * ((g.W + 3) / 4);
        k_jpeg_color<<<(unsigned int)((nq + 255) / 256), 256, 0, st>>>(ca);
        CKJ(cudaGetLastError());
    }
    CKJ(cudaEventRecord(j->ev1, st));
    j->last_h2d_bytes = h2d;
    j->timed = true;
    return worst == FID_OK ? FID_OK : (first >= 0 ? FID_OK : worst);  // per-image failures are reported through status[]
}

int fid_jpeg_sync(fid_jpeg* j) {
    if (!j) return FID_ERR_INVALID_ARG;
    CKJ(cudaSetDevice(j->device));
    CKJ(cudaStreamSynchronize(j->stream));
    return FID_OK;
}

int fid_jpeg_stream(fid_jpeg* j, void** cuda_stream) {
    if (!j || !cuda_stream) return FID_ERR_INVALID_ARG;
    *cuda_stream = (void*)j->stream;
    return FID_OK;
}

int fid_jpeg_last_stats(fid_jpeg* j, double* host_decode_ms, double* h2d_bytes, double* device_ms) {
    if (!j) return FID_ERR_INVALID_ARG;
    if (host_decode_ms) *host_decode_ms = j->last_host_ms;
    if (h2d_bytes) *h2d_bytes = j->last_h2d_bytes;
    if (device_ms) {
        *device_ms = 0;
        if (j->timed) {
            CKJ(cudaSetDevice(j->device));
            CKJ(cudaEventSynchronize(j->ev1));
            float ms = 0;
            CKJ(cudaEventElapsedTime(&ms, j->ev0, j->ev1));
            *device_ms = ms;
        }
    }
    return FID_OK;
}

}  // extern "C"
