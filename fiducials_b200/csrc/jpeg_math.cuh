// JPEG ingest, arithmetic shared by the device kernels (kernels_jpeg.cuh) and the CPU unit-test harness (tests/hostsim):
// the integer pipeline cv::imdecode runs after entropy decoding (libjpeg's default decompression settings: accurate integer
// inverse DCT, "fancy" triangle-filter chroma upsampling, 16-bit fixed-point YCbCr -> RGB), restated from the published
// algorithm descriptions (Loeffler-Ligtenberg-Moschytz 8-point DCT with 13-bit constants; ITU-T T.81 / JFIF colour equations).
// Everything is integer arithmetic: results are bit-exact against cv2.imdecode (tests/test_hostsim_jpeg.py on the CPU,
// tests/test_gpu_jpeg.py on the device).
#pragma once
#include <stdint.h>

#include "common.cuh"

namespace fid {

FID_HD int jpeg_clamp255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// 1-D 8-point inverse DCT (LLM), 13-bit fixed-point constants; in/out stride given.  SHIFT_OUT and the rounding of the two passes
// are the caller's.
struct JpegIdctConst {
    static constexpr int C0_298 = 2446, C0_390 = 3196, C0_541 = 4433, C0_765 = 6270, C0_899 = 7373, C1_175 = 9633, C1_501 = 12299, C1_847 = 15137, C1_961 = 16069,
                         C2_053 = 16819, C2_562 = 20995, C3_072 = 25172;
};

FID_HD void jpeg_idct_1d(const int in[8], int out[8]) {
    typedef JpegIdctConst K;
    // even part
    int z2 = in[2], z3 = in[6];
    int z1 = (z2 + z3) * K::C0_541;
    const int t2 = z1 + z3 * (-K::C1_847);
    const int t3 = z1 + z2 * K::C0_765;
    z2 = in[0];
    z3 = in[4];
    const int t0 = (z2 + z3) * 8192;
    const int t1 = (z2 - z3) * 8192;
    const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    // odd part
    int o0 = in[7], o1 = in[5], o2 = in[3], o3 = in[1];
    z1 = o0 + o3;
    z2 = o1 + o2;
    z3 = o0 + o2;
    int z4 = o1 + o3;
    const int z5 = (z3 + z4) * K::C1_175;
    o0 *= K::C0_298;
    o1 *= K::C2_053;
    o2 *= K::C3_072;
    o3 *= K::C1_501;
    z1 *= -K::C0_899;
    z2 *= -K::C2_562;
    z3 *= -K::C1_961;
    z4 *= -K::C0_390;
    z3 += z5;
    z4 += z5;
    o0 += z1 + z3;
    o1 += z2 + z4;
    o2 += z2 + z3;
    o3 += z1 + z4;
    out[0] = t10 + o3;
    out[7] = t10 - o3;
    out[1] = t11 + o2;
    out[6] = t11 - o2;
    out[2] = t12 + o1;
    out[5] = t12 - o1;
    out[3] = t13 + o0;
    out[4] = t13 - o0;
}

// coef: dequantised coefficients, natural (row-major) order; out: 64 samples.  Pass 1 over columns keeps 2 extra bits
// (descale by 13 - 2), pass 2 over rows removes them together with the factor 8 of the 2-D transform (13 + 2 + 3) and adds the
// level shift of 128.
FID_HD void jpeg_idct_block(const int coef[64], uint8_t out[64]) {
    int ws[64];
    for (int c = 0; c < 8; c++) {
        int in[8], o[8];
        for (int r = 0; r < 8; r++) in[r] = coef[r * 8 + c];
        jpeg_idct_1d(in, o);
        for (int r = 0; r < 8; r++) ws[r * 8 + c] = (o[r] + (1 << 10)) >> 11;
    }
    for (int r = 0; r < 8; r++) {
        int o[8];
        jpeg_idct_1d(ws + r * 8, o);
        for (int c = 0; c < 8; c++) out[r * 8 + c] = (uint8_t)jpeg_clamp255(((o[c] + (1 << 17)) >> 18) + 128);
    }
}

// Chroma sample at full-resolution pixel (x, y) of a component plane `p` (pitch in bytes) whose real extent is cw x ch samples.
//   mode 0: no subsampling   mode 1: h2v1 (4:2:2)   mode 2: h2v2 (4:2:0)
// Triangle filter: 3/4 nearer sample + 1/4 farther sample in each subsampled direction; at the image edges the missing
// neighbour is the edge sample itself.  Rounding alternates (+1/+2, +8/+7) between even and odd output columns.
FID_HD int jpeg_chroma_at(const uint8_t* p, int pitch, int cw, int ch, int mode, int x, int y) {
    if (mode == 0) return p[(size_t)y * pitch + x];
    const int cx = x >> 1;
    if (mode == 1) {
        const uint8_t* row = p + (size_t)y * pitch;
        const int cur = row[cx];
        if (!(x & 1)) return cx == 0 ? cur : (3 * cur + row[cx - 1] + 1) >> 2;
        return cx == cw - 1 ? cur : (3 * cur + row[cx + 1] + 2) >> 2;
    }
    const int cy = y >> 1;
    int fy = (y & 1) ? cy + 1 : cy - 1;
    fy = fy < 0 ? 0 : (fy > ch - 1 ? ch - 1 : fy);
    const uint8_t* near_row = p + (size_t)cy * pitch;
    const uint8_t* far_row = p + (size_t)fy * pitch;
    const int cur = 3 * near_row[cx] + far_row[cx];
    if (!(x & 1)) {
        if (cx == 0) return (cur * 4 + 8) >> 4;
        return (3 * cur + (3 * near_row[cx - 1] + far_row[cx - 1]) + 8) >> 4;
    }
    if (cx == cw - 1) return (cur * 4 + 7) >> 4;
    return (3 * cur + (3 * near_row[cx + 1] + far_row[cx + 1]) + 7) >> 4;
}

// YCbCr -> B, G, R with 16-bit fixed-point coefficients 1.402, 0.71414, 0.34414, 1.772
FID_HD void jpeg_ycc_to_bgr(int y, int cb, int cr, uint8_t* bgr) {
    const int b = cb - 128, r = cr - 128;
    const int rr = (91881 * r + 32768) >> 16;
    const int bb = (116130 * b + 32768) >> 16;
    const int gg = (-22554 * b + 32768 - 46802 * r) >> 16;
    bgr[0] = (uint8_t)jpeg_clamp255(y + bb);
    bgr[1] = (uint8_t)jpeg_clamp255(y + gg);
    bgr[2] = (uint8_t)jpeg_clamp255(y + rr);
}

}  // namespace fid
