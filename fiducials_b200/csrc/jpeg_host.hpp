// JPEG ingest, host half: the reference's default image transport is `compressed`
// (aruco_detect/launch/aruco_detect.launch:6,28) -- compressed_image_transport hands cv::imdecode's BGR8 frame to
// imageCallback (aruco_detect.cpp:332,348).  Entropy decoding is a serial bit stream per image and stays on the host (one
// thread per image); everything after it -- dequantisation, inverse DCT, chroma upsampling, colour conversion -- runs on the
// device (kernels_jpeg.cuh) and writes BGR8 frames straight into HBM.
//
// This header: marker parser + baseline Huffman decoder.  Output per image = the quantised coefficients in a SPARSE form
// (per 8x8 block a 64-bit mask of the non-zero zig-zag positions + the values in zig-zag order), typically 4-6x smaller than
// the decoded frame: that, not the frame, is what crosses PCIe.
//
// Supported: baseline / extended-sequential 8-bit Huffman JPEG (SOF0, SOF1), one interleaved scan, grey or YCbCr with
// 4:4:4, 4:2:2 (h2v1) or 4:2:0 (h2v2) sampling, restart intervals.  Anything else (progressive, arithmetic, 12-bit, CMYK, other
// sampling, multi-scan) returns JPEG_UNSUPPORTED -- the caller falls back to its own decoder for that image.
#pragma once
#include <stdint.h>
#include <string.h>

namespace fidjpeg {

enum { JPEG_OK = 0, JPEG_BAD = -1, JPEG_UNSUPPORTED = -2, JPEG_CAPACITY = -3 };

// zig-zag position -> natural (row-major) index
static const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                                    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct FrameInfo {
    int W, H, ncomp;
    int hs[3], vs[3], tq[3];
    int hmax, vmax;
    int mcux, mcuy;          // MCUs per row / column
    int bw[3], bh[3];        // blocks per row / column of every component plane (whole MCUs)
    int blk_base[3], nblk;   // component-major block numbering
    int cw[3], ch[3];        // "downsampled" width / height of every component: ceil(W * hs / hmax), ceil(H * vs / vmax)
    uint16_t q[3][64];       // quantisation tables per component, ZIG-ZAG order (as stored in the file)
};

#define FIDJPEG_LOOK 10  // look-ahead bits of the symbol tables

struct HuffTable {
    bool present = false;
    uint8_t vals[256];
    int mincode[17], maxcode[18], valptr[17];
    uint16_t look[1 << FIDJPEG_LOOK];    // (length << 8) | symbol, 0 = code longer than the look-ahead
    int32_t fast_ac[1 << FIDJPEG_LOOK];  // AC tables: code AND magnitude bits inside the look-ahead -> value * 256 + run * 16 + total bits, else 0

    // false = not a prefix code (more codes of some length than the code space holds): the stream is rejected
    bool build(const uint8_t bits[17], const uint8_t* symbols, int nsym) {
        memset(vals, 0, sizeof(vals));
        memcpy(vals, symbols, nsym);
        memset(look, 0, sizeof(look));
        int code = 0, k = 0;
        for (int l = 1; l <= 16; l++) {
            if (code + bits[l] > (1 << l)) return false;
            mincode[l] = code;
            valptr[l] = k;
            for (int i = 0; i < bits[l]; i++, k++, code++) {
                if (l <= FIDJPEG_LOOK) {
                    const int first = code << (FIDJPEG_LOOK - l), n = 1 << (FIDJPEG_LOOK - l);
                    for (int j = 0; j < n && first + j < (1 << FIDJPEG_LOOK); j++) look[first + j] = (uint16_t)((l << 8) | vals[k]);
                }
            }
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        for (int i = 0; i < (1 << FIDJPEG_LOOK); i++) {
            fast_ac[i] = 0;
            const int e = look[i];
            if (!e) continue;
            const int l = e >> 8, rs = e & 255, run = rs >> 4, mag = rs & 15;
            if (mag == 0 || l + mag > FIDJPEG_LOOK) continue;
            int v = ((i << l) & ((1 << FIDJPEG_LOOK) - 1)) >> (FIDJPEG_LOOK - mag);
            if (v < (1 << (mag - 1))) v -= (1 << mag) - 1;
            fast_ac[i] = v * 256 + run * 16 + (l + mag);  // |v| < 512: fits with room to spare
        }
        present = true;
        return true;
    }
};

struct BitReader {
    const uint8_t* p;
    const uint8_t* end;
    uint64_t acc = 0;  // bits left aligned
    int cnt = 0;

    inline void refill() {
        if (p + 8 <= end) {  // fast path: none of the next 8 bytes is 0xFF -> take as many whole bytes as fit
            uint64_t v;
            memcpy(&v, p, 8);
            if (!(((~v) - 0x0101010101010101ull) & v & 0x8080808080808080ull)) {
                const int k = (64 - cnt) >> 3, nb = k * 8;
                const uint64_t be = __builtin_bswap64(v);
                const uint64_t chunk = nb == 64 ? be : (be >> (64 - nb)) << (64 - nb);
                acc |= cnt ? chunk >> cnt : chunk;
                cnt += nb;
                p += k;
                return;
            }
        }
        while (cnt <= 56) {
            uint32_t b = 0;
            if (p < end) {
                b = *p;
                if (b == 0xFF) {
                    if (p + 1 < end && p[1] == 0) {
                        p += 2;  // stuffed zero
                    } else {
                        b = 0;  // a marker: feed zeros, do not run past it
                    }
                } else {
                    p++;
                }
            }
            acc |= (uint64_t)b << (56 - cnt);
            cnt += 8;
        }
    }
    inline void consume(int n) {
        acc <<= n;
        cnt -= n;
    }
    // caller guarantees cnt >= 16 (refill() beforehand)
    inline int decode_nofill(const HuffTable& t) {
        const uint16_t e = t.look[acc >> (64 - FIDJPEG_LOOK)];
        if (e) {
            consume(e >> 8);
            return e & 255;
        }
        for (int l = FIDJPEG_LOOK + 1; l <= 16; l++) {
            const int code = (int)(acc >> (64 - l));
            if (code <= t.maxcode[l]) {
                consume(l);
                return t.vals[(t.valptr[l] + code - t.mincode[l]) & 255];
            }
        }
        return -1;
    }
    inline int decode(const HuffTable& t) {
        if (cnt < 32) refill();
        return decode_nofill(t);
    }
    // caller guarantees cnt >= s
    inline int receive_extend_nofill(int s) {
        int v = (int)(acc >> (64 - s));
        consume(s);
        if (v < (1 << (s - 1))) v -= (1 << s) - 1;
        return v;
    }
    inline int receive_extend(int s) {
        if (cnt < 32) refill();
        return receive_extend_nofill(s);
    }
    // restart: drop the padding bits, step over RSTn
    inline bool restart() {
        acc = 0;
        cnt = 0;
        while (p + 1 < end && !(p[0] == 0xFF && p[1] >= 0xD0 && p[1] <= 0xD7)) {
            if (p[0] == 0xFF && p[1] != 0 && p[1] != 0xFF) return false;  // some other marker
            p++;
        }
        if (p + 1 >= end) return false;
        p += 2;
        return true;
    }
};

static inline int be16(const uint8_t* p) { return (p[0] << 8) | p[1]; }

// Parse the headers up to the start of the entropy-coded data.  *scan = first byte after the SOS header.
static inline int parse_headers(const uint8_t* data, size_t size, FrameInfo* fi, HuffTable dc[4], HuffTable ac[4], int comp_dc[3], int comp_ac[3], int* restart_interval,
                                const uint8_t** scan) {
    if (size < 4 || data[0] != 0xFF || data[1] != 0xD8) return JPEG_BAD;
    uint16_t qt[4][64];
    bool have_q[4] = {false, false, false, false};
    bool have_sof = false;
    int comp_id[3] = {0, 0, 0};
    *restart_interval = 0;
    size_t pos = 2;
    while (pos + 4 <= size) {
        if (data[pos] != 0xFF) return JPEG_BAD;
        while (pos < size && data[pos] == 0xFF) pos++;  // fill bytes
        if (pos >= size) return JPEG_BAD;
        const int m = data[pos++];
        if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
        if (m == 0xD9) return JPEG_BAD;
        if (pos + 2 > size) return JPEG_BAD;
        const int len = be16(data + pos);
        if (len < 2 || pos + len > size) return JPEG_BAD;
        const uint8_t* seg = data + pos + 2;
        const int n = len - 2;
        if (m == 0xDB) {  // DQT
            int o = 0;
            while (o < n) {
                const int pq = seg[o] >> 4, tq = seg[o] & 15;
                o++;
                if (tq > 3 || o + (pq ? 128 : 64) > n) return JPEG_BAD;
                for (int i = 0; i < 64; i++) {
                    qt[tq][i] = pq ? (uint16_t)be16(seg + o + 2 * i) : seg[o + i];
                }
                o += pq ? 128 : 64;
                have_q[tq] = true;
            }
        } else if (m == 0xC0 || m == 0xC1) {  // SOF0 / SOF1
            if (n < 6 || seg[0] != 8) return JPEG_UNSUPPORTED;
            fi->H = be16(seg + 1);
            fi->W = be16(seg + 3);
            fi->ncomp = seg[5];
            if (fi->W <= 0 || fi->H <= 0) return JPEG_UNSUPPORTED;
            if (fi->ncomp != 1 && fi->ncomp != 3) return JPEG_UNSUPPORTED;
            if (n < 6 + 3 * fi->ncomp) return JPEG_BAD;
            for (int c = 0; c < fi->ncomp; c++) {
                comp_id[c] = seg[6 + 3 * c];
                fi->hs[c] = seg[7 + 3 * c] >> 4;
                fi->vs[c] = seg[7 + 3 * c] & 15;
                fi->tq[c] = seg[8 + 3 * c];
                if (fi->tq[c] > 3) return JPEG_BAD;
            }
            have_sof = true;
        } else if (m == 0xC2 || (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC)) {
            return JPEG_UNSUPPORTED;  // progressive, lossless, arithmetic, differential
        } else if (m == 0xC4) {  // DHT
            int o = 0;
            while (o < n) {
                if (o + 17 > n) return JPEG_BAD;
                const int tc = seg[o] >> 4, th = seg[o] & 15;
                if (tc > 1 || th > 3) return JPEG_BAD;
                uint8_t bits[17];
                bits[0] = 0;
                int nsym = 0;
                for (int i = 1; i <= 16; i++) {
                    bits[i] = seg[o + i];
                    nsym += bits[i];
                }
                o += 17;
                if (nsym > 256 || o + nsym > n) return JPEG_BAD;
                if (!(tc ? ac : dc)[th].build(bits, seg + o, nsym)) return JPEG_BAD;
                o += nsym;
            }
        } else if (m == 0xDD) {  // DRI
            if (n < 2) return JPEG_BAD;
            *restart_interval = be16(seg);
        } else if (m == 0xDA) {  // SOS
            if (!have_sof) return JPEG_BAD;
            if (n < 1 || seg[0] != fi->ncomp) return JPEG_UNSUPPORTED;  // multi-scan files
            if (n < 1 + 2 * fi->ncomp + 3) return JPEG_BAD;
            for (int c = 0; c < fi->ncomp; c++) {
                if (seg[1 + 2 * c] != comp_id[c]) return JPEG_UNSUPPORTED;
                comp_dc[c] = seg[2 + 2 * c] >> 4;
                comp_ac[c] = seg[2 + 2 * c] & 15;
                if (comp_dc[c] > 3 || comp_ac[c] > 3 || !dc[comp_dc[c]].present || !ac[comp_ac[c]].present) return JPEG_BAD;
            }
            // geometry
            if (fi->ncomp == 1) {
                fi->hs[0] = fi->vs[0] = 1;
            } else {
                if (fi->hs[1] != 1 || fi->vs[1] != 1 || fi->hs[2] != 1 || fi->vs[2] != 1) return JPEG_UNSUPPORTED;
                const bool ok = (fi->hs[0] == 1 && fi->vs[0] == 1) || (fi->hs[0] == 2 && fi->vs[0] == 1) || (fi->hs[0] == 2 && fi->vs[0] == 2);
                if (!ok) return JPEG_UNSUPPORTED;
            }
            fi->hmax = fi->hs[0];
            fi->vmax = fi->vs[0];
            fi->mcux = (fi->W + 8 * fi->hmax - 1) / (8 * fi->hmax);
            fi->mcuy = (fi->H + 8 * fi->vmax - 1) / (8 * fi->vmax);
            fi->nblk = 0;
            for (int c = 0; c < fi->ncomp; c++) {
                if (!have_q[fi->tq[c]]) return JPEG_BAD;
                memcpy(fi->q[c], qt[fi->tq[c]], sizeof(qt[0]));
                fi->bw[c] = fi->mcux * fi->hs[c];
                fi->bh[c] = fi->mcuy * fi->vs[c];
                fi->blk_base[c] = fi->nblk;
                fi->nblk += fi->bw[c] * fi->bh[c];
                fi->cw[c] = (fi->W * fi->hs[c] + fi->hmax - 1) / fi->hmax;
                fi->ch[c] = (fi->H * fi->vs[c] + fi->vmax - 1) / fi->vmax;
            }
            *scan = data + pos + len;
            return JPEG_OK;
        }
        pos += len;
    }
    return JPEG_BAD;
}

// Entropy-decode one image.  mask[nblk], off[nblk] (offset of the block's first value), vals[cap] (zig-zag order).
static inline int decode_image(const uint8_t* data, size_t size, FrameInfo* fi, uint64_t* mask, uint32_t* off, int16_t* vals, size_t cap, size_t max_blk, size_t* n_vals) {
    HuffTable dc[4], ac[4];
    int cdc[3], cac[3], ri = 0;
    const uint8_t* scan = nullptr;
    const int rc = parse_headers(data, size, fi, dc, ac, cdc, cac, &ri, &scan);
    if (rc != JPEG_OK) return rc;
    if ((size_t)fi->nblk > max_blk || (size_t)fi->nblk * 64 > cap) return JPEG_CAPACITY;
    BitReader br;
    br.p = scan;
    br.end = data + size;
    int pred[3] = {0, 0, 0};
    size_t nv = 0;
    int to_restart = ri;
    for (int my = 0; my < fi->mcuy; my++) {
        for (int mx = 0; mx < fi->mcux; mx++) {
            if (ri) {
                if (to_restart == 0) {
                    if (!br.restart()) return JPEG_BAD;
                    pred[0] = pred[1] = pred[2] = 0;
                    to_restart = ri;
                }
                to_restart--;
            }
            for (int c = 0; c < fi->ncomp; c++) {
                const HuffTable& tdc = dc[cdc[c]];
                const HuffTable& tac = ac[cac[c]];
                for (int v = 0; v < fi->vs[c]; v++) {
                    for (int h = 0; h < fi->hs[c]; h++) {
                        const int b = fi->blk_base[c] + (my * fi->vs[c] + v) * fi->bw[c] + mx * fi->hs[c] + h;
                        uint64_t m = 0;
                        off[b] = (uint32_t)nv;
                        int s = br.decode(tdc);
                        if (s < 0 || s > 11) return JPEG_BAD;
                        if (s) pred[c] += br.receive_extend(s);
                        if (pred[c]) {
                            m |= 1ull;
                            vals[nv++] = (int16_t)pred[c];
                        }
                        for (int k = 1; k < 64;) {
                            if (br.cnt < 32) br.refill();  // one refill per coefficient: code (<= 16) + magnitude (<= 10) bits
                            const int32_t fa = tac.fast_ac[br.acc >> (64 - FIDJPEG_LOOK)];
                            if (fa) {  // code and magnitude inside the look-ahead
                                k += (fa >> 4) & 15;
                                if (k > 63) return JPEG_BAD;
                                br.consume(fa & 15);
                                m |= 1ull << k;
                                vals[nv++] = (int16_t)(fa >> 8);
                                k++;
                                continue;
                            }
                            const int rs = br.decode_nofill(tac);
                            if (rs < 0) return JPEG_BAD;
                            const int r = rs >> 4;
                            s = rs & 15;
                            if (s == 0) {
                                if (r == 15) {
                                    k += 16;
                                    continue;
                                }
                                break;
                            }
                            k += r;
                            if (k > 63) return JPEG_BAD;
                            const int val = br.receive_extend_nofill(s);
                            m |= 1ull << k;
                            vals[nv++] = (int16_t)val;
                            k++;
                        }
                        mask[b] = m;
                    }
                }
            }
        }
    }
    *n_vals = nv;
    return JPEG_OK;
}

}  // namespace fidjpeg
