// cv::cornerSubPix for one corner, as run by detectMarkers' CORNER_REFINE_SUBPIX
// step that the reference enables at aruco_detect/src/aruco_detect.cpp:700-711 (criteria :694-696).
// float32 / float64 mix and accumulation order follow OpenCV so that the iteration count -- and
// with it the converged point -- is reproduced, not just approximated.
#pragma once
#include "common.cuh"

namespace fid {

#define FID_SUBPIX_MAX_WIN 5  // cornerRefinementWinSize upper bound supported (reference uses 5)

// getRectSubPix(8u -> 32f) of a (2*win+3)^2 patch centred at (cx,cy); replicate border.
template <class Img>
FID_HD void rect_subpix(const Img& gray, int W, int H, float cx, float cy, int win, float* patch) {
    const int pw = 2 * win + 3;
    cx -= (pw - 1) * 0.5f;
    cy -= (pw - 1) * 0.5f;
    const int ix = (int)floorf(cx), iy = (int)floorf(cy);
    const float a = cx - ix, b = cy - iy;
    const float a11 = (1.f - a) * (1.f - b), a12 = a * (1.f - b), a21 = (1.f - a) * b, a22 = a * b;
    const bool inside = 0 <= ix && ix < W - pw && 0 <= iy && iy < H - pw;
    for (int i = 0; i < pw; i++) {
        int y0 = iy + i, y1 = iy + i + 1;
        if (!inside) {
            y0 = y0 < 0 ? 0 : (y0 > H - 1 ? H - 1 : y0);
            y1 = y1 < 0 ? 0 : (y1 > H - 1 ? H - 1 : y1);
        }
        for (int j = 0; j < pw; j++) {
            int x0 = ix + j, x1 = ix + j + 1;
            if (!inside) {
                x0 = x0 < 0 ? 0 : (x0 > W - 1 ? W - 1 : x0);
                x1 = x1 < 0 ? 0 : (x1 > W - 1 ? W - 1 : x1);
            }
            patch[i * pw + j] = gray.at(x0, y0) * a11 + gray.at(x1, y0) * a12 + gray.at(x0, y1) * a21 + gray.at(x1, y1) * a22;
        }
    }
}

// mask[(2win+1)^2] is the separable exp window computed on the host (libm expf, as OpenCV does).
template <class Img>
FID_HD void corner_subpix(const Img& gray, int W, int H, float* px, float* py, int win, const float* mask, int max_iters, double eps_sq,
                          float* patch /* (2win+3)^2 scratch */) {
    const int ww = 2 * win + 1, pw = ww + 2;
    const float tx = *px, ty = *py;
    float cx = tx, cy = ty;
    int iter = 0;
    double err = 0.0;
    do {
        rect_subpix(gray, W, H, cx, cy, win, patch);
        double a = 0, b = 0, c = 0, bb1 = 0, bb2 = 0;
        for (int i = 0, k = 0; i < ww; i++) {
            const float* sp = patch + (i + 1) * pw + 1;
            const double py_ = i - win;
            for (int j = 0; j < ww; j++, k++) {
                const double m = mask[k];
                const double tgx = sp[j + 1] - sp[j - 1];
                const double tgy = sp[j + pw] - sp[j - pw];
                const double gxx = tgx * tgx * m;
                const double gxy = tgx * tgy * m;
                const double gyy = tgy * tgy * m;
                const double px_ = j - win;
                a += gxx;
                b += gxy;
                c += gyy;
                bb1 += gxx * px_ + gxy * py_;
                bb2 += gxy * px_ + gyy * py_;
            }
        }
        const double det = a * c - b * b;
        if (fabs(det) <= 2.220446049250313e-16 * 2.220446049250313e-16) break;
        const double scale = 1.0 / det;
        const float nx = (float)(cx + c * scale * bb1 - b * scale * bb2);
        const float ny = (float)(cy - b * scale * bb1 + a * scale * bb2);
        err = (nx - cx) * (nx - cx) + (ny - cy) * (ny - cy);
        cx = nx;
        cy = ny;
        if (cx < 0 || cx >= W || cy < 0 || cy >= H) break;
    } while (++iter < max_iters && err > eps_sq);
    if (fabsf(cx - tx) > win || fabsf(cy - ty) > win) {
        cx = tx;
        cy = ty;
    }
    *px = cx;
    *py = cy;
}

}  // namespace fid
