// CORNER_REFINE_CONTOUR (cv::aruco, OpenCV 4.13 semantics): the reference selects it with doCornerRefinement = true,
// cornerRefinementSubPix = false (aruco_detect/src/aruco_detect.cpp:700-711, 274-281).  The four corners of an identified
// marker are replaced by the intersections of four least-squares lines through the candidate's contour points:
//   * the contour (findContours order) is cut at every point that coincides with one of the four corners: a point belongs to
//     the corner seen last before it, the points before the first corner to the corner seen last overall;
//   * every group gets the line y = a x + b (or x = a y + b when the group is taller than wide), solved through the NORMAL
//     equations in float32 -- the sums are integers, exact in double, and are rounded to float once -- with a pivoted 2x2
//     elimination in float;
//   * corner i = intersection of line i with line i+1 or i-1 (depending on the direction in which the contour visits the
//     corners), 2x2 Cramer in float.
// Sums and extrema are order independent, so the device version (kernels_marker.cuh, k_finish) strides the points over a CTA and
// reduces; this header holds the arithmetic after the reduction and a serial driver for the CPU harness.
#pragma once
#include "common.cuh"
#include "approx_quad.cuh"  // Pt16

namespace fid {

struct LineSums {  // one group of contour points
    long long n, sx, sy, sxx, syy, sxy;
    int minx, maxx, miny, maxy;
};

FID_HD void line_sums_init(LineSums* s) {
    s->n = s->sx = s->sy = s->sxx = s->syy = s->sxy = 0;
    s->minx = s->miny = 0x7fffffff;
    s->maxx = s->maxy = -0x7fffffff;
}
FID_HD void line_sums_add(LineSums* s, int x, int y) {
    s->n++;
    s->sx += x;
    s->sy += y;
    s->sxx += (long long)x * x;
    s->syy += (long long)y * y;
    s->sxy += (long long)x * y;
    s->minx = x < s->minx ? x : s->minx;
    s->maxx = x > s->maxx ? x : s->maxx;
    s->miny = y < s->miny ? y : s->miny;
    s->maxy = y > s->maxy ? y : s->maxy;
}

// cv::solve(A (N x 2: [u, 1]), B (N x 1: v), DECOMP_NORMAL) in float: A^T A = [[suu, su], [su, N]], A^T B = [suv, sv]; 2x2 LU with
// partial pivoting (cv::LU semantics: d = -1/pivot, row update by addition).  Returns false for a singular system.
FID_HD bool solve_normal_2x2(float a00, float a01, float a11, float b0, float b1, float* x0, float* x1) {
    float A[2][2] = {{a00, a01}, {a01, a11}}, B[2] = {b0, b1};
    // column 0
    if (fabsf(A[1][0]) > fabsf(A[0][0])) {
        float t = A[0][0];
        A[0][0] = A[1][0];
        A[1][0] = t;
        t = A[0][1];
        A[0][1] = A[1][1];
        A[1][1] = t;
        t = B[0];
        B[0] = B[1];
        B[1] = t;
    }
    if (fabsf(A[0][0]) < 1.1920928955078125e-06f) return false;  // FLT_EPSILON * 10
    const float d = -1.0f / A[0][0];
    const float alpha = A[1][0] * d;
    A[1][1] += alpha * A[0][1];
    B[1] += alpha * B[0];
    if (fabsf(A[1][1]) < 1.1920928955078125e-06f) return false;
    // back substitution
    B[1] = B[1] / A[1][1];
    B[0] = (B[0] - A[0][1] * B[1]) / A[0][0];
    *x0 = B[0];
    *x1 = B[1];
    return true;
}

// line as (a, b, c): a x + b y + c = 0 in the parametrisation cv::aruco uses
FID_HD void line_from_sums(const LineSums& s, float L[3]) {
    float k = 0.f, m = 0.f;
    if ((float)s.maxx - (float)s.minx > (float)s.maxy - (float)s.miny) {
        solve_normal_2x2((float)(double)s.sxx, (float)(double)s.sx, (float)(double)s.n, (float)(double)s.sxy, (float)(double)s.sy, &k, &m);
        L[0] = k;
        L[1] = -1.f;
        L[2] = m;
    } else {
        solve_normal_2x2((float)(double)s.syy, (float)(double)s.sy, (float)(double)s.n, (float)(double)s.sxy, (float)(double)s.sx, &k, &m);
        L[0] = -1.f;
        L[1] = k;
        L[2] = m;
    }
}

FID_HD void line_cross(const float L1[3], const float L2[3], float* x, float* y) {
    // Matx22f(L1.x, L1.y, L2.x, L2.y).solve(Vec2f(-L1.z, -L2.z)): Cramer in float, zero vector when singular
    const float b0 = -L1[2], b1 = -L2[2];
    float d = L1[0] * L2[1] - L1[1] * L2[0];
    if (d == 0.f) {
        *x = 0.f;
        *y = 0.f;
        return;
    }
    d = 1.f / d;
    *x = (b0 * L2[1] - b1 * L1[1]) * d;
    *y = (b1 * L1[0] - b0 * L2[0]) * d;
}

// corner_index[j] = LAST contour position that coincides with corner j; sums[j] as described above.
FID_HD void corners_from_lines(const LineSums sums[4], const int corner_index[4], float cx[4], float cy[4]) {
    int inc = 1;
    if (corner_index[0] > corner_index[1] && corner_index[3] > corner_index[0]) inc = -1;
    if (corner_index[2] > corner_index[3] && corner_index[1] > corner_index[2]) inc = -1;
    float L[4][3];
    for (int i = 0; i < 4; i++) line_from_sums(sums[i], L[i]);
    for (int i = 0; i < 4; i++) {
        if (inc < 0)
            line_cross(L[i], L[(i + 1) & 3], &cx[i], &cy[i]);
        else
            line_cross(L[i], L[(i + 3) & 3], &cx[i], &cy[i]);
    }
}

// serial driver (CPU harness): pts = the candidate's contour in findContours order; cx, cy = its corners (integers as float)
FID_HD bool refine_candidate_lines_serial(const Pt16* pts, int n, float cx[4], float cy[4]) {
    LineSums sums[5];
    for (int g = 0; g < 5; g++) line_sums_init(&sums[g]);
    int corner_index[4] = {-1, -1, -1, -1};
    int group = 4;
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < 4; j++)
            if ((float)pts[i].x == cx[j] && (float)pts[i].y == cy[j]) {
                corner_index[j] = i;
                group = j;
            }
        line_sums_add(&sums[group], pts[i].x, pts[i].y);
    }
    for (int j = 0; j < 4; j++)
        if (corner_index[j] < 0) return false;
    if (sums[4].n) {  // the points before the first corner belong to the corner seen last
        LineSums& d = sums[group];
        const LineSums& e = sums[4];
        d.n += e.n;
        d.sx += e.sx;
        d.sy += e.sy;
        d.sxx += e.sxx;
        d.syy += e.syy;
        d.sxy += e.sxy;
        d.minx = e.minx < d.minx ? e.minx : d.minx;
        d.maxx = e.maxx > d.maxx ? e.maxx : d.maxx;
        d.miny = e.miny < d.miny ? e.miny : d.miny;
        d.maxy = e.maxy > d.maxy ? e.maxy : d.maxy;
    }
    corners_from_lines(sums, corner_index, cx, cy);
    return true;
}

}  // namespace fid
