// Polygonal approximation of one closed contour + the quad-candidate filters of
// cv::aruco::_findMarkerContours (reached from aruco_detect/src/aruco_detect.cpp:350).
// Algorithm restated from SURVEY.md A.4 / E.2 (cv::approxPolyDP, closed curve, integer points):
//   1. three "farthest point from the current start" sweeps pick the initial split,
//   2. Douglas-Peucker with an explicit slice stack (left slice first),
//   3. a clean-up pass that drops nearly-collinear vertices.
// Every sweep is an arg-max over a cyclic range of the contour, first maximum wins.  The sweeps
// are delegated to a Reducer so that a CUDA block can run them as parallel reductions while the
// (uniform) control flow stays identical on every thread; tests/hostsim plugs a serial reducer.
//
// Only the "is it a 4-gon" outcome matters to the detector, and the clean-up pass removes at most
// half of the vertices, so the walk aborts as soon as more than 8 vertices are certain.
#pragma once
#include "common.cuh"

namespace fid {

struct ArgMax {
    int value;  // maximum (>= 0)
    int index;  // cyclic offset j in [1, len] from the sweep origin at which it was first reached
};

// Serial reducer (host tests; also usable by a single device thread).
struct SerialReducer {
    // max over j=1..len-1 of |P[(pos0+j)%n] - P[pos0]|^2
    FID_HD ArgMax farthest(const Pt16* p, int n, int pos0, int len) const {
        ArgMax r = {0, 0};
        const int sx = p[pos0].x, sy = p[pos0].y;
        int pos = pos0;
        for (int j = 1; j < len; j++) {
            pos = pos + 1 == n ? 0 : pos + 1;
            const int dx = p[pos].x - sx, dy = p[pos].y - sy;
            const int d = dx * dx + dy * dy;
            if (d > r.value) {
                r.value = d;
                r.index = j;
            }
        }
        return r;
    }
    // max over the open cyclic range (s0, s1) of |(P.y-S.y)*dx - (P.x-S.x)*dy|, S = P[s0], (dx,dy) = P[s1]-P[s0]
    FID_HD ArgMax off_chord(const Pt16* p, int n, int s0, int s1) const {
        ArgMax r = {0, 0};
        const int sx = p[s0].x, sy = p[s0].y;
        const int dx = p[s1].x - sx, dy = p[s1].y - sy;
        int pos = s0 + 1 == n ? 0 : s0 + 1;
        int j = 1;
        while (pos != s1) {
            int d = (p[pos].y - sy) * dx - (p[pos].x - sx) * dy;
            d = d < 0 ? -d : d;
            if (d > r.value) {
                r.value = d;
                r.index = j;
            }
            pos = pos + 1 == n ? 0 : pos + 1;
            j++;
        }
        return r;
    }
};

#define FID_APPROX_MAX_V 8

// Returns the vertex count after clean-up (vertices in out[]), or -1 when the polygon certainly has
// more than FID_APPROX_MAX_V vertices before clean-up (=> cannot end up a 4-gon).
template <class Reducer>
FID_HD int approx_poly_closed(const Reducer& red, const Pt16* p, int n, double eps, Pt16* out) {
    if (n <= 0) return 0;
    const double eps2 = eps * eps;
    int cnt = 0;
    // -- 1. initial split
    int pos = 0, right_start = 0;
    bool le_eps = false;
    for (int it = 0; it < 3; it++) {
        pos = (pos + right_start) % n;
        ArgMax am = red.farthest(p, n, pos, n);
        right_start = am.index;
        le_eps = (double)am.value <= eps2;
    }
    int stack_a[FID_APPROX_MAX_V + 4], stack_b[FID_APPROX_MAX_V + 4];
    int sp = 0;
    if (!le_eps) {
        const int left_start = pos;
        const int split = (right_start + left_start) % n;
        stack_a[sp] = split;  // right slice, popped second
        stack_b[sp++] = left_start;
        stack_a[sp] = left_start;  // left slice, popped first
        stack_b[sp++] = split;
    } else {
        out[cnt++] = p[pos];
    }
    // -- 2. Douglas-Peucker
    while (sp > 0) {
        sp--;
        const int s0 = stack_a[sp], s1 = stack_b[sp];
        bool le = true;
        int split = 0;
        const int next = s0 + 1 == n ? 0 : s0 + 1;
        if (next != s1) {
            ArgMax am = red.off_chord(p, n, s0, s1);
            const int dx = p[s1].x - p[s0].x, dy = p[s1].y - p[s0].y;
            le = (double)am.value * (double)am.value <= eps2 * (double)(dx * dx + dy * dy);
            split = (s0 + am.index) % n;
        }
        if (le) {
            if (cnt >= FID_APPROX_MAX_V) return -1;
            out[cnt++] = p[s0];
        } else {
            if (cnt + sp + 2 > FID_APPROX_MAX_V) return -1;  // every pending slice yields >= 1 vertex
            stack_a[sp] = split;
            stack_b[sp++] = s1;
            stack_a[sp] = s0;
            stack_b[sp++] = split;
        }
    }
    // -- 3. clean-up (serial, <= 8 vertices)
    const int count = cnt;
    int new_count = count;
    int rp = count - 1;
    Pt16 start_pt = out[rp];
    rp = rp + 1 == count ? 0 : rp + 1;
    int wp = rp;
    Pt16 pt = out[rp];
    rp = rp + 1 == count ? 0 : rp + 1;
    for (int i = 0; i < count && new_count > 2; i++) {
        const Pt16 end_pt = out[rp];
        rp = rp + 1 == count ? 0 : rp + 1;
        const int dx = end_pt.x - start_pt.x, dy = end_pt.y - start_pt.y;
        int dist = (pt.x - start_pt.x) * dy - (pt.y - start_pt.y) * dx;
        dist = dist < 0 ? -dist : dist;
        const int successive_inner_product = (pt.x - start_pt.x) * (end_pt.x - pt.x) + (pt.y - start_pt.y) * (end_pt.y - pt.y);
        if ((double)dist * (double)dist <= 0.5 * eps2 * (double)(dx * dx + dy * dy) && dx != 0 && dy != 0 && successive_inner_product >= 0) {
            new_count--;
            out[wp] = start_pt = end_pt;
            wp = wp + 1 == count ? 0 : wp + 1;
            pt = out[rp];
            rp = rp + 1 == count ? 0 : rp + 1;
            i++;
            continue;
        }
        out[wp] = start_pt = pt;
        wp = wp + 1 == count ? 0 : wp + 1;
        pt = end_pt;
    }
    return new_count;
}

// cv::isContourConvex for an integer polygon (sign-consistency of consecutive edge cross products).
FID_HD bool is_convex_int(const Pt16* q, int n) {
    Pt16 prev = q[(n - 2 + n) % n], cur = q[n - 1];
    int dx0 = cur.x - prev.x, dy0 = cur.y - prev.y;
    int orientation = 0;
    for (int i = 0; i < n; i++) {
        prev = cur;
        cur = q[i];
        const int dx = cur.x - prev.x, dy = cur.y - prev.y;
        const int dxdy0 = dx * dy0, dydx0 = dy * dx0;
        orientation |= (dydx0 > dxdy0) ? 1 : ((dydx0 < dxdy0) ? 2 : 3);
        if (orientation == 3) return false;
        dx0 = dx;
        dy0 = dy;
    }
    return true;
}

// The remaining filters of _findMarkerContours applied to a 4-vertex approximation.
// NOTE (found by black-box probing of cv2 4.13, DESIGN.md "border rule"): the minDistanceToBorder
// test is NOT applied here.  In OpenCV 4.13 a quad touching the border still takes part in the
// grouping step (it can become a group leader and un-select its neighbours) and only then is the
// selected candidate discarded -- silently, together with its group.  See quad_near_border().
FID_HD bool quad_passes_filters(const Pt16* q, int n_contour, int W, int H, double min_corner_dist_rate) {
    if (!is_convex_int(q, 4)) return false;
    const int mx = W > H ? W : H;
    double min_d = (double)mx * (double)mx;
    for (int j = 0; j < 4; j++) {
        const int dx = q[j].x - q[(j + 1) & 3].x, dy = q[j].y - q[(j + 1) & 3].y;
        const double d = (double)(dx * dx + dy * dy);
        min_d = d < min_d ? d : min_d;
    }
    const double min_corner_px = (double)n_contour * min_corner_dist_rate;
    if (min_d < min_corner_px * min_corner_px) return false;
    return true;
}

}  // namespace fid
