// Host-side parameter handling shared by the C-ABI (fid_api.cu) and tests/hostsim: fid_params
// (include/fiducials_b200.h, mirrors aruco_detect.cpp:690-727) -> fid::DevParams, dictionary
// packing and the cornerSubPix window table.
#pragma once
#include <math.h>
#include <string.h>
#include <vector>

#include "../../include/fiducials_b200.h"
#include "common.cuh"
#include "dict_tables.h"

namespace fid {

inline void default_params(fid_params* p) {
    memset(p, 0, sizeof(*p));
    p->dictionary = 7;  // aruco_detect.cpp:611
    p->adaptiveThreshConstant = 7;
    p->adaptiveThreshWinSizeMax = 53;
    p->adaptiveThreshWinSizeMin = 3;
    p->adaptiveThreshWinSizeStep = 4;
    p->cornerRefinementMaxIterations = 30;
    p->cornerRefinementMinAccuracy = 0.01;
    p->cornerRefinementWinSize = 5;
    p->cornerRefinementMethod = 1;
    p->errorCorrectionRate = 0.6;
    p->minCornerDistanceRate = 0.05;
    p->markerBorderBits = 1;
    p->maxErroneousBitsInBorderRate = 0.04;
    p->minDistanceToBorder = 3;
    p->minMarkerDistanceRate = 0.05;
    p->minMarkerPerimeterRate = 0.1;
    p->maxMarkerPerimeterRate = 4.0;
    p->minOtsuStdDev = 5.0;
    p->perspectiveRemoveIgnoredMarginPerCell = 0.13;
    p->perspectiveRemovePixelPerCell = 8;
    p->polygonalApproxAccuracyRate = 0.01;
    p->relativeCornerRefinmentWinSize = 100.0;
    p->minGroupDistance = 0.21;
}


// Returns FID_OK or an error status; fills dp.
inline int make_dev_params(const fid_params& p, DevParams* dp) {
    memset(dp, 0, sizeof(*dp));
    const FidDictInfo* info = nullptr;
    for (int i = 0; i < kNumDictInfo; i++)
        if (kDictInfo[i].id == p.dictionary) info = &kDictInfo[i];
    if (!info) return FID_ERR_UNSUPPORTED;
    if (p.adaptiveThreshWinSizeMin < 3 || p.adaptiveThreshWinSizeMax < p.adaptiveThreshWinSizeMin || p.adaptiveThreshWinSizeStep <= 0) return FID_ERR_INVALID_ARG;
    const int n_scales = (p.adaptiveThreshWinSizeMax - p.adaptiveThreshWinSizeMin) / p.adaptiveThreshWinSizeStep + 1;
    if (n_scales > FID_MAX_SCALES) return FID_ERR_UNSUPPORTED;
    dp->n_scales = n_scales;
    for (int i = 0; i < n_scales; i++) {
        int w = p.adaptiveThreshWinSizeMin + i * p.adaptiveThreshWinSizeStep;
        if (w % 2 == 0) w++;
        if (w / 2 > FID_MAX_WIN_RADIUS) return FID_ERR_UNSUPPORTED;
        dp->win[i] = w;
    }
    dp->thresh_c = (int)floor(p.adaptiveThreshConstant);
    dp->min_perimeter_rate = p.minMarkerPerimeterRate;
    dp->max_perimeter_rate = p.maxMarkerPerimeterRate;
    dp->poly_accuracy_rate = p.polygonalApproxAccuracyRate;
    dp->min_corner_dist_rate = p.minCornerDistanceRate;
    dp->min_dist_to_border = p.minDistanceToBorder;
    dp->min_marker_dist_rate = p.minMarkerDistanceRate;
    dp->marker_border_bits = p.markerBorderBits;
    dp->px_per_cell = p.perspectiveRemovePixelPerCell;
    dp->ignored_margin_per_cell = p.perspectiveRemoveIgnoredMarginPerCell;
    dp->max_err_border_rate = p.maxErroneousBitsInBorderRate;
    dp->min_otsu_stddev = p.minOtsuStdDev;
    dp->error_correction_rate = p.errorCorrectionRate;
    if (p.cornerRefinementMethod < 0 || p.cornerRefinementMethod > 2) return FID_ERR_UNSUPPORTED;  // NONE, SUBPIX, CONTOUR (APRILTAG: not in the reference)
    dp->corner_refine = p.cornerRefinementMethod;
    dp->refine_win = p.cornerRefinementWinSize;
    dp->refine_max_iter = p.cornerRefinementMaxIterations;
    dp->refine_min_acc = p.cornerRefinementMinAccuracy;
    dp->rel_refine_win = p.relativeCornerRefinmentWinSize;
    dp->min_group_dist = p.minGroupDistance;
    dp->marker_size = info->marker_size;
    dp->n_markers = info->n_markers;
    dp->max_correction_bits = info->max_correction_bits;
    dp->dict_nbytes = (info->marker_size * info->marker_size + 7) / 8;
    dp->dict_table = info->table;
    if (p.markerBorderBits < 1 || p.perspectiveRemovePixelPerCell < 1) return FID_ERR_INVALID_ARG;
    const int cells = dp->marker_size + 2 * dp->marker_border_bits;
    if (cells * dp->px_per_cell > FID_MAX_WARP_SIDE || cells * cells > 128 || dp->marker_size * dp->marker_size > 64) return FID_ERR_UNSUPPORTED;
    if (dp->corner_refine == 1 && (dp->refine_win < 1 || dp->refine_win > 5 || dp->refine_max_iter < 1 || !(dp->refine_min_acc > 0))) return FID_ERR_UNSUPPORTED;
    return FID_OK;
}

// Dictionary as n_markers x 4 rotations of 64-bit words (byte k of the rotation in bits 8k..8k+7).
inline void pack_dictionary(const DevParams& dp, std::vector<unsigned long long>* out) {
    const uint8_t* tab = kDictTables[dp.dict_table];
    const int nb = dp.dict_nbytes;
    out->assign((size_t)dp.n_markers * 4, 0ull);
    for (int m = 0; m < dp.n_markers; m++)
        for (int r = 0; r < 4; r++) {
            unsigned long long v = 0;
            for (int k = 0; k < nb; k++) v |= (unsigned long long)tab[(size_t)m * 4 * nb + r * nb + k] << (8 * k);
            (*out)[(size_t)m * 4 + r] = v;
        }
}

// cornerSubPix weighting window for half-size `win`, computed with the host libm exactly like
// OpenCV does (float y; float vy = expf(-y*y); mask = (float)(vy * expf(-x*x))).
inline void subpix_mask(int win, float* mask /* (2win+1)^2 */) {
    const int ww = 2 * win + 1;
    for (int i = 0; i < ww; i++) {
        const float y = (float)(i - win) / win;
        const float vy = expf(-y * y);
        for (int j = 0; j < ww; j++) {
            const float x = (float)(j - win) / win;
            mask[i * ww + j] = (float)(vy * expf(-x * x));
        }
    }
}

}  // namespace fid
