// Per-match geometry gates of the stereo drivers, evaluated on the device so the keep-mask comes
// back with the match vector (K4 in SURVEY.md 2.1).  The reference is built for baseline x86-64
// without FMA (CMakeLists.txt:41), so every product feeding a comparison uses explicit
// round-to-nearest intrinsics that the compiler never contracts.
#pragma once
#include "plm_common.cuh"

namespace plm {

// std::min / std::max semantics on doubles (NaN behaviour differs from fmin/fmax).
__device__ __forceinline__ double std_min(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double std_max(double a, double b) { return (a < b) ? b : a; }

// StereoFrame::matchStereoPoints gates (stvo-pl/src/stereoFrame.cpp:168-171): float subtraction,
// |dy| <= maxDistEpip and disparity >= minDisp compared in double.
__device__ __forceinline__ bool stereo_point_gate(float2 l, float2 r, double max_dist_epip, double min_disp, double &disp) {
    const float dy = __fsub_rn(l.y, r.y);
    if (static_cast<double>(fabsf(dy)) <= max_dist_epip) {
        const double d = static_cast<double>(__fsub_rn(l.x, r.x));
        if (d >= min_disp) {
            disp = d;
            return true;
        }
    }
    return false;
}

__global__ void stereo_filter_points_kernel(const float2 *__restrict__ kp_l, const float2 *__restrict__ kp_r, int n2,
                                            const int32_t *__restrict__ m12, int n1, double max_dist_epip,
                                            double min_disp, uint8_t *__restrict__ keep, double *__restrict__ disp,
                                            int32_t *__restrict__ count) {
    const int i1 = blockIdx.x * blockDim.x + threadIdx.x;
    bool kept = false;
    if (i1 < n1) {
        double dsp = 0.0;
        const int i2 = m12[i1];
        if (i2 >= 0 && i2 < n2) kept = stereo_point_gate(kp_l[i1], kp_r[i2], max_dist_epip, min_disp, dsp);
        keep[i1] = kept ? 1 : 0;
        disp[i1] = dsp;
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, kept);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, __popc(m));
}

// StereoFrame::lineSegmentOverlapStereo (stereoFrame.cpp:484-519).
__device__ __forceinline__ double line_overlap_stereo(double spl_obs, double epl_obs, double spl_proj, double epl_proj,
                                                      double line_horiz_th) {
    double overlap = 1.0;
    if (fabs(__dsub_rn(epl_obs, spl_obs)) > line_horiz_th) {
        const double sln = std_min(spl_obs, epl_obs), eln = std_max(spl_obs, epl_obs);
        const double spn = std_min(spl_proj, epl_proj), epn = std_max(spl_proj, epl_proj);
        const double length = __dsub_rn(eln, spn);
        if ((epn < sln) || (spn > eln))
            overlap = 0.0;
        else if ((epn > eln) && (spn < sln))
            overlap = __dsub_rn(eln, sln);
        else
            overlap = __dsub_rn(std_min(eln, epn), std_max(sln, spn));
        if (length > static_cast<double>(0.01f))
            overlap = __ddiv_rn(overlap, length);
        else
            overlap = 0.0;
        if (overlap > 1.0) overlap = 1.0;
    }
    return overlap;
}

// StereoFrame::matchStereoLines gates (stereoFrame.cpp:366-385) with filterLineSegmentDisparity
// (:416-426).  sp_r is overwritten before ep_r is interpolated, as in the reference (:377-378).
// disp_s / disp_e come back as computed (including the -1 / -1 rejection marker).
__device__ __forceinline__ bool stereo_line_gate(float4 l, float4 r, double min_disp, double line_horiz_th,
                                                 double stereo_overlap_th, double ls_min_disp_ratio, double &disp_s,
                                                 double &disp_e) {
    const double sp_l0 = l.x, sp_l1 = l.y, ep_l0 = l.z, ep_l1 = l.w;
    double sp_r0 = r.x, sp_r1 = r.y, ep_r0 = r.z, ep_r1 = r.w;
    const double overlap = line_overlap_stereo(sp_l1, ep_l1, sp_r1, ep_r1, line_horiz_th);
    {
        const double a = __dmul_rn(sp_r0, __dsub_rn(sp_l1, ep_r1));
        const double b = __dmul_rn(ep_r0, __dsub_rn(sp_r1, sp_l1));
        sp_r0 = __ddiv_rn(__dadd_rn(a, b), __dsub_rn(sp_r1, ep_r1));
        sp_r1 = sp_l1;
    }
    {
        const double a = __dmul_rn(sp_r0, __dsub_rn(ep_l1, ep_r1));
        const double b = __dmul_rn(ep_r0, __dsub_rn(sp_r1, ep_l1));
        ep_r0 = __ddiv_rn(__dadd_rn(a, b), __dsub_rn(sp_r1, ep_r1));
        ep_r1 = ep_l1;
    }
    disp_s = __dsub_rn(sp_l0, sp_r0);
    disp_e = __dsub_rn(ep_l0, ep_r0);
    if (__ddiv_rn(std_min(disp_s, disp_e), std_max(disp_s, disp_e)) < ls_min_disp_ratio) {
        disp_s = -1.0;
        disp_e = -1.0;
    }
    return disp_s >= min_disp && disp_e >= min_disp && fabs(__dsub_rn(sp_l1, ep_l1)) > line_horiz_th &&
           fabs(__dsub_rn(sp_r1, ep_r1)) > line_horiz_th && overlap > stereo_overlap_th;
}

__global__ void stereo_filter_lines_kernel(const float4 *__restrict__ ln_l, const float4 *__restrict__ ln_r, int n2,
                                           const int32_t *__restrict__ m12, int n1, double min_disp,
                                           double line_horiz_th, double stereo_overlap_th, double ls_min_disp_ratio,
                                           uint8_t *__restrict__ keep, double *__restrict__ disp_se,
                                           int32_t *__restrict__ count) {
    const int i1 = blockIdx.x * blockDim.x + threadIdx.x;
    bool kept = false;
    if (i1 < n1) {
        double disp_s = 0.0, disp_e = 0.0;
        const int i2 = m12[i1];
        if (i2 >= 0 && i2 < n2)
            kept = stereo_line_gate(ln_l[i1], ln_r[i2], min_disp, line_horiz_th, stereo_overlap_th, ls_min_disp_ratio,
                                    disp_s, disp_e);
        keep[i1] = kept ? 1 : 0;
        disp_se[2 * i1] = disp_s;
        disp_se[2 * i1 + 1] = disp_e;
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, kept);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, __popc(m));
}

// ---- opt-in geometric filter for matched line pairs (BASELINE config 2; SURVEY 8 note 6) ---------------------------

__device__ __forceinline__ double overlap_from_lambdas(double lambda_s, double lambda_e) {
    const double lambda_min = std_min(lambda_s, lambda_e);
    const double lambda_max = std_max(lambda_s, lambda_e);
    if (lambda_min < 0.0 && lambda_max > 1.0) return 1.0;
    if (lambda_max < 0.0 || lambda_min > 1.0) return 0.0;
    if (lambda_min < 0.0) return lambda_max;
    if (lambda_max > 1.0) return __dsub_rn(1.0, lambda_min);
    return __dsub_rn(lambda_max, lambda_min);
}

// StereoFrame::lineSegmentOverlap (stvo-pl/src/stereoFrame.cpp:521-627): fraction of the observed segment (so, eo)
// covered by the other segment (sp, ep) projected onto its line; every product / sum rounded separately (the
// reference is built without FMA contraction).
__device__ __forceinline__ double line_segment_overlap(double sox, double soy, double eox, double eoy, double spx, double spy,
                                                       double epx, double epy) {
    const double l0 = __dsub_rn(eox, sox), l1 = __dsub_rn(eoy, soy);
    if (fabs(__dsub_rn(sox, eox)) < 1.0) // vertical lines
        return overlap_from_lambdas(__ddiv_rn(__dsub_rn(spy, soy), l1), __ddiv_rn(__dsub_rn(epy, soy), l1));
    if (fabs(__dsub_rn(soy, eoy)) < 1.0) // horizontal lines
        return overlap_from_lambdas(__ddiv_rn(__dsub_rn(spx, sox), l0), __ddiv_rn(__dsub_rn(epx, sox), l0));
    const double a = __dsub_rn(soy, eoy), b = __dsub_rn(eox, sox);
    const double c = __dsub_rn(__dmul_rn(sox, eoy), __dmul_rn(eox, soy));
    const double lxy = __ddiv_rn(1.0, __dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)));
    const double ac = __dmul_rn(a, c);
    const double sx = __dmul_rn(__dsub_rn(__dmul_rn(b, __dsub_rn(__dmul_rn(b, spx), __dmul_rn(a, spy))), ac), lxy);
    const double ex = __dmul_rn(__dsub_rn(__dmul_rn(b, __dsub_rn(__dmul_rn(b, epx), __dmul_rn(a, epy))), ac), lxy);
    return overlap_from_lambdas(__ddiv_rn(__dsub_rn(sx, sox), l0), __ddiv_rn(__dsub_rn(ex, sox), l0));
}

// keep[i1] = matched && overlap > overlap_th && !(|cos| < line_sim_th)   (a NaN similarity passes, matching.cpp:221)
__global__ void line_pair_filter_kernel(const float4 *__restrict__ ln1, int n1, const float4 *__restrict__ ln2, int n2,
                                        const int32_t *__restrict__ m12, double overlap_th, double line_sim_th,
                                        uint8_t *__restrict__ keep, double *__restrict__ overlap, double *__restrict__ sim,
                                        int32_t *__restrict__ count) {
    const int i1 = blockIdx.x * blockDim.x + threadIdx.x;
    bool kept = false;
    if (i1 < n1) {
        double ov = 0.0, s = 0.0;
        const int i2 = m12[i1];
        if (i2 >= 0 && i2 < n2) {
            const float4 o = ln1[i1], p = ln2[i2];
            ov = line_segment_overlap(o.x, o.y, o.z, o.w, p.x, p.y, p.z, p.w);
            // dot / normalize of stvo-pl/include/matching.h:39-48
            double vx = __dsub_rn(double(o.z), double(o.x)), vy = __dsub_rn(double(o.w), double(o.y));
            double wx = __dsub_rn(double(p.z), double(p.x)), wy = __dsub_rn(double(p.w), double(p.y));
            const double mv = __dsqrt_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)));
            const double mw = __dsqrt_rn(__dadd_rn(__dmul_rn(wx, wx), __dmul_rn(wy, wy)));
            vx = __ddiv_rn(vx, mv);
            vy = __ddiv_rn(vy, mv);
            wx = __ddiv_rn(wx, mw);
            wy = __ddiv_rn(wy, mw);
            s = fabs(__dadd_rn(__dmul_rn(vx, wx), __dmul_rn(vy, wy)));
            kept = ov > overlap_th && !(s < line_sim_th);
        }
        keep[i1] = kept ? 1 : 0;
        overlap[i1] = ov;
        sim[i1] = s;
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, kept);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, __popc(m));
}

} // namespace plm
