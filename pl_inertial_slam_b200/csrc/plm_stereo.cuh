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

} // namespace plm
