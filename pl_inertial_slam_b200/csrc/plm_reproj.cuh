// Selection of the local map and the reprojection gates around the map -> keyframe matcher
// (MapHandler::matchMap2KFPoints / matchMap2KFLines, src/mapHandler.cpp:583-682 and :685-803), on the device so
// that a resident local map never travels to the host between the projection, the matcher and the gate.
//
//   select   Pf = R X + t (Twf = [R | t], :600 / :701-703), pf = cam->projection(Pf) = (cx + fx Pf0 / Pf2,
//            cy + fy Pf1 / Pf2) (stvo-pl/src/pinholeStereoCamera.cpp:239-245); a landmark is kept when it is active
//            (local and not yet observed from this keyframe -- a host-side flag) and pf lies strictly inside the
//            image with positive depth (:602 / :705-706).  Kept landmarks are compacted IN ORDER (the matcher is
//            order dependent) together with their grid-cell coordinates: the reference stores pf * inv_width /
//            inv_height into std::pair<int,int> (:605, :709-710), i.e. a double -> int truncation.
//   gate     points: |pf_map - pl_kf| < maxKFEpipP (:661-662); lines: l . (spf, 1) < maxKFEpipL and l . (epf, 1) <
//            maxKFEpipL, signed, exactly as written (:784-786).  A failing pair is dropped and decrements the count.
//
// Arithmetic: fp64, every product and sum rounded separately (no FMA: the reference is built for baseline x86-64).
// The 3x3 * 3x1 product is evaluated left to right, ((r0 x0 + r1 x1) + r2 x2) + t.  Eigen's own evaluation order for
// this fixed-size product is not reproducible without Eigen (absent from this image), so the fp64 outputs are
// "parity unpinned" (see DESIGN.md 5); selections, cell coordinates and gate decisions are integer / boolean results.
#pragma once
#include "plm_common.cuh"

namespace plm {

struct MapView {
    double T[12];  // Twf rows 0..2, row-major: r00 r01 r02 tx | r10 r11 r12 ty | r20 r21 r22 tz
    double fx, fy, cx, cy;
    double inv_width, inv_height; // GRID_COLS / image width, GRID_ROWS / image height
    int32_t width, height;
};

__device__ __forceinline__ double3 view_transform(const MapView &v, double x, double y, double z) {
    double3 p;
    p.x = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(v.T[0], x), __dmul_rn(v.T[1], y)), __dmul_rn(v.T[2], z)), v.T[3]);
    p.y = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(v.T[4], x), __dmul_rn(v.T[5], y)), __dmul_rn(v.T[6], z)), v.T[7]);
    p.z = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(v.T[8], x), __dmul_rn(v.T[9], y)), __dmul_rn(v.T[10], z)), v.T[11]);
    return p;
}
__device__ __forceinline__ double2 view_project(const MapView &v, double3 P) {
    double2 uv;
    uv.x = __dadd_rn(v.cx, __ddiv_rn(__dmul_rn(v.fx, P.x), P.z));
    uv.y = __dadd_rn(v.cy, __ddiv_rn(__dmul_rn(v.fy, P.y), P.z));
    return uv;
}
__device__ __forceinline__ bool view_inside(const MapView &v, double2 uv, double3 P) {
    return uv.x > 0 && uv.x < static_cast<double>(v.width) && uv.y > 0 && uv.y < static_cast<double>(v.height) && P.z > 0.0;
}
// double -> int as the C++ conversion the reference relies on (truncation; values are inside the image here)
__device__ __forceinline__ int32_t cell_of(double px, double inv) { return __double2int_rz(__dmul_rn(px, inv)); }

// pass 0: keep flags + per-CTA counts; pass 1 (after an exclusive scan of the counts): ordered compaction.
// PTS = 1: landmarks are points (3 doubles), PTS = 2: segments (6 doubles, both endpoints must be inside).
template <int PTS>
__global__ void __launch_bounds__(256) map_select_kernel(const double *__restrict__ X, const uint8_t *__restrict__ active, int n, MapView v,
                                                         int pass, int32_t *__restrict__ cta_count, int32_t *__restrict__ sel,
                                                         int32_t *__restrict__ coords, double *__restrict__ pf, int32_t *__restrict__ n_sel) {
    __shared__ int s_warp[8];
    const int i = blockIdx.x * 256 + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    bool keep = false;
    double2 uv[PTS];
    if (i < n && (!active || active[i])) {
        keep = true;
#pragma unroll
        for (int k = 0; k < PTS; ++k) {
            const double *p = X + static_cast<size_t>(i) * (3 * PTS) + 3 * k;
            const double3 P = view_transform(v, p[0], p[1], p[2]);
            uv[k] = view_project(v, P);
            keep = keep && view_inside(v, uv[k], P);
        }
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, keep);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        if (w < warp) before += s_warp[w];
        total += s_warp[w];
    }
    if (pass == 0) {
        if (threadIdx.x == 0) cta_count[blockIdx.x] = total;
        return;
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *n_sel = cta_count[blockIdx.x] + total; // cta_count = exclusive prefix
    if (!keep) return;
    const int o = cta_count[blockIdx.x] + before + __popc(m & ((1u << lane) - 1));
    sel[o] = i;
#pragma unroll
    for (int k = 0; k < PTS; ++k) {
        coords[static_cast<size_t>(o) * (2 * PTS) + 2 * k] = cell_of(uv[k].x, v.inv_width);
        coords[static_cast<size_t>(o) * (2 * PTS) + 2 * k + 1] = cell_of(uv[k].y, v.inv_height);
        pf[static_cast<size_t>(o) * (2 * PTS) + 2 * k] = uv[k].x;
        pf[static_cast<size_t>(o) * (2 * PTS) + 2 * k + 1] = uv[k].y;
    }
}

// exclusive scan of the per-CTA counts in place (one CTA; the map has at most a few thousand CTAs of 256 landmarks)
__global__ void __launch_bounds__(1024) map_count_scan_kernel(int32_t *__restrict__ cta_count, int n_cta) {
    __shared__ int s_part[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n_cta; base += 1024) {
        const int i = base + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int v = (i < n_cta) ? cta_count[i] : 0;
        int incl = v;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const int t = __shfl_up_sync(0xFFFFFFFFu, incl, s);
            if (lane >= s) incl += t;
        }
        if (lane == 31) s_part[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = s_part[lane];
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const int t = __shfl_up_sync(0xFFFFFFFFu, w, s);
                if (lane >= s) w += t;
            }
            s_part[lane] = w;
        }
        __syncthreads();
        const int excl = s_carry + (warp ? s_part[warp - 1] : 0) + incl - v;
        if (i < n_cta) cta_count[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
}

// out[j] = rows[sel[j]] for j < *n_sel: the representative descriptors of the selected landmarks (:606, :711)
__global__ void gather_rows_kernel(const uint4 *__restrict__ rows, const int32_t *__restrict__ sel, const int32_t *__restrict__ n_sel,
                                   uint4 *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= *n_sel) return;
    const Desc d = load_desc(rows, sel[j]);
    out[2 * static_cast<size_t>(j)] = d.lo;
    out[2 * static_cast<size_t>(j) + 1] = d.hi;
}

// Gates over the compacted rows.  ok[i1] = 1 for an accepted pair, 0 otherwise; *count -= rejected pairs.
// LINES = 0: feat = n2 x 2 (pl of the keyframe points); LINES = 1: feat = n2 x 3 (line equations le).
template <int LINES>
__global__ void map_gate_kernel(const double *__restrict__ pf, const int32_t *__restrict__ m12, const int32_t *__restrict__ n_sel,
                                const double *__restrict__ feat, int n2, double max_epip, uint8_t *__restrict__ ok,
                                int32_t *__restrict__ count) {
    const int i1 = blockIdx.x * blockDim.x + threadIdx.x;
    bool reject = false;
    if (i1 < *n_sel) {
        const int i2 = m12[i1];
        bool pass = false;
        if (i2 >= 0 && i2 < n2) {
            if (!LINES) {
                const double dx = __dsub_rn(pf[2 * static_cast<size_t>(i1)], feat[2 * static_cast<size_t>(i2)]);
                const double dy = __dsub_rn(pf[2 * static_cast<size_t>(i1) + 1], feat[2 * static_cast<size_t>(i2) + 1]);
                pass = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))) < max_epip;
            } else {
                const double *l = feat + 3 * static_cast<size_t>(i2);
                const double *p = pf + 4 * static_cast<size_t>(i1);
                const double e0 = __dadd_rn(__dadd_rn(__dmul_rn(l[0], p[0]), __dmul_rn(l[1], p[1])), l[2]);
                const double e1 = __dadd_rn(__dadd_rn(__dmul_rn(l[0], p[2]), __dmul_rn(l[1], p[3])), l[2]);
                pass = e0 < max_epip && e1 < max_epip;
            }
            reject = !pass;
        }
        ok[i1] = pass ? 1 : 0;
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, reject);
    if ((threadIdx.x & 31) == 0 && m) atomicSub(count, __popc(m));
}

} // namespace plm
