// Map landmarks: representative ("median") descriptor and mean observation direction, batched.
//
// PLSLAM::MapPoint::updateAverageDescDir (src/mapFeatures.cpp:51-93) and MapLine::updateAverageDescDir
// (:121-163) -- the producer of the med_desc rows that map-to-keyframe matching reads
// (mapHandler.cpp:596-609, :698-714).  For a landmark with n observations the reference builds the n x n
// Hamming matrix, takes per row the element at sorted position int(1 + 0.5*(n-1)) (self distance
// included) and keeps the FIRST row whose value is strictly smallest; the direction is the plain mean of
// the observation directions, summed in list order.
//
// Layout: observations of landmark l are rows obs_start[l] .. obs_start[l+1]-1 of one descriptor arena
// (n_obs x 32 B) and one direction arena (n_obs x 3 doubles).  The work is HBM / latency bound
// (n*n/2 pairs on n*56 bytes), so the kernels are organised around one coalesced pass over the arenas:
//
//  * med_desc_warp_kernel: lists of up to 32 observations (every list the SLAM system produces in practice: a
//    landmark is seen from a few dozen keyframes at most).  A warp takes 4 consecutive landmarks and packs them
//    by their longest list -- 4 landmarks x 8 lanes, 2 x 16 or 1 x 32 -- so short lists do not idle most of the
//    warp.  Sub-lane i of a group owns observation i in registers; row j is re-read from L1; the n distances of a
//    lane go to a private column of shared memory; the order statistic is a 9-step binary search on the value
//    range [0, 256]; the winner is a butterfly min over the group on (value << 5 | sub-lane) -- lowest row on ties
//    like the reference's strict `<`.
//  * med_desc_cta_kernel: landmarks with more than 32 observations are appended to a device work list by
//    the first kernel and handled by persistent CTAs (one thread per row, distances recomputed in each
//    search step from L1-resident rows).  No host round trip between the two launches.
#pragma once
#include "plm_common.cuh"

namespace plm {

constexpr int MED_WARPS = 4;   // landmarks in flight per CTA of the warp kernel
constexpr int MED_CTA_THREADS = 256;

struct MedArgs {
    const uint4 *desc;        // n_obs x 32 B
    const double *dirs;       // n_obs x 3, or null
    const int32_t *obs_start; // n_lm + 1
    long long n_obs;
    int n_lm;
    int32_t *med_idx;        // n_lm: winning position inside the list, -1 for an empty / malformed list
    uint4 *med_desc;         // n_lm x 32 B, or null
    const int32_t *dst_rows; // optional: landmark l writes row dst_rows[l] of med_desc (< 0: skip) -- scatter into a map DB
    double *med_dir;         // n_lm x 3, or null
    int32_t *work;           // [0] = number of long lists, [1 ..] = their landmark ids
};

// int(1 + 0.5 * (n - 1)) of mapFeatures.cpp:79 for n >= 2
__device__ __forceinline__ int med_rank(int n) { return 1 + (n - 1) / 2; }

__device__ __forceinline__ Desc shfl_desc(const Desc &q, int src) {
    Desc t;
    t.lo.x = __shfl_sync(0xFFFFFFFFu, q.lo.x, src);
    t.lo.y = __shfl_sync(0xFFFFFFFFu, q.lo.y, src);
    t.lo.z = __shfl_sync(0xFFFFFFFFu, q.lo.z, src);
    t.lo.w = __shfl_sync(0xFFFFFFFFu, q.lo.w, src);
    t.hi.x = __shfl_sync(0xFFFFFFFFu, q.hi.x, src);
    t.hi.y = __shfl_sync(0xFFFFFFFFu, q.hi.y, src);
    t.hi.z = __shfl_sync(0xFFFFFFFFu, q.hi.z, src);
    t.hi.w = __shfl_sync(0xFFFFFFFFu, q.hi.w, src);
    return t;
}

// Mean direction of one landmark, component c (mapFeatures.cpp:88-91): sequential fp64 sum in list
// order from zero, one IEEE division by n.  A single observation keeps its direction (constructor, :38).
__device__ __forceinline__ double med_dir_component(const double *__restrict__ dirs, long long lo, int n, int c) {
    if (n <= 0) return 0.0;
    if (n == 1) return __ldg(dirs + 3 * lo + c);
    double s = 0.0;
    for (int i = 0; i < n; ++i) s = __dadd_rn(s, __ldg(dirs + 3 * (lo + i) + c));
    return __ddiv_rn(s, static_cast<double>(n));
}

// One pass of a warp over 32 / G landmarks, G lanes each (G = 8, 16 or 32): sub-lane i of a group owns observation i.
// `n_in` / `lo_in` are this lane's group values (n_in = 0 for a group without a landmark in this pass; lists longer
// than G never get here).  `nmax` is the warp-uniform maximum of n over the groups of the pass.
template <int G>
__device__ __forceinline__ void med_desc_pass(const MedArgs &a, uint16_t *sd, long long lm, long long lo, int n, int nmax,
                                              bool valid) {
    const int lane = threadIdx.x & 31, sl = lane & (G - 1);
    Desc q{make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
    if (sl < n) q = load_desc(a.desc, lo + sl);
    int best = n > 0 ? 0 : -1;
    if (nmax >= 2) {
        // distances of row `sl` to every row j of its own landmark (L1-resident re-read of row j; the 4-POPC
        // carry-save form keeps the POPC pipe, 16 lanes/clk/SM, from binding)
        for (int j = 0; j < nmax; ++j) {
            if (j < n) {
                const Desc t = load_desc(a.desc, lo + j);
                sd[j * 32 + lane] = static_cast<uint16_t>(hamming256_csa4(q, t.lo, t.hi));
            }
        }
        // smallest v with #{j : d_j <= v} >= rank + 1  ==  the element at sorted position `rank`
        const int need = med_rank(n) + 1;
        int v_lo = 0, v_hi = 256;
#pragma unroll 1
        for (int step = 0; step < 9; ++step) { // 2^9 > 257 values: the interval is a single value after 9 halvings
            const int mid = (v_lo + v_hi) >> 1;
            int c = 0;
            for (int j = 0; j < nmax; ++j) c += (j < n && sd[j * 32 + lane] <= mid);
            if (v_lo < v_hi) {
                if (c >= need) v_hi = mid;
                else v_lo = mid + 1;
            }
        }
        uint32_t key = (sl < n && n >= 2) ? (static_cast<uint32_t>(v_lo) << 5) | sl : KEY32_ABSENT;
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) key = min(key, __shfl_xor_sync(0xFFFFFFFFu, key, off));
        if (n >= 2) best = static_cast<int>(key & 31u);
    }
    if (!valid) return;
    if (sl == 0) a.med_idx[lm] = best;
    if (a.med_desc) {
        const long long row = a.dst_rows ? __ldg(a.dst_rows + lm) : lm;
        if (row >= 0 && sl == (best < 0 ? 0 : best)) { // the winner writes its own registers (zeros for an empty list)
            a.med_desc[2 * row] = q.lo;
            a.med_desc[2 * row + 1] = q.hi;
        }
    }
    if (a.med_dir && a.dirs && sl < 3) a.med_dir[3 * lm + sl] = med_dir_component(a.dirs, lo, n, sl);
}

// A warp takes 4 consecutive landmarks and packs them by the longest list among them: 4 x 8 lanes, 2 x 16 lanes
// (two passes) or 1 x 32 lanes (four passes).
__global__ void __launch_bounds__(32 * MED_WARPS) med_desc_warp_kernel(MedArgs a) {
    __shared__ uint16_t sd_all[MED_WARPS][32 * 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint16_t *sd = sd_all[warp];
    const long long n_blocks = (static_cast<long long>(a.n_lm) + 3) / 4;
    for (long long blk = static_cast<long long>(blockIdx.x) * MED_WARPS + warp; blk < n_blocks;
         blk += static_cast<long long>(gridDim.x) * MED_WARPS) {
        // lanes 0..3 read the block's landmarks
        const long long my_lm = blk * 4 + (lane & 3);
        long long my_lo = 0, my_n = 0;
        if (my_lm < a.n_lm) {
            my_lo = __ldg(a.obs_start + my_lm);
            my_n = static_cast<long long>(__ldg(a.obs_start + my_lm + 1)) - my_lo;
            if (my_lo < 0 || my_n < 0 || my_lo + my_n > a.n_obs) my_n = 0; // malformed range: treated as empty
            if (my_n > 32) {
                if (lane < 4) a.work[1 + atomicAdd(a.work, 1)] = static_cast<int32_t>(my_lm);
                my_n = -1; // handled by med_desc_cta_kernel
            }
        } else {
            my_n = -2; // past the end
        }
        int nmax = static_cast<int>(my_n);
        nmax = max(nmax, __shfl_xor_sync(0xFFFFFFFFu, nmax, 1));
        nmax = max(nmax, __shfl_xor_sync(0xFFFFFFFFu, nmax, 2));
        nmax = __shfl_sync(0xFFFFFFFFu, nmax, 0);
        if (nmax <= 8) {
            const int g = lane >> 3;
            const int n = static_cast<int>(__shfl_sync(0xFFFFFFFFu, my_n, g));
            const long long lo = __shfl_sync(0xFFFFFFFFu, my_lo, g);
            med_desc_pass<8>(a, sd, blk * 4 + g, lo, max(n, 0), nmax, n >= 0);
        } else if (nmax <= 16) {
            for (int pass = 0; pass < 2; ++pass) {
                const int g = 2 * pass + (lane >> 4);
                const int n = static_cast<int>(__shfl_sync(0xFFFFFFFFu, my_n, g));
                const long long lo = __shfl_sync(0xFFFFFFFFu, my_lo, g);
                int pmax = max(n, __shfl_xor_sync(0xFFFFFFFFu, n, 16));
                med_desc_pass<16>(a, sd, blk * 4 + g, lo, max(n, 0), pmax, n >= 0);
                __syncwarp();
            }
        } else {
            for (int pass = 0; pass < 4; ++pass) {
                const int n = static_cast<int>(__shfl_sync(0xFFFFFFFFu, my_n, pass));
                const long long lo = __shfl_sync(0xFFFFFFFFu, my_lo, pass);
                if (n >= 0) med_desc_pass<32>(a, sd, blk * 4 + pass, lo, n, n, true);
                __syncwarp();
            }
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(MED_CTA_THREADS) med_desc_cta_kernel(MedArgs a) {
    __shared__ unsigned long long s_best;
    const int n_work = a.work[0];
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const long long lm = a.work[1 + w];
        const long long lo = __ldg(a.obs_start + lm);
        const int n = static_cast<int>(static_cast<long long>(__ldg(a.obs_start + lm + 1)) - lo);
        if (threadIdx.x == 0) s_best = KEY64_ABSENT;
        __syncthreads();
        const int need = med_rank(n) + 1;
        for (int i = threadIdx.x; i < n; i += MED_CTA_THREADS) {
            const Desc q = load_desc(a.desc, lo + i);
            int v_lo = 0, v_hi = 256;
            while (v_lo < v_hi) {
                const int mid = (v_lo + v_hi) >> 1;
                int c = 0;
                for (int j = 0; j < n; ++j) c += (hamming256(q, load_desc(a.desc, lo + j)) <= mid);
                if (c >= need) v_hi = mid;
                else v_lo = mid + 1;
            }
            atomicMin(&s_best, make_key64(static_cast<uint32_t>(v_lo), static_cast<uint32_t>(i)));
        }
        __syncthreads();
        const int best = static_cast<int>(s_best & 0xFFFFFFFFull);
        if (threadIdx.x == 0) a.med_idx[lm] = best;
        if (a.med_desc && threadIdx.x < 2) {
            const long long row = a.dst_rows ? __ldg(a.dst_rows + lm) : lm;
            if (row >= 0) a.med_desc[2 * row + threadIdx.x] = __ldg(a.desc + 2 * (lo + best) + threadIdx.x);
        }
        if (a.med_dir && a.dirs && threadIdx.x >= 32 && threadIdx.x < 35)
            a.med_dir[3 * lm + (threadIdx.x - 32)] = med_dir_component(a.dirs, lo, n, threadIdx.x - 32);
        __syncthreads();
    }
}

} // namespace plm
