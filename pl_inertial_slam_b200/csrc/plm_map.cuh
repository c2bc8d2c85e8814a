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
//  * med_desc_warp_kernel: one warp per landmark with n <= 32 (every list the SLAM system produces in
//    practice: a landmark is seen from a few dozen keyframes at most).  Lane i owns observation i in
//    registers; row j reaches all lanes by shuffle; the n distances of a lane go to a private column of
//    shared memory; the order statistic is a 9-step binary search on the value range [0, 256]; the
//    winner is one __reduce_min_sync on (value << 5 | lane) -- lowest row on ties like the reference's
//    strict `<`.
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

__global__ void __launch_bounds__(32 * MED_WARPS) med_desc_warp_kernel(MedArgs a) {
    __shared__ uint16_t sd_all[MED_WARPS][32 * 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint16_t *sd = sd_all[warp];
    for (long long lm = static_cast<long long>(blockIdx.x) * MED_WARPS + warp; lm < a.n_lm;
         lm += static_cast<long long>(gridDim.x) * MED_WARPS) {
        const long long lo = __ldg(a.obs_start + lm);
        long long n64 = static_cast<long long>(__ldg(a.obs_start + lm + 1)) - lo;
        if (lo < 0 || n64 < 0 || lo + n64 > a.n_obs) n64 = 0; // malformed range: treated as empty
        if (n64 > 32) {
            if (lane == 0) a.work[1 + atomicAdd(a.work, 1)] = static_cast<int32_t>(lm);
            continue;
        }
        const int n = static_cast<int>(n64);
        Desc q{make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
        if (lane < n) q = load_desc(a.desc, lo + lane);
        int best = n > 0 ? 0 : -1;
        if (n >= 2) {
            for (int j = 0; j < n; ++j) {
                const Desc t = shfl_desc(q, j);
                sd[j * 32 + lane] = static_cast<uint16_t>(hamming256(q, t));
            }
            // smallest v with #{j : d_j <= v} >= rank + 1  ==  the element at sorted position `rank`
            const int need = med_rank(n) + 1;
            int v_lo = 0, v_hi = 256;
            while (v_lo < v_hi) {
                const int mid = (v_lo + v_hi) >> 1;
                int c = 0;
                for (int j = 0; j < n; ++j) c += (sd[j * 32 + lane] <= mid);
                if (c >= need) v_hi = mid;
                else v_lo = mid + 1;
            }
            const uint32_t key = lane < n ? (static_cast<uint32_t>(v_lo) << 5) | lane : KEY32_ABSENT;
            best = static_cast<int>(__reduce_min_sync(0xFFFFFFFFu, key) & 31u);
        }
        if (lane == 0) a.med_idx[lm] = best;
        if (a.med_desc) {
            const Desc w = shfl_desc(q, best < 0 ? 0 : best); // empty list: lane 0 holds zeros
            const long long row = a.dst_rows ? __ldg(a.dst_rows + lm) : lm;
            if (row >= 0 && lane < 2) a.med_desc[2 * row + lane] = lane ? w.hi : w.lo;
        }
        if (a.med_dir && a.dirs && lane < 3) a.med_dir[3 * lm + lane] = med_dir_component(a.dirs, lo, n, lane);
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
