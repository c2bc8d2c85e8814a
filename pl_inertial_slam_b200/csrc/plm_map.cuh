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
//    warp.  Sub-lane i of a group owns observation i in registers; row j is re-read from L1; the distances of a
//    lane's row stay in registers, padded to the group width, and go through a bitonic sorting network (all
//    indices compile-time); the winner is a butterfly min over the group on (value << 5 | sub-lane) -- lowest row on ties
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

// One pass of a warp over 32 / G landmarks, G lanes each (G = 8, 16 or 32): sub-lane i of a group owns observation i
// and keeps the G distances of its row in REGISTERS (every loop below is unrolled over compile-time indices).
// `n` / `lo` are this lane's group values (n = 0 for a group without a landmark in this pass; lists longer than G
// never get here); `nmax` is the warp-uniform maximum of n over the groups of the pass.
template <int G>
__device__ __forceinline__ void med_desc_pass(const MedArgs &a, double *sdir, long long lm, long long lo, int n, int nmax,
                                              bool valid) {
    const int lane = threadIdx.x & 31, sl = lane & (G - 1);
    Desc q{make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
    const bool want_dir = a.med_dir && a.dirs;
    if (sl < n) {
        q = load_desc(a.desc, lo + sl);
        if (want_dir) { // the owner lane fetches its observation's direction now, next to the descriptor (one HBM round trip)
            const double *dp = a.dirs + 3 * (lo + sl);
            sdir[3 * lane] = __ldg(dp);
            sdir[3 * lane + 1] = __ldg(dp + 1);
            sdir[3 * lane + 2] = __ldg(dp + 2);
        }
    }
    int best = n > 0 ? 0 : -1;
    if (nmax >= 2) {
        // distances of row `sl` to every row j of its own landmark (L1-resident re-read of row j; the 4-POPC
        // carry-save form keeps the POPC pipe, 16 lanes/clk/SM, from binding); 0xFFFF pads the row to G entries
        int d[G];
#pragma unroll
        for (int j = 0; j < G; ++j) {
            d[j] = 0xFFFF;
            if (j < nmax) { // warp-uniform
                if (j < n) {
                    const Desc t = load_desc(a.desc, lo + j);
                    d[j] = hamming256_csa4(q, t.lo, t.hi);
                }
            }
        }
        // bitonic sorting network over the register row (ascending; the pads sink to the end)
#pragma unroll
        for (int k = 2; k <= G; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
                for (int i = 0; i < G; ++i) {
                    const int l = i ^ j;
                    if (l > i) {
                        const int x = d[i], y = d[l];
                        const bool up = (i & k) == 0;
                        d[i] = up ? min(x, y) : max(x, y);
                        d[l] = up ? max(x, y) : min(x, y);
                    }
                }
            }
        }
        // the element at sorted position int(1 + 0.5*(n-1)) (mapFeatures.cpp:79)
        const int rank = med_rank(n);
        int med = d[0];
#pragma unroll
        for (int j = 1; j < G; ++j) med = (j == rank) ? d[j] : med;
        uint32_t key = (sl < n && n >= 2) ? (static_cast<uint32_t>(med) << 5) | sl : KEY32_ABSENT;
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) key = min(key, __shfl_xor_sync(0xFFFFFFFFu, key, off));
        if (n >= 2) best = static_cast<int>(key & 31u);
    }
    if (valid) {
        if (sl == 0) a.med_idx[lm] = best;
        if (a.med_desc) {
            const long long row = a.dst_rows ? __ldg(a.dst_rows + lm) : lm;
            if (row >= 0 && sl == (best < 0 ? 0 : best)) { // the winner writes its own registers (zeros for an empty list)
                a.med_desc[2 * row] = q.lo;
                a.med_desc[2 * row + 1] = q.hi;
            }
        }
    }
    if (want_dir) {
        __syncwarp();
        if (valid && sl < 3) { // mean direction (mapFeatures.cpp:88-91): sequential sum in list order from zero, / n
            const double *g = sdir + 3 * (lane - sl) + sl; // component sl of the group's observation 0
            double r = 0.0;
            if (n == 1) {
                r = g[0]; // a single observation keeps its direction (constructor, :38)
            } else if (n >= 2) {
                double acc = 0.0;
                for (int j = 0; j < n; ++j) acc = __dadd_rn(acc, g[3 * j]);
                r = __ddiv_rn(acc, static_cast<double>(n));
            }
            a.med_dir[3 * lm + sl] = r;
        }
        __syncwarp();
    }
}

// A warp takes 4 consecutive landmarks and packs them by the longest list among them: 4 x 8 lanes, 2 x 16 lanes
// (two passes) or 1 x 32 lanes (four passes).
__global__ void __launch_bounds__(32 * MED_WARPS) med_desc_warp_kernel(MedArgs a) {
    __shared__ double sdir_all[MED_WARPS][32 * 3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *sdir = sdir_all[warp];
    const long long n_blocks = (static_cast<long long>(a.n_lm) + 3) / 4;
    const long long stride = static_cast<long long>(gridDim.x) * MED_WARPS;
    // lanes read the offsets of the block's 4 landmarks; the next block's are fetched one iteration ahead
    auto fetch = [&](long long blk, long long &lo, long long &n) {
        const long long my_lm = blk * 4 + (lane & 3);
        lo = 0;
        n = -2; // past the end
        if (blk < n_blocks && my_lm < a.n_lm) {
            lo = __ldg(a.obs_start + my_lm);
            n = static_cast<long long>(__ldg(a.obs_start + my_lm + 1)) - lo;
        }
    };
    long long blk = static_cast<long long>(blockIdx.x) * MED_WARPS + warp;
    long long nx_lo, nx_n;
    fetch(blk, nx_lo, nx_n);
    for (; blk < n_blocks; blk += stride) {
        long long my_lo = nx_lo, my_n = nx_n;
        fetch(blk + stride, nx_lo, nx_n);
        if (my_n > -2) {
            if (my_lo < 0 || my_n < 0 || my_lo + my_n > a.n_obs) my_n = 0; // malformed range: treated as empty
            if (my_n > 32) {
                if (lane < 4) a.work[1 + atomicAdd(a.work, 1)] = static_cast<int32_t>(blk * 4 + lane);
                my_n = -1; // handled by med_desc_cta_kernel
            }
        }
        int nmax = static_cast<int>(my_n);
        nmax = max(nmax, __shfl_xor_sync(0xFFFFFFFFu, nmax, 1));
        nmax = max(nmax, __shfl_xor_sync(0xFFFFFFFFu, nmax, 2));
        nmax = __shfl_sync(0xFFFFFFFFu, nmax, 0);
        if (nmax <= 8) {
            const int g = lane >> 3;
            const int n = static_cast<int>(__shfl_sync(0xFFFFFFFFu, my_n, g));
            const long long lo = __shfl_sync(0xFFFFFFFFu, my_lo, g);
            med_desc_pass<8>(a, sdir, blk * 4 + g, lo, max(n, 0), nmax, n >= 0);
        } else if (nmax <= 16) {
            for (int pass = 0; pass < 2; ++pass) {
                const int g = 2 * pass + (lane >> 4);
                const int n = static_cast<int>(__shfl_sync(0xFFFFFFFFu, my_n, g));
                const long long lo = __shfl_sync(0xFFFFFFFFu, my_lo, g);
                const int pmax = max(n, __shfl_xor_sync(0xFFFFFFFFu, n, 16));
                med_desc_pass<16>(a, sdir, blk * 4 + g, lo, max(n, 0), pmax, n >= 0);
            }
        } else {
            for (int pass = 0; pass < 4; ++pass) {
                const int n = static_cast<int>(__shfl_sync(0xFFFFFFFFu, my_n, pass));
                const long long lo = __shfl_sync(0xFFFFFFFFu, my_lo, pass);
                if (n >= 0) med_desc_pass<32>(a, sdir, blk * 4 + pass, lo, n, n, true);
            }
        }
    }
}

constexpr int MED_CTA_CACHE_N = 128; // lists up to this length keep their distance matrix in shared memory

__global__ void __launch_bounds__(MED_CTA_THREADS) med_desc_cta_kernel(MedArgs a) {
    __shared__ unsigned long long s_best;
    __shared__ uint16_t s_dist[MED_CTA_CACHE_N * MED_CTA_CACHE_N];
    const int n_work = a.work[0];
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const long long lm = a.work[1 + w];
        const long long lo = __ldg(a.obs_start + lm);
        const int n = static_cast<int>(static_cast<long long>(__ldg(a.obs_start + lm + 1)) - lo);
        if (threadIdx.x == 0) s_best = KEY64_ABSENT;
        const bool cached = n <= MED_CTA_CACHE_N;
        if (cached) { // every pair once, all threads (row-major n x n)
            for (int idx = threadIdx.x; idx < n * n; idx += MED_CTA_THREADS) {
                const int i = idx / n, j = idx - i * n;
                s_dist[idx] = static_cast<uint16_t>(hamming256(load_desc(a.desc, lo + i), load_desc(a.desc, lo + j)));
            }
        }
        __syncthreads();
        const int need = med_rank(n) + 1;
        for (int i = threadIdx.x; i < n; i += MED_CTA_THREADS) {
            const Desc q = load_desc(a.desc, lo + i);
            int v_lo = 0, v_hi = 256;
            while (v_lo < v_hi) {
                const int mid = (v_lo + v_hi) >> 1;
                int c = 0;
                if (cached) {
                    for (int j = 0; j < n; ++j) c += (s_dist[i * n + j] <= mid);
                } else {
                    for (int j = 0; j < n; ++j) c += (hamming256(q, load_desc(a.desc, lo + j)) <= mid);
                }
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
