// Grid-windowed matching: StVO::matchGrid for points (stvo-pl/src/matching.cpp:111-177) and lines
// (:179-258) over a CSR copy of StVO::GridStructure (stvo-pl/src/gridStructure.cpp:43-83).
//
// The reference walks the query rows i1 in order and keeps distances[i2] = running minimum per
// train feature; a pair only counts for row i1 when it strictly improves that minimum.  Stated
// without the loop-carried dependence (SURVEY 8a note 1):
//     live(i1,i2)  <=>  D(i1,i2) < min{ D(i1',i2) : i1' < i1, i2 in cand(i1') }
//     m21[i2]       =   the live pair of column i2 with the smallest D  (= lowest i1 at the minimum)
//     row result    =   two smallest D over the live pairs of the row, fp64 ratio test, mutual check
// Parallel form used here: consecutive rows are split into chunks, one warp per chunk.
//   phase A  every warp computes the column minima of its own chunk          (wmin[w][i2], u16)
//   scan     exclusive prefix-min over the chunks in row order  -> wmin[w][i2] becomes the
//            threshold a pair of chunk w must beat before the chunk starts
//   phase C  every warp replays its rows IN ORDER against its private threshold array, which
//            reproduces the sequential semantics inside the chunk exactly; lanes own candidates.
// A frame-sized job (n1 <= ~1-2k rows) is one CTA and one launch including the mutual check; a
// map-sized job (200k rows) uses one CTA per 32*W rows with the scan done by a second kernel.
// Distances are recomputed in phase C instead of being stored: a pair list would need a dynamic
// allocation, and 8 POPC are cheaper than the round trip.
#pragma once
#include <cooperative_groups.h>
#include <type_traits>

#include "plm_common.cuh"

namespace plm {

namespace cg = cooperative_groups;

// Phase timeline of the one-launch frame kernel (debug builds only: PLM_BUILD_DEFINES=-DPLM_TIMELINE): thread 0 of a
// CTA stamps clock64() at phase boundaries, tools/latency_bench prints the deltas.
#ifdef PLM_TIMELINE
__device__ long long g_timeline[128][24];
#define PLM_TL(k) do { if (threadIdx.x == 0) g_timeline[blockIdx.x < 127 ? blockIdx.x : 127][k] = clock64(); } while (0)
#else
#define PLM_TL(k) do { } while (0)
#endif

struct GridJob {
    const int32_t *coords;     // n1 x 2 (points) or n1 x 4 (lines), grid-cell coordinates
    const uint4 *d1;           // n1 query descriptors
    const int32_t *cell_start; // grid_rows*grid_cols + 1
    const int32_t *cell_items;
    const uint4 *d2;           // n2 train descriptors
    const double *dirs2;       // n2 x 2 (lines only)
    int32_t *m12;              // in/out, n1
    int32_t *count;            // accepted - culled
    int32_t n1, n2, is_lines, pad_;
    int32_t win[4];
    long long i1_base;         // global row index of local row 0 (row-sharded map; 0 otherwise)
    int32_t q_row_base, pad2_; // d1 / coords point at row q_row_base (staged per-CTA row windows)
};

struct GridParams {
    int grid_rows, grid_cols;
    int best_lr;
    int rows_per_warp;          // chunked launch only; the fused kernel derives it from n1
    int rows_per_cta, cap_pairs; // row-parallel launch (grid_rows_kernel): rows per CTA, pair-list capacity
    // The rows at the START of the map have no thresholds yet (every pair survives into pass 1 and the record chains of
    // a column are longest there), so the first head_ctas CTAs only take head_rows rows each: the spare CTA slots of
    // the one-wave launch shorten the tail of pass 1.  0 = uniform.
    int head_ctas, head_rows;
    // pass 0 -> pass 1 hand-over of the row-parallel launch: pass 0 leaves every pair it evaluated (i2 | row | D) in
    // ent_g (a fixed region of ent_per_cta entries per CTA) with one seg_tab record per list segment, so that pass 1
    // only streams them against its thresholds instead of walking the windows and computing the distances again.
    // seg_cnt[cta] = number of records, -1 when the CTA's pairs did not fit (pass 1 then recomputes).  Null = off.
    uint32_t *ent_g;
    int4 *seg_tab;      // [n_cta][GRID_SEG_TAB]: (offset in ent_g, entries, block of rows, 0)
    int32_t *seg_cnt;   // [n_cta]
    int ent_per_cta;
    int init_flags; // pass 0 of the multi-launch form also initialises: bit 0 m12 = -1 for its rows, bit 1 *count = 0, bit 2 m21key = absent
    // fused kernel: when staged != 0 every input of the job is first copied into shared memory with
    // coalesced 128-bit loads (capacities below, in elements), so the per-row dependent accesses
    // (coords -> cell_start -> cell_items -> descriptor / threshold) cost shared-memory latency
    int staged, cap_n1, cap_items;
    int n_items_p1; // row-parallel kernels: number of grid items + 1 when the host knows it (0: read cell_start[n_cells])
    double ratio, line_sim_th;
    // chunked (multi-CTA) launch only:
    uint16_t *cta_min;          // [n_cta][n2] per-CTA column minima, turned into thresholds in place
    unsigned long long *m21key; // [n2] (D << 32 | i1) of the best live pair per column
};

constexpr uint16_t D_INF = 0xFFFFu;
constexpr int GRID_KEY_BITS = 22;

struct RowWindows {
    int n_win;
    int min_x[2], nx[2], min_y[2], max_y[2];
    int n_ranges;
};

// GridStructure::get window clamp (gridStructure.cpp:67-71).  One window per point, two per line
// (start and end cell, matching.cpp:213-215).
__device__ __forceinline__ void clamp_window(RowWindows &rw, int k, int x, int y, const int32_t win[4],
                                             int grid_rows, int grid_cols) {
    const long long ax = static_cast<long long>(x) - win[0], bx = static_cast<long long>(x) + win[1] + 1;
    const long long ay = static_cast<long long>(y) - win[2], by = static_cast<long long>(y) + win[3] + 1;
    const int min_x = static_cast<int>(max(0ll, ax)), max_x = static_cast<int>(min(static_cast<long long>(grid_cols), bx));
    const int min_y = static_cast<int>(max(0ll, ay)), max_y = static_cast<int>(min(static_cast<long long>(grid_rows), by));
    const bool any = max_x > min_x && max_y > min_y;
    rw.min_x[k] = min_x;
    rw.nx[k] = any ? max_x - min_x : 0;
    rw.min_y[k] = min_y;
    rw.max_y[k] = max_y;
}

// Calls body(i2) once per candidate slot of the row's window(s), 32 slots at a time with lane ==
// slot; lanes without a slot get i2 = -1.  All 32 lanes call body together (it may use warp
// collectives).  Because cell ids are x-major, one window column is one contiguous item range:
// lanes first fetch the <= 32 range bounds, scan their lengths, then map flat slots to items with
// a shuffle binary search -- no per-range serial loop.
template <class Body>
__device__ __forceinline__ void for_each_candidate(const RowWindows &rw, const GridJob &job, int grid_rows,
                                                   int lane, Body &&body) {
    for (int rg0 = 0; rg0 < rw.n_ranges; rg0 += 32) {
        const int t = rg0 + lane;
        int lo = 0, hi = 0;
        if (t < rw.n_ranges) {
            const int k = (t < rw.nx[0]) ? 0 : 1;
            const int x = rw.min_x[k] + (k ? t - rw.nx[0] : t);
            lo = job.cell_start[x * grid_rows + rw.min_y[k]];
            hi = job.cell_start[x * grid_rows + rw.max_y[k]];
        }
        const int cnt = max(hi - lo, 0);
        int incl = cnt;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const int v = __shfl_up_sync(0xFFFFFFFFu, incl, s);
            if (lane >= s) incl += v;
        }
        const int tot = __shfl_sync(0xFFFFFFFFu, incl, 31);
        const int excl = incl - cnt;
        for (int c0 = 0; c0 < tot; c0 += 32) {
            const int p = c0 + lane;
            int r = 0;
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) {
                const int v = __shfl_sync(0xFFFFFFFFu, incl, r + s - 1);
                if (v <= p) r += s;
            }
            const int lo_r = __shfl_sync(0xFFFFFFFFu, lo, r);
            const int ex_r = __shfl_sync(0xFFFFFFFFu, excl, r);
            int i2 = -1;
            if (p < tot) i2 = job.cell_items[lo_r + (p - ex_r)];
            body(i2);
        }
    }
}

struct RowQuery {
    RowWindows rw;
    Desc q;
    double vx, vy; // normalised query direction (lines)
};

// Window(s) of row i1 only (no descriptor, no direction).
__device__ __forceinline__ RowWindows load_windows(const GridJob &job, const GridParams &gp, int i1) {
    RowWindows rw;
    if (!job.is_lines) {
        const long long ci = i1 - job.q_row_base;
        rw.n_win = 1;
        clamp_window(rw, 0, job.coords[2 * ci], job.coords[2 * ci + 1], job.win, gp.grid_rows, gp.grid_cols);
        rw.nx[1] = 0;
        rw.min_x[1] = rw.min_y[1] = rw.max_y[1] = 0;
    } else {
        const int32_t *cp = job.coords + 4 * static_cast<long long>(i1 - job.q_row_base);
        rw.n_win = 2;
        clamp_window(rw, 0, cp[0], cp[1], job.win, gp.grid_rows, gp.grid_cols);
        clamp_window(rw, 1, cp[2], cp[3], job.win, gp.grid_rows, gp.grid_cols);
    }
    rw.n_ranges = rw.nx[0] + rw.nx[1];
    return rw;
}

__device__ __forceinline__ RowQuery load_row(const GridJob &job, const GridParams &gp, int i1) {
    RowQuery r;
    r.q = load_desc_any(job.d1, i1 - job.q_row_base);
    r.vx = r.vy = 0.0;
    r.rw = load_windows(job, gp, i1);
    if (job.is_lines) {
        const int32_t *cp = job.coords + 4 * static_cast<long long>(i1 - job.q_row_base);
        // matching.cpp:210-211 + matching.h:43-48: v = (ep - sp) as ints -> double, divided by
        // sqrt(x*x + y*y).  Separate roundings (the reference has no FMA); 0/0 = NaN is kept.
        const double dx = static_cast<double>(cp[2] - cp[0]), dy = static_cast<double>(cp[3] - cp[1]);
        const double mag = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        r.vx = __ddiv_rn(dx, mag);
        r.vy = __ddiv_rn(dy, mag);
    }
    return r;
}

// matching.cpp:141 (range check) and :221 (direction filter; `fabs(NaN) < th` is false -> kept).
__device__ __forceinline__ bool candidate_ok(const GridJob &job, const GridParams &gp, const RowQuery &r, int i2) {
    if (i2 < 0 || i2 >= job.n2) return false;
    if (job.is_lines) {
        const double2 d = make_double2(job.dirs2[2 * static_cast<long long>(i2)], job.dirs2[2 * static_cast<long long>(i2) + 1]);
        const double dp = __dadd_rn(__dmul_rn(r.vx, d.x), __dmul_rn(r.vy, d.y));
        if (fabs(dp) < gp.line_sim_th) return false;
    }
    return true;
}

// Phase A for the rows [row0, row1) of one warp: wmin[i2] = min D over the chunk's pairs.
__device__ __forceinline__ void chunk_minima(const GridJob &job, const GridParams &gp, int row0, int row1,
                                             uint16_t *wmin, int lane) {
    for (int i1 = row0; i1 < row1; ++i1) {
        const RowQuery r = load_row(job, gp, i1);
        for_each_candidate(r.rw, job, gp.grid_rows, lane, [&](int i2) {
            if (candidate_ok(job, gp, r, i2)) {
                const int d = hamming256(r.q, load_desc_any(job.d2, i2));
                // lanes holding the same i2 (a line listed in several cells) write the same value
                if (d < wmin[i2]) wmin[i2] = static_cast<uint16_t>(d);
            }
            __syncwarp();
        });
    }
}

// Phase C for the rows [row0, row1) of one warp.  wthr is the warp's private threshold array
// (best_lr) or duplicate-stamp array (!best_lr).  M21 receives (D, i1) of live pairs.
// Returns the number of accepted rows (valid in lane 0).
template <class M21>
__device__ __forceinline__ int chunk_match(const GridJob &job, const GridParams &gp, int row0, int row1,
                                           uint16_t *wthr, int lane, M21 &&m21_update) {
    int accepted = 0;
    for (int i1 = row0; i1 < row1; ++i1) {
        const RowQuery r = load_row(job, gp, i1);
        uint32_t b0 = KEY32_ABSENT, b1 = KEY32_ABSENT;
        const uint16_t stamp = static_cast<uint16_t>(i1 - row0);
        for_each_candidate(r.rw, job, gp.grid_rows, lane, [&](int i2) {
            bool ok = candidate_ok(job, gp, r, i2);
            // the candidate set is a set (std::unordered_set, matching.cpp:136/213): drop repeats
            // inside this batch of 32 ...
            const unsigned peers = __match_any_sync(0xFFFFFFFFu, ok ? i2 : -1 - lane);
            ok = ok && (lane == __ffs(peers) - 1);
            if (ok) {
                const int d = hamming256(r.q, load_desc_any(job.d2, i2));
                bool live;
                if (gp.best_lr) {
                    // ... and across batches: a repeat no longer beats the threshold it set itself
                    live = d < wthr[i2];
                    if (live) {
                        wthr[i2] = static_cast<uint16_t>(d);
                        m21_update(i2, d, i1);
                    }
                } else {
                    live = wthr[i2] != stamp;
                    wthr[i2] = stamp;
                }
                if (live) top2_insert(b0, b1, (static_cast<uint32_t>(d) << GRID_KEY_BITS) | static_cast<uint32_t>(i2));
            }
            __syncwarp();
        });
        // warp top-2: keys are distinct (one per train index) except the ABSENT sentinel
        const uint32_t m0 = __reduce_min_sync(0xFFFFFFFFu, b0);
        const uint32_t m1 = __reduce_min_sync(0xFFFFFFFFu, (b0 == m0) ? b1 : b0);
        if (lane == 0 && r.rw.n_ranges > 0) {
            // matching.cpp:160 / :241 -- int -> double, one double multiply, strict compare
            const int best_d = (m0 == KEY32_ABSENT) ? 0x7FFFFFFF : static_cast<int>(m0 >> GRID_KEY_BITS);
            const int best_d2 = (m1 == KEY32_ABSENT) ? 0x7FFFFFFF : static_cast<int>(m1 >> GRID_KEY_BITS);
            if (m0 != KEY32_ABSENT &&
                static_cast<double>(best_d) < __dmul_rn(static_cast<double>(best_d2), gp.ratio)) {
                job.m12[i1] = static_cast<int32_t>(m0 & ((1u << GRID_KEY_BITS) - 1));
                ++accepted;
            }
        }
    }
    return accepted;
}

// Cooperative copy of n_bytes (any alignment of the byte count; src and dst 16-byte aligned when
// n_bytes >= 16) into shared memory.
__device__ __forceinline__ void stage_bytes(unsigned char *dst, const void *src, size_t n_bytes) {
    const size_t n16 = n_bytes >> 4;
    const uint4 *s4 = static_cast<const uint4 *>(src);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst);
    if ((reinterpret_cast<size_t>(src) & 15) == 0) {
        for (size_t i = threadIdx.x; i < n16; i += blockDim.x) d4[i] = __ldg(s4 + i);
        const unsigned char *sb = static_cast<const unsigned char *>(src);
        for (size_t i = (n16 << 4) + threadIdx.x; i < n_bytes; i += blockDim.x) dst[i] = sb[i];
    } else {
        const uint32_t *s1 = static_cast<const uint32_t *>(src); // every arena is at least 4-byte aligned
        uint32_t *d1 = reinterpret_cast<uint32_t *>(dst);
        for (size_t i = threadIdx.x; i < (n_bytes >> 2); i += blockDim.x) d1[i] = __ldg(s1 + i);
    }
}

// Same copy issued as 16-byte cp.async (LDGSTS): every region of a staging block is in flight before the first one
// lands; stage_wait() + a CTA barrier make the data visible.  Falls back to the synchronous copy for a source that is
// not 16-byte aligned.
__device__ __forceinline__ void stage_bytes_async(unsigned char *dst, const void *src, size_t n_bytes) {
    if ((reinterpret_cast<size_t>(src) & 15) != 0) {
        stage_bytes(dst, src, n_bytes);
        return;
    }
    const size_t n16 = n_bytes >> 4;
    const unsigned char *sb = static_cast<const unsigned char *>(src);
    const uint32_t d0 = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
    for (size_t i = threadIdx.x; i < n16; i += blockDim.x)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + static_cast<uint32_t>(i << 4)), "l"(sb + (i << 4)) : "memory");
    for (size_t i = (n16 << 4) + threadIdx.x; i < n_bytes; i += blockDim.x) dst[i] = sb[i];
}
__device__ __forceinline__ void stage_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__host__ __device__ inline size_t grid_align16(size_t v) { return (v + 15) & ~size_t(15); }

// Shared-memory footprint of the fused kernel (also used by the host to decide whether to stage).
__host__ __device__ inline size_t grid_fused_smem(int warps, int n2_max, int staged, int cap_n1, int cap_items,
                                                  int n_cells, bool any_lines) {
    size_t b = grid_align16(static_cast<size_t>(warps) * n2_max * 2) + grid_align16(static_cast<size_t>(n2_max) * 4);
    if (staged) {
        b += grid_align16(static_cast<size_t>(n_cells + 1) * 4) + grid_align16(static_cast<size_t>(cap_items) * 4);
        b += static_cast<size_t>(n2_max) * 32 + static_cast<size_t>(cap_n1) * 32;
        b += grid_align16(static_cast<size_t>(cap_n1) * 16);
        if (any_lines) b += static_cast<size_t>(n2_max) * 16;
    }
    return b;
}

// ---- pair-list form for frame-sized jobs ---------------------------------------------------------------
// With a few candidates per row (stereo window: 2-3, +-3 cells: 10-15) the warp-per-chunk phases above spend
// their time in warp-synchronous bookkeeping.  The same semantics without any loop-carried dependence:
//   P1  thread per row: number of candidate slots of its window(s)              -> block scan -> rowoff
//   P2  thread per row: walk the slots, range check, in-row de-duplication (the candidate set is a set,
//       matching.cpp:136 / :213; a line sits in several cells), direction filter, distance
//                                                                                -> pair[p] = (i2, D), prow[p] = i1
//   P3  counting sort of the valid pairs by column i2                            -> colstart, cpair
//   P4  thread per pair: with bestLRMatches the pair is DEAD when an earlier row i1' < i1 holds D' <= D in its
//       column bucket (it does not improve distances[i2], :145-150)
//       thread per column: m21[i2] = lowest i1 at the column minimum
//   P5  thread per row: two smallest D over its live pairs, fp64 ratio test (:160 / :241) -> m12
//   P6  mutual check (:166-174)
// Everything lives in shared memory; a job with more slots than the arrays hold reports false and the caller
// falls back to the chunk phases.
constexpr uint32_t PAIR_INVALID = 0xFFFFFFFFu;
constexpr uint16_t PROW_DEAD = 0x8000u;

struct PairArrays {
    uint32_t *pair;    // [cap]  i2 << 9 | D
    uint16_t *prow;    // [cap]  i1 | PROW_DEAD
    uint16_t *cpair;   // [cap]  pair slots grouped by column
    int32_t *rowoff;   // [n1 + 1]
    int32_t *colstart; // [n2 + 1]
    int32_t *colaux;   // [n2]     fill cursors, then m21
    int cap;           // <= 65535
};

__host__ __device__ inline size_t pairlist_smem(int cap_pairs, int cap_n1, int cap_n2) {
    return grid_align16(static_cast<size_t>(cap_pairs) * 4) + 2 * grid_align16(static_cast<size_t>(cap_pairs) * 2) +
           grid_align16(static_cast<size_t>(cap_n1 + 1) * 4) + grid_align16(static_cast<size_t>(cap_n2 + 1) * 4) +
           grid_align16(static_cast<size_t>(cap_n2) * 4);
}

__device__ __forceinline__ PairArrays pairlist_carve(unsigned char *p, int cap_pairs, int cap_n1, int cap_n2) {
    PairArrays A;
    A.pair = reinterpret_cast<uint32_t *>(p); p += grid_align16(static_cast<size_t>(cap_pairs) * 4);
    A.prow = reinterpret_cast<uint16_t *>(p); p += grid_align16(static_cast<size_t>(cap_pairs) * 2);
    A.cpair = reinterpret_cast<uint16_t *>(p); p += grid_align16(static_cast<size_t>(cap_pairs) * 2);
    A.rowoff = reinterpret_cast<int32_t *>(p); p += grid_align16(static_cast<size_t>(cap_n1 + 1) * 4);
    A.colstart = reinterpret_cast<int32_t *>(p); p += grid_align16(static_cast<size_t>(cap_n2 + 1) * 4);
    A.colaux = reinterpret_cast<int32_t *>(p);
    A.cap = cap_pairs;
    return A;
}

// Inclusive block scan of v[0..n) in shared memory; thread t owns a contiguous strip.
// s_part needs blockDim.x / 32 ints.  Ends with a barrier.
__device__ __forceinline__ void block_inclusive_scan(int32_t *v, int n, int32_t *s_part) {
    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (n + T - 1) / T;
    const int lo = min(n, tid * per), hi = min(n, lo + per);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += v[i];
    int incl = sum;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const int t = __shfl_up_sync(0xFFFFFFFFu, incl, s);
        if (lane >= s) incl += t;
    }
    if (lane == 31) s_part[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int nw = T >> 5;
        int w = (lane < nw) ? s_part[lane] : 0;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const int t = __shfl_up_sync(0xFFFFFFFFu, w, s);
            if (lane >= s) w += t;
        }
        if (lane < nw) s_part[lane] = w; // inclusive over warps
    }
    __syncthreads();
    int run = incl - sum + (warp ? s_part[warp - 1] : 0);
    for (int i = lo; i < hi; ++i) {
        run += v[i];
        v[i] = run;
    }
    __syncthreads();
}

// Calls f(item slot t) for every slot of the row's window(s), one thread walking them in order.
template <class F>
__device__ __forceinline__ void for_each_slot(const RowWindows &rw, const GridJob &job, int grid_rows, F &&f) {
#pragma unroll
    for (int k = 0; k < 2; ++k) // nx[1] == 0 for points; constant indices keep the window in registers
        for (int x = rw.min_x[k]; x < rw.min_x[k] + rw.nx[k]; ++x) {
            const int lo = job.cell_start[x * grid_rows + rw.min_y[k]], hi = job.cell_start[x * grid_rows + rw.max_y[k]];
            for (int t = lo; t < hi; ++t) f(t);
        }
}

// Block-wide.  job.m12 must be readable / writable by the whole CTA (shared or global memory); *s_count
// (shared) receives accepted - culled.  Returns false when the slots do not fit (nothing written).
__device__ __forceinline__ bool pairlist_match(const GridJob &job, const GridParams &gp, const PairArrays &A,
                                               int32_t *s_part, int *s_count) {
    const int T = blockDim.x, tid = threadIdx.x;
    const int n1 = job.n1, n2 = job.n2;
    // P1
    for (int i1 = tid; i1 < n1; i1 += T) {
        const RowWindows rw = load_windows(job, gp, i1);
        int cnt = 0;
        for (int k = 0; k < rw.n_win; ++k)
            for (int x = rw.min_x[k]; x < rw.min_x[k] + rw.nx[k]; ++x)
                cnt += max(0, job.cell_start[x * gp.grid_rows + rw.max_y[k]] - job.cell_start[x * gp.grid_rows + rw.min_y[k]]);
        A.rowoff[i1 + 1] = cnt;
    }
    if (tid == 0) A.rowoff[0] = 0;
    for (int i = tid; i <= n2; i += T) A.colstart[i] = 0;
    __syncthreads();
    block_inclusive_scan(A.rowoff + 1, n1, s_part);
    const int S = A.rowoff[n1];
    if (S > A.cap) return false;
    // P2: a train feature listed in several cells of the window(s) (a line) is kept once -- the candidate
    // set is a set (matching.cpp:136 / :213); repeats stay PAIR_INVALID and cost no distance
    for (int i1 = tid; i1 < n1; i1 += T) {
        const RowQuery r = load_row(job, gp, i1);
        const int p0 = A.rowoff[i1];
        int p = p0;
        for_each_slot(r.rw, job, gp.grid_rows, [&](int t) {
            const int i2 = job.cell_items[t];
            uint32_t v = PAIR_INVALID;
            if (i2 >= 0 && i2 < n2) {
                bool dup = false;
                if (job.is_lines)
                    for (int pp = p0; pp < p; ++pp) dup = dup || (A.pair[pp] >> 9) == static_cast<uint32_t>(i2);
                if (!dup && candidate_ok(job, gp, r, i2)) {
                    const int d = hamming256(r.q, load_desc_any(job.d2, i2));
                    v = (static_cast<uint32_t>(i2) << 9) | static_cast<uint32_t>(d);
                    atomicAdd(&A.colstart[i2 + 1], 1);
                }
            }
            A.pair[p] = v;
            A.prow[p] = static_cast<uint16_t>(i1);
            ++p;
        });
    }
    __syncthreads();
    // P3
    block_inclusive_scan(A.colstart + 1, n2, s_part);
    for (int i = tid; i < n2; i += T) A.colaux[i] = A.colstart[i];
    __syncthreads();
    for (int p = tid; p < S; p += T) {
        const uint32_t v = A.pair[p];
        if (v != PAIR_INVALID) A.cpair[atomicAdd(&A.colaux[v >> 9], 1)] = static_cast<uint16_t>(p);
    }
    __syncthreads();
    // P4
    for (int p = tid; p < S; p += T) {
        const uint32_t v = A.pair[p];
        if (v == PAIR_INVALID) continue;
        const int i2 = static_cast<int>(v >> 9), d = static_cast<int>(v & 511u);
        const int i1 = A.prow[p] & ~PROW_DEAD;
        bool dead = false;
        if (gp.best_lr)
            for (int b = A.colstart[i2]; b < A.colstart[i2 + 1]; ++b) {
                const int q = A.cpair[b];
                const int i1q = A.prow[q] & ~PROW_DEAD;
                dead = dead || (i1q < i1 && static_cast<int>(A.pair[q] & 511u) <= d);
            }
        if (dead) A.prow[p] = static_cast<uint16_t>(i1 | PROW_DEAD);
    }
    if (gp.best_lr) {
        for (int i2 = tid; i2 < n2; i2 += T) {
            uint32_t best = 0xFFFFFFFFu; // D << 16 | i1
            for (int b = A.colstart[i2]; b < A.colstart[i2 + 1]; ++b) {
                const int q = A.cpair[b];
                best = min(best, ((A.pair[q] & 511u) << 16) | static_cast<uint32_t>(A.prow[q] & ~PROW_DEAD));
            }
            A.colaux[i2] = (best == 0xFFFFFFFFu) ? -1 : static_cast<int32_t>(best & 0xFFFFu);
        }
    }
    __syncthreads();
    // P5
    int accepted = 0;
    for (int i1 = tid; i1 < n1; i1 += T) {
        uint32_t b0 = KEY32_ABSENT, b1 = KEY32_ABSENT;
        for (int p = A.rowoff[i1]; p < A.rowoff[i1 + 1]; ++p) {
            const uint32_t v = A.pair[p];
            if (v != PAIR_INVALID && !(A.prow[p] & PROW_DEAD))
                top2_insert(b0, b1, ((v & 511u) << GRID_KEY_BITS) | (v >> 9));
        }
        if (b0 != KEY32_ABSENT) {
            const int best_d = static_cast<int>(b0 >> GRID_KEY_BITS);
            const int best_d2 = (b1 == KEY32_ABSENT) ? 0x7FFFFFFF : static_cast<int>(b1 >> GRID_KEY_BITS);
            if (static_cast<double>(best_d) < __dmul_rn(static_cast<double>(best_d2), gp.ratio)) {
                job.m12[i1] = static_cast<int32_t>(b0 & ((1u << GRID_KEY_BITS) - 1));
                ++accepted;
            }
        }
    }
    if (accepted) atomicAdd(s_count, accepted);
    __syncthreads();
    // P6: every entry >= 0, stale ones included
    if (gp.best_lr) {
        int culled = 0;
        for (int i1 = tid; i1 < n1; i1 += T) {
            const int32_t i2 = job.m12[i1];
            if (i2 >= 0) {
                const int back = (i2 < n2) ? A.colaux[i2] : -1;
                if (back != i1) {
                    job.m12[i1] = -1;
                    ++culled;
                }
            }
        }
        if (culled) atomicSub(s_count, culled);
        __syncthreads();
    }
    return true;
}

// ---- fused single-CTA kernel: one job per CTA (blockIdx.x), everything in one launch --------------
// dynamic smem: uint16 wmin[W][n2_max] | uint32 m21key[n2_max] | staged copies of the job's inputs
__global__ void __launch_bounds__(1024)
grid_match_fused_kernel(const GridJob *__restrict__ jobs, GridParams gp, int n2_max) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_count;
    GridJob job = jobs[blockIdx.x];
    const int W = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint16_t *wmin = reinterpret_cast<uint16_t *>(smem_raw);
    uint32_t *m21key = reinterpret_cast<uint32_t *>(smem_raw + grid_align16(static_cast<size_t>(W) * n2_max * 2));
    const int n1 = job.n1, n2 = job.n2;
    if (gp.staged) {
        const int n_cells = gp.grid_rows * gp.grid_cols;
        unsigned char *p = reinterpret_cast<unsigned char *>(m21key) + grid_align16(static_cast<size_t>(n2_max) * 4);
        const int n_items = job.cell_start[n_cells];
        unsigned char *s_cs = p; p += grid_align16(static_cast<size_t>(n_cells + 1) * 4);
        unsigned char *s_ci = p; p += grid_align16(static_cast<size_t>(gp.cap_items) * 4);
        unsigned char *s_d2 = p; p += static_cast<size_t>(n2_max) * 32;
        unsigned char *s_d1 = p; p += static_cast<size_t>(gp.cap_n1) * 32;
        unsigned char *s_co = p; p += grid_align16(static_cast<size_t>(gp.cap_n1) * 16);
        unsigned char *s_dir = p;
        stage_bytes(s_cs, job.cell_start, static_cast<size_t>(n_cells + 1) * 4);
        stage_bytes(s_ci, job.cell_items, static_cast<size_t>(n_items) * 4);
        stage_bytes(s_d2, job.d2, static_cast<size_t>(n2) * 32);
        stage_bytes(s_d1, job.d1, static_cast<size_t>(n1) * 32);
        stage_bytes(s_co, job.coords, static_cast<size_t>(n1) * (job.is_lines ? 16 : 8));
        if (job.is_lines) stage_bytes(s_dir, job.dirs2, static_cast<size_t>(n2) * 16);
        job.cell_start = reinterpret_cast<const int32_t *>(s_cs);
        job.cell_items = reinterpret_cast<const int32_t *>(s_ci);
        job.d2 = reinterpret_cast<const uint4 *>(s_d2);
        job.d1 = reinterpret_cast<const uint4 *>(s_d1);
        job.coords = reinterpret_cast<const int32_t *>(s_co);
        if (job.is_lines) job.dirs2 = reinterpret_cast<const double *>(s_dir);
    }
    if (threadIdx.x == 0) s_count = 0;
    for (int i = threadIdx.x; i < W * n2_max; i += blockDim.x) wmin[i] = D_INF;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) m21key[i] = KEY32_ABSENT;
    __syncthreads();

    const int rpw = (n1 + W - 1) / W;
    const int row0 = min(n1, warp * rpw), row1 = min(n1, row0 + rpw);
    uint16_t *mine = wmin + static_cast<size_t>(warp) * n2_max;

    if (gp.best_lr) {
        chunk_minima(job, gp, row0, row1, mine, lane);
        __syncthreads();
        // exclusive prefix-min over the chunks, per column
        for (int i2 = threadIdx.x; i2 < n2; i2 += blockDim.x) {
            uint16_t run = D_INF;
            for (int w = 0; w < W; ++w) {
                const uint16_t t = wmin[static_cast<size_t>(w) * n2_max + i2];
                wmin[static_cast<size_t>(w) * n2_max + i2] = run;
                run = min(run, t);
            }
        }
        __syncthreads();
    }

    const int acc = chunk_match(job, gp, row0, row1, mine, lane, [&](int i2, int d, int i1) {
        atomicMin(&m21key[i2], (static_cast<uint32_t>(d) << GRID_KEY_BITS) | static_cast<uint32_t>(i1));
    });
    if (lane == 0 && acc) atomicAdd(&s_count, acc);
    __syncthreads();

    if (gp.best_lr) {
        // mutual check (matching.cpp:166-174) over every entry >= 0, stale ones included
        int culled = 0;
        for (int i1 = threadIdx.x; i1 < n1; i1 += blockDim.x) {
            const int32_t i2 = job.m12[i1];
            if (i2 >= 0) {
                const uint32_t k = (i2 < n2) ? m21key[i2] : KEY32_ABSENT;
                const int back = (k == KEY32_ABSENT) ? -1 : static_cast<int>(k & ((1u << GRID_KEY_BITS) - 1));
                if (back != i1) {
                    job.m12[i1] = -1;
                    ++culled;
                }
            }
        }
        if (culled) atomicSub(&s_count, culled);
        __syncthreads();
    }
    if (threadIdx.x == 0) *job.count = s_count;
}

// ---- cluster kernel: ONE frame-sized job spread over a thread-block cluster of CLUSTER CTAs -----------
// A single call must not sit on one SM of 148: the chunks (warps) of the job are dealt to the CTAs of a
// cluster in row order, every CTA keeps its chunk arrays in its own shared memory, and the two places
// where chunks talk to each other go through distributed shared memory:
//   * threshold seeding: CTA r reads the per-CTA column minima of the CTAs r' < r (their rows come first);
//   * mutual check: the best live pair of a column is the minimum over the CTAs' m21key arrays.
// grid = CLUSTER * n_jobs CTAs; job = blockIdx.x / CLUSTER.  *job.count must be zero on entry.
// dynamic smem per CTA: uint16 wmin[W][n2_max] | uint16 ctamin[n2_max] | uint32 m21key[n2_max] | staged
__host__ __device__ inline size_t grid_cluster_smem(int warps, int n2_max, int staged, int rows_per_cta, int cap_items,
                                                    int n_cells, bool any_lines) {
    size_t b = grid_align16(static_cast<size_t>(warps) * n2_max * 2) + grid_align16(static_cast<size_t>(n2_max) * 2) +
               grid_align16(static_cast<size_t>(n2_max) * 4);
    if (staged) {
        b += grid_align16(static_cast<size_t>(n_cells + 1) * 4) + grid_align16(static_cast<size_t>(cap_items) * 4);
        b += static_cast<size_t>(n2_max) * 32 + static_cast<size_t>(rows_per_cta) * 32;
        b += grid_align16(static_cast<size_t>(rows_per_cta) * 16);
        if (any_lines) b += static_cast<size_t>(n2_max) * 16;
    }
    return b;
}

template <int CLUSTER>
__global__ void __launch_bounds__(512)
grid_match_cluster_kernel(const GridJob *__restrict__ jobs, GridParams gp, int n2_max) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = static_cast<int>(cluster.block_rank());
    GridJob job = jobs[blockIdx.x / CLUSTER];
    const int W = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n1 = job.n1, n2 = job.n2;
    uint16_t *wmin = reinterpret_cast<uint16_t *>(smem_raw);
    uint16_t *ctamin = reinterpret_cast<uint16_t *>(smem_raw + grid_align16(static_cast<size_t>(W) * n2_max * 2));
    uint32_t *m21key = reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(ctamin) + grid_align16(static_cast<size_t>(n2_max) * 2));

    // chunk g = rank * W + warp owns rows [g * rpw, (g + 1) * rpw): CTA order == row order
    const int rpw = (n1 + CLUSTER * W - 1) / (CLUSTER * W);
    const int cta_row0 = min(n1, rank * W * rpw), cta_row1 = min(n1, cta_row0 + W * rpw);
    const int row0 = min(n1, cta_row0 + warp * rpw), row1 = min(n1, row0 + rpw);

    for (int i = threadIdx.x; i < W * n2_max; i += blockDim.x) wmin[i] = D_INF;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) m21key[i] = KEY32_ABSENT;
    if (gp.staged) {
        const int n_cells = gp.grid_rows * gp.grid_cols;
        unsigned char *p = reinterpret_cast<unsigned char *>(m21key) + grid_align16(static_cast<size_t>(n2_max) * 4);
        const int n_items = job.cell_start[n_cells];
        unsigned char *s_cs = p; p += grid_align16(static_cast<size_t>(n_cells + 1) * 4);
        unsigned char *s_ci = p; p += grid_align16(static_cast<size_t>(gp.cap_items) * 4);
        unsigned char *s_d2 = p; p += static_cast<size_t>(n2_max) * 32;
        unsigned char *s_d1 = p; p += static_cast<size_t>(gp.cap_n1) * 32;
        unsigned char *s_co = p; p += grid_align16(static_cast<size_t>(gp.cap_n1) * 16);
        unsigned char *s_dir = p;
        const int cpq = job.is_lines ? 4 : 2;
        stage_bytes(s_cs, job.cell_start, static_cast<size_t>(n_cells + 1) * 4);
        stage_bytes(s_ci, job.cell_items, static_cast<size_t>(n_items) * 4);
        stage_bytes(s_d2, job.d2, static_cast<size_t>(n2) * 32);
        stage_bytes(s_d1, job.d1 + 2 * static_cast<size_t>(cta_row0), static_cast<size_t>(cta_row1 - cta_row0) * 32);
        stage_bytes(s_co, job.coords + static_cast<size_t>(cta_row0) * cpq, static_cast<size_t>(cta_row1 - cta_row0) * cpq * 4);
        if (job.is_lines) stage_bytes(s_dir, job.dirs2, static_cast<size_t>(n2) * 16);
        job.cell_start = reinterpret_cast<const int32_t *>(s_cs);
        job.cell_items = reinterpret_cast<const int32_t *>(s_ci);
        job.d2 = reinterpret_cast<const uint4 *>(s_d2);
        job.d1 = reinterpret_cast<const uint4 *>(s_d1);
        job.coords = reinterpret_cast<const int32_t *>(s_co);
        job.q_row_base = cta_row0;
        if (job.is_lines) job.dirs2 = reinterpret_cast<const double *>(s_dir);
    }
    __syncthreads();

    uint16_t *mine = wmin + static_cast<size_t>(warp) * n2_max;
    if (gp.best_lr) {
        chunk_minima(job, gp, row0, row1, mine, lane);
        __syncthreads();
        // exclusive prefix-min over this CTA's chunks; the CTA total goes to ctamin for the higher ranks
        for (int i2 = threadIdx.x; i2 < n2; i2 += blockDim.x) {
            uint16_t run = D_INF;
            for (int w = 0; w < W; ++w) {
                const uint16_t t = wmin[static_cast<size_t>(w) * n2_max + i2];
                wmin[static_cast<size_t>(w) * n2_max + i2] = run;
                run = min(run, t);
            }
            ctamin[i2] = run;
        }
        cluster.sync();
        if (rank > 0) {
            for (int i2 = threadIdx.x; i2 < n2; i2 += blockDim.x) {
                uint16_t seed = D_INF;
                for (int r = 0; r < rank; ++r) seed = min(seed, cluster.map_shared_rank(ctamin, r)[i2]);
                if (seed != D_INF)
                    for (int w = 0; w < W; ++w) {
                        uint16_t &t = wmin[static_cast<size_t>(w) * n2_max + i2];
                        t = min(t, seed);
                    }
            }
        }
        __syncthreads();
    }

    const int acc = chunk_match(job, gp, row0, row1, mine, lane, [&](int i2, int d, int i1) {
        atomicMin(&m21key[i2], (static_cast<uint32_t>(d) << GRID_KEY_BITS) | static_cast<uint32_t>(i1));
    });
    if (lane == 0 && acc) atomicAdd(job.count, acc);

    if (gp.best_lr) {
        cluster.sync();
        // mutual check (matching.cpp:166-174) over this CTA's rows, stale entries included
        int culled = 0;
        for (int i1 = cta_row0 + threadIdx.x; i1 < cta_row1; i1 += blockDim.x) {
            const int32_t i2 = job.m12[i1];
            if (i2 >= 0) {
                uint32_t k = KEY32_ABSENT;
                if (i2 < n2)
                    for (int r = 0; r < CLUSTER; ++r) k = min(k, cluster.map_shared_rank(m21key, r)[i2]);
                const int back = (k == KEY32_ABSENT) ? -1 : static_cast<int>(k & ((1u << GRID_KEY_BITS) - 1));
                if (back != i1) {
                    job.m12[i1] = -1;
                    ++culled;
                }
            }
        }
        const unsigned msk = __ballot_sync(0xFFFFFFFFu, culled != 0);
        (void)msk;
        if (culled) atomicSub(job.count, culled);
    }
    cluster.sync(); // nobody leaves while its shared memory may still be read
}

// ---- chunked launch for map-sized jobs: CTA c owns rows [c*W*rpw, (c+1)*W*rpw) --------------------
// pass 0: per-CTA column minima -> gp.cta_min[c][:]
// (then grid_scan_kernel turns cta_min into exclusive thresholds)
// pass 1: phase A again in shared memory, seeded prefix over the warps, phase C, accepts counted
// dynamic smem: uint16 wmin[W][n2]
template <int PASS>
__global__ void __launch_bounds__(1024)
grid_match_chunked_kernel(GridJob job, GridParams gp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int W = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n1 = job.n1, n2 = job.n2;
    uint16_t *wmin = reinterpret_cast<uint16_t *>(smem_raw);
    for (int i = threadIdx.x; i < W * n2; i += blockDim.x) wmin[i] = D_INF;
    __syncthreads();
    const long long cta_row0 = static_cast<long long>(blockIdx.x) * W * gp.rows_per_warp;
    const int row0 = static_cast<int>(min(static_cast<long long>(n1), cta_row0 + static_cast<long long>(warp) * gp.rows_per_warp));
    const int row1 = min(n1, row0 + gp.rows_per_warp);
    uint16_t *mine = wmin + static_cast<size_t>(warp) * n2;
    uint16_t *cta_min = gp.cta_min + static_cast<size_t>(blockIdx.x) * n2;

    if (gp.best_lr) {
        chunk_minima(job, gp, row0, row1, mine, lane);
        __syncthreads();
        if (PASS == 0) {
            for (int i2 = threadIdx.x; i2 < n2; i2 += blockDim.x) {
                uint16_t run = D_INF;
                for (int w = 0; w < W; ++w) run = min(run, wmin[static_cast<size_t>(w) * n2 + i2]);
                cta_min[i2] = run;
            }
            return;
        }
        for (int i2 = threadIdx.x; i2 < n2; i2 += blockDim.x) {
            uint16_t run = cta_min[i2]; // threshold inherited from all earlier CTAs (and lower shards)
            for (int w = 0; w < W; ++w) {
                const uint16_t t = wmin[static_cast<size_t>(w) * n2 + i2];
                wmin[static_cast<size_t>(w) * n2 + i2] = run;
                run = min(run, t);
            }
        }
        __syncthreads();
    }
    if (PASS == 1) {
        const int acc = chunk_match(job, gp, row0, row1, mine, lane, [&](int i2, int d, int i1) {
            atomicMin(&gp.m21key[i2], make_key64(static_cast<uint32_t>(d), static_cast<uint32_t>(job.i1_base + i1)));
        });
        if (lane == 0 && acc) atomicAdd(job.count, acc);
    }
}

// ---- row-parallel launch for map-sized jobs ---------------------------------------------------------------------
// The warp-per-chunk kernels above replay the rows of a chunk in order; at map scale (200 000 rows x ~10 candidates)
// that serial dependent chain (coords -> cell_start -> cell_items -> descriptor) is what the time goes into.  Here
// a CTA takes blocks of 256 consecutive rows and turns each block into a FLAT PAIR LIST in shared memory:
//   phase A  thread = row: window clamp, slot count, block scan, then the row's (row, i2) entries are written out
//   phase B  thread = entry (stride 256, no divergence): direction filter, distance, and
//              pass 0:  cmin[i2] = min D                                          -> cta_min[c][:]
//              pass 1:  D < T[i2] ("survivor": it beats every earlier block, CTA and lower-ranked shard) ?
//                       the entry stays in the list as (i2, row, D) and proposes K[i2] = min (D << 16 | row)
//   rounds   among the survivors of a block the live pairs are the strict prefix minima of their column in row
//            order: the entry that owns K[i2] is live, later rows of the column are dead, earlier rows stay
//            undecided and propose again (one record of the column's chain per round, ~log of the survivors per
//            column rounds; after the first CTAs a block has almost no survivors).  Live pairs go into the row's
//            top-2 with a linearisable shared-memory insert; fp64 ratio test and the m21 key follow as before.
//   scan     between the passes grid_scan_kernel turns cta_min into exclusive thresholds (+ seed of lower shards).
// A train feature listed in several cells of a row's window(s) (a line) yields repeated entries: every update is
// an idempotent min and the row top-2 ignores a key it already holds, so no de-duplication pass is needed.
// A block whose slots exceed the list capacity (very dense windows) is handled by the same rounds with every
// thread re-walking its own row instead (slower, same result).
constexpr int GRID_ROW_THREADS = 256;
constexpr int GRID_SEG_TAB = 32;          // list segments a CTA can hand from pass 0 to pass 1
constexpr int GRID_CLUSTER_MAX_CTAS = 8;   // CTAs of the one-launch cluster form (MODE 2)
constexpr int GRID_ENT_PER_ROW = 64;      // hand-over capacity per row (slots; ~10 are used for points, ~50 for lines)

struct GridRowsSmem { // layout of the work area, in bytes from the start of dynamic shared memory
    size_t K, K2, T, Tnew, B, cmin, rb0, rb1, d1s, rowdir, ent, stage, total;
};
__host__ __device__ inline GridRowsSmem grid_rows_layout(int n2, bool lines, int cap_pairs, int n_cells, int cap_items, int staged) {
    GridRowsSmem L;
    size_t o = 0;
    L.K = o; o += grid_align16(static_cast<size_t>(n2) * 4);
    L.K2 = o; o += grid_align16(static_cast<size_t>(n2) * 4); // second proposal array: the rounds over a pair list ping-pong
    L.T = o; o += grid_align16(static_cast<size_t>(n2) * 2);
    L.Tnew = o; o += grid_align16(static_cast<size_t>(n2) * 2);
    L.B = o; o += grid_align16(static_cast<size_t>(n2) * 2);
    L.cmin = o; o += grid_align16(static_cast<size_t>(n2) * 2); // cluster form: this CTA's column minima, read by the higher ranks
    L.rb0 = o; o += static_cast<size_t>(GRID_ROW_THREADS) * 4;
    L.rb1 = o; o += static_cast<size_t>(GRID_ROW_THREADS) * 4;
    L.d1s = o; o += static_cast<size_t>(GRID_ROW_THREADS) * 32;
    L.rowdir = o; o += lines ? static_cast<size_t>(GRID_ROW_THREADS) * 16 : 0;
    L.ent = o; o += grid_align16(static_cast<size_t>(cap_pairs) * 4);
    L.stage = o;
    if (staged)
        o += grid_align16(static_cast<size_t>(n_cells + 1) * 4) + static_cast<size_t>(n2) * 32 +
             (lines ? static_cast<size_t>(n2) * 16 : 0) + grid_align16(static_cast<size_t>(cap_items) * 4);
    L.total = o;
    return L;
}

// Concurrent insert of key k into the sorted pair (b0 <= b1) held in shared memory.  Linearisable: the insert that
// lowers b0 hands the previous minimum down to b1, every later insert sees the new minimum and hands itself down, so
// b1 ends as the second smallest DISTINCT key (a repeated key is ignored).
__device__ __forceinline__ void top2_insert_atomic(uint32_t *b0, uint32_t *b1, uint32_t k) {
    const uint32_t old = atomicMin(b0, k);
    if (old != k) atomicMin(b1, max(old, k));
}

// f(i2, D) for every candidate of the row that passes the range check and the direction filter.
template <class F>
__device__ __forceinline__ void row_walk(const GridJob &job, const GridParams &gp, const RowQuery &r, F &&f) {
    for_each_slot(r.rw, job, gp.grid_rows, [&](int t) {
        const int i2 = job.cell_items[t];
        if (candidate_ok(job, gp, r, i2)) f(i2, hamming256(r.q, load_desc_any(job.d2, i2)));
    });
}

// STAGED: the frame side (cell_start, d2, dirs2 and -- when they fit -- cell_items) is copied into shared memory first;
// the accesses below then go through pointers the compiler can prove to be shared (LDS instead of generic loads).
// MODE 0 / 1: pass 0 / pass 1 of the multi-launch form (thresholds travel through gp.cta_min and grid_scan_kernel).
// MODE 2: the whole matchGrid of a frame-sized job in ONE launch -- the CTAs (256 rows each, <= 8) form one
//         thread-block cluster: pass 0, cluster barrier, thresholds = minima of the lower-ranked CTAs read through
//         distributed shared memory, pass 1, cluster barrier, mutual check from the global m21 keys.
template <int STAGED, int MODE>
__device__ __forceinline__ void grid_rows_device(GridJob job, GridParams gp, const int cta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_warp_tot[GRID_ROW_THREADS / 32];
    __shared__ int s_seg_end;
    const int tid = threadIdx.x, NT = GRID_ROW_THREADS, lane = tid & 31, warp = tid >> 5;
    if (MODE != 2) PLM_TL(MODE == 1 ? 15 : 0);
    const int n1 = job.n1, n2 = job.n2;
    const int n_cells = gp.grid_rows * gp.grid_cols;
    const GridRowsSmem L = grid_rows_layout(n2, job.is_lines != 0, gp.cap_pairs, n_cells, gp.cap_items, STAGED);
    uint32_t *K = reinterpret_cast<uint32_t *>(smem_raw + L.K);
    uint32_t *K2 = reinterpret_cast<uint32_t *>(smem_raw + L.K2);
    uint16_t *T = reinterpret_cast<uint16_t *>(smem_raw + L.T);
    uint16_t *Tnew = reinterpret_cast<uint16_t *>(smem_raw + L.Tnew);
    uint16_t *B = reinterpret_cast<uint16_t *>(smem_raw + L.B);
    uint32_t *rb0 = reinterpret_cast<uint32_t *>(smem_raw + L.rb0), *rb1 = reinterpret_cast<uint32_t *>(smem_raw + L.rb1);
    uint4 *d1s = reinterpret_cast<uint4 *>(smem_raw + L.d1s);
    double2 *rowdir = reinterpret_cast<double2 *>(smem_raw + L.rowdir);
    uint32_t *ent = reinterpret_cast<uint32_t *>(smem_raw + L.ent);
    const size_t o_cs = L.stage, o_d2 = o_cs + grid_align16(static_cast<size_t>(n_cells + 1) * 4),
                 o_dir = o_d2 + static_cast<size_t>(n2) * 32, o_ci = o_dir + (job.is_lines ? static_cast<size_t>(n2) * 16 : 0);
    const int32_t *cs = STAGED ? reinterpret_cast<const int32_t *>(smem_raw + o_cs) : job.cell_start;
    const uint4 *d2p = STAGED ? reinterpret_cast<const uint4 *>(smem_raw + o_d2) : job.d2;
    const double2 *dirp = STAGED ? reinterpret_cast<const double2 *>(smem_raw + o_dir) : reinterpret_cast<const double2 *>(job.dirs2);
    const int32_t *s_items = reinterpret_cast<const int32_t *>(smem_raw + o_ci);
    bool items_staged = false;
    // pass 1 in its replay form (every pair of this CTA was handed over by pass 0) never touches the frame side
    const bool replay_only = MODE == 1 && gp.best_lr && gp.ent_g && gp.seg_cnt[cta] >= 0;
    if (MODE == 0 && gp.init_flags) { // plm_dev_match_grid: no separate memsets in front of the pass
        if (cta == 0) {
            if ((gp.init_flags & 4) && gp.m21key)
                for (int i = tid; i < n2; i += NT) gp.m21key[i] = KEY64_ABSENT;
            if ((gp.init_flags & 2) && tid == 0) *job.count = 0;
        }
    }
    if (MODE == 2) { // the rows this thread will load in phase A: on their way into L2 while the frame side is staged
        const long long i1 = static_cast<long long>(cta) * gp.rows_per_cta + tid;
        if (tid < gp.rows_per_cta && i1 < n1) {
            prefetch_l2(job.d1 + 2 * (i1 - job.q_row_base));
            prefetch_l2(job.coords + (job.is_lines ? 4 : 2) * (i1 - job.q_row_base));
        }
    }
    if (STAGED && !replay_only) {
        stage_bytes_async(smem_raw + o_cs, job.cell_start, static_cast<size_t>(n_cells + 1) * 4);
        stage_bytes_async(smem_raw + o_d2, job.d2, static_cast<size_t>(n2) * 32);
        if (job.is_lines) stage_bytes_async(smem_raw + o_dir, job.dirs2, static_cast<size_t>(n2) * 16);
        const int n_items = gp.n_items_p1 > 0 ? gp.n_items_p1 - 1 : job.cell_start[n_cells];
        items_staged = n_items <= gp.cap_items;
        if (items_staged) stage_bytes_async(smem_raw + o_ci, job.cell_items, static_cast<size_t>(max(n_items, 0)) * 4);
        stage_wait(); // visible to the CTA after the barrier that opens the first pass
        // the re-walk form (list overflow) goes through the job's pointers
        job.cell_start = cs;
        job.d2 = d2p;
        if (job.is_lines) job.dirs2 = reinterpret_cast<const double *>(dirp);
        if (items_staged) job.cell_items = s_items;
    }
    PLM_TL(MODE == 1 ? 16 : 1);
    // the first head_ctas CTAs take head_rows rows each (short blocks at the start of the map, see GridParams)
    const long long cta_row0 = cta < gp.head_ctas ? static_cast<long long>(cta) * gp.head_rows
                                                  : static_cast<long long>(gp.head_ctas) * gp.head_rows +
                                                        static_cast<long long>(cta - gp.head_ctas) * gp.rows_per_cta;
    const int row_end = static_cast<int>(min(static_cast<long long>(n1), cta_row0 + (cta < gp.head_ctas ? gp.head_rows : gp.rows_per_cta)));
    if (MODE == 0 && (gp.init_flags & 1))
        for (long long i = cta_row0 + tid; i < row_end; i += NT) job.m12[i] = -1;
    uint16_t *cta_min = MODE == 2 ? nullptr : gp.cta_min + static_cast<size_t>(cta) * n2;
    uint16_t *cmin16 = reinterpret_cast<uint16_t *>(smem_raw + L.cmin);

    // MODE 2 with a single block of rows per CTA: the pair list pass 0 evaluated stays in shared memory (entries
    // re-encoded as i2 | row | D) and pass 1 only streams it against its thresholds; -1 = not available
    int smem_rec_total = -1;
    // one pass over this CTA's rows; P = std::integral_constant<int, 0 / 1>
    auto run_pass = [&](auto P, uint16_t *cmin_out, const uint16_t *thr_in) {
    constexpr int PASS = decltype(P)::value;
    const bool thresholds = PASS == 1 && gp.best_lr;

    if (PASS == 0) {
        for (int i = tid; i < n2; i += NT) K[i] = KEY32_ABSENT;
    } else if (thresholds) {
        for (int i = tid; i < n2; i += NT) {
            const uint16_t t = thr_in[i]; // exclusive prefix over the earlier CTAs (and lower-ranked shards)
            T[i] = t;
            Tnew[i] = t;
            B[i] = D_INF;
            K[i] = KEY32_ABSENT;
            K2[i] = KEY32_ABSENT;
        }
    }
    __syncthreads();
    if (PASS == 1) PLM_TL(7);
    const uint32_t utid = static_cast<uint32_t>(tid);
    // Rounds over the survivors in ent[0, seg_total) (entry = i2 << 17 | row << 9 | D; the proposals of round 0 are
    // already in K).  Per round ONE pass over the list: the entry that owns the column's minimum key is live, later
    // rows of the column are dead, earlier rows stay undecided and propose straight into the other array for the next
    // round; then the touched columns are folded (round 0: new threshold + m21 candidate) and cleared.
    auto list_rounds = [&](int seg_total, long long base) {
        uint32_t *Kc = K, *Kn = K2;
        for (int round = 0;; ++round) {
            bool again = false;
            for (int e = tid; e < seg_total; e += NT) {
                const uint32_t v = ent[e];
                if (v == PAIR_INVALID) continue;
                const uint32_t i2 = v >> 17, row = (v >> 9) & 0xFFu, d = v & 0x1FFu;
                const uint32_t key = (d << 16) | row, k = Kc[i2];
                if (key == k) {
                    top2_insert_atomic(&rb0[row], &rb1[row], (d << GRID_KEY_BITS) | i2);
                    ent[e] = PAIR_INVALID;
                } else if (row > (k & 0xFFFFu)) {
                    ent[e] = PAIR_INVALID;
                } else {
                    atomicMin(&Kn[i2], key);
                    again = true;
                }
            }
            const int more = __syncthreads_or(again ? 1 : 0);
            for (int i2 = tid; i2 < n2; i2 += NT) {
                const uint32_t k = Kc[i2];
                if (k != KEY32_ABSENT) {
                    if (round == 0) { // the best pair of the column in this block / segment: new threshold, m21 candidate
                        Tnew[i2] = static_cast<uint16_t>(k >> 16);
                        atomicMin(&gp.m21key[i2], make_key64(k >> 16, static_cast<uint32_t>(job.i1_base + base + (k & 0xFFFFu))));
                    }
                    Kc[i2] = KEY32_ABSENT;
                }
            }
            __syncthreads();
            if (!more) break;
            uint32_t *t = Kc;
            Kc = Kn;
            Kn = t;
        }
    };
    int accepted = 0;
    // pass 0 records its pairs for pass 1 (MODE 0 / 1 only; the one-launch cluster form keeps everything on chip)
    const bool rec_smem = PASS == 0 && MODE == 2 && row_end - cta_row0 <= NT;
    const bool recording = (PASS == 0 && MODE == 0 && gp.ent_g != nullptr) || rec_smem;
    const bool replay_smem = PASS == 1 && MODE == 2 && thresholds && smem_rec_total >= 0;
    uint32_t *rec_base = (recording && !rec_smem) || (PASS == 1 && MODE == 1 && gp.ent_g) ? gp.ent_g + static_cast<size_t>(cta) * gp.ent_per_cta : nullptr;
    int4 *rec_tab = gp.seg_tab ? gp.seg_tab + static_cast<size_t>(cta) * GRID_SEG_TAB : nullptr;
    bool rec_ok = true;   // uniform
    int rec_n = 0, rec_used = 0;
    const int n_rec = (PASS == 1 && MODE == 1 && thresholds && gp.ent_g) ? gp.seg_cnt[cta] : -1;
    int rec_i = 0;
    for (long long base = cta_row0; base < row_end; base += NT) { // uniform over the CTA
        const long long i1 = base + tid;
        const bool has_row = i1 < row_end;
        const int blk = static_cast<int>((base - cta_row0) / NT);
        if (n_rec >= 0 || replay_smem) {
            // ---- pass 1, replay form: stream the pairs pass 0 left behind against the thresholds ----
            rb0[tid] = KEY32_ABSENT;
            rb1[tid] = KEY32_ABSENT;
            __syncthreads();
            bool smem_pending = replay_smem;
            while (smem_pending || (rec_i < n_rec && rec_tab[rec_i].z == blk)) {
                const uint32_t *src = ent;
                int seg_total = smem_rec_total;
                if (smem_pending) {
                    smem_pending = false;
                } else {
                    const int4 rec = rec_tab[rec_i];
                    ++rec_i;
                    src = rec_base + rec.x;
                    seg_total = rec.y;
                }
                // four entries per trip, loaded before any of them is processed: the list comes from global memory (L2)
                // and the stores below would otherwise serialise the loads behind them
                for (int e0 = tid; e0 < seg_total; e0 += 4 * NT) {
                    uint32_t v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) v[u] = (e0 + u * NT < seg_total) ? src[e0 + u * NT] : PAIR_INVALID;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (e0 + u * NT >= seg_total) break;
                        if (v[u] != PAIR_INVALID) {
                            const uint32_t i2 = v[u] >> 17, d = v[u] & 0x1FFu;
                            if (d < T[i2]) atomicMin(&K[i2], (d << 16) | ((v[u] >> 9) & 0xFFu));
                            else v[u] = PAIR_INVALID;
                        }
                        ent[e0 + u * NT] = v[u];
                    }
                }
                __syncthreads();
                PLM_TL(8);
                list_rounds(seg_total, base);
                PLM_TL(9);
                for (int i2 = tid; i2 < n2; i2 += NT) { // thresholds for the rows that follow
                    T[i2] = Tnew[i2];
                    B[i2] = D_INF;
                }
                __syncthreads();
            }
            const uint32_t b0 = rb0[tid], b1 = rb1[tid];
            __syncthreads(); // rb / ent are rewritten by the next block
            if (has_row && b0 != KEY32_ABSENT) {
                const int best_d = static_cast<int>(b0 >> GRID_KEY_BITS);
                const int best_d2 = (b1 == KEY32_ABSENT) ? 0x7FFFFFFF : static_cast<int>(b1 >> GRID_KEY_BITS);
                if (static_cast<double>(best_d) < __dmul_rn(static_cast<double>(best_d2), gp.ratio)) { // matching.cpp:160 / :241
                    job.m12[i1] = static_cast<int32_t>(b0 & ((1u << GRID_KEY_BITS) - 1));
                    ++accepted;
                }
            }
            continue;
        }
        // ---- phase A: windows, slot counts, block scan ----
        RowQuery r;
        int S = 0;
        if (has_row) {
            r = load_row(job, gp, static_cast<int>(i1));
#pragma unroll
            for (int k = 0; k < 2; ++k) // nx[1] == 0 for points
                for (int x = r.rw.min_x[k]; x < r.rw.min_x[k] + r.rw.nx[k]; ++x)
                    S += max(0, cs[x * gp.grid_rows + r.rw.max_y[k]] - cs[x * gp.grid_rows + r.rw.min_y[k]]);
            d1s[2 * tid] = r.q.lo;
            d1s[2 * tid + 1] = r.q.hi;
            if (job.is_lines) rowdir[tid] = make_double2(r.vx, r.vy);
        }
        int incl = S;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const int v = __shfl_up_sync(0xFFFFFFFFu, incl, s);
            if (lane >= s) incl += v;
        }
        if (lane == 31) s_warp_tot[warp] = incl;
        if (PASS == 1) {
            rb0[tid] = KEY32_ABSENT;
            rb1[tid] = KEY32_ABSENT;
        }
        __syncthreads();
        if (PASS == 0) PLM_TL(2);
        int off = incl - S, total = 0;
#pragma unroll
        for (int w = 0; w < GRID_ROW_THREADS / 32; ++w) {
            const int t = s_warp_tot[w];
            if (w < warp) off += t;
            total += t;
        }
        // The block goes through the pair list in one piece when its slots fit, else as consecutive row ranges
        // ("segments": a row joins the segment its first slot falls into, every segment holds <= cap entries as long as
        // no single row has more than cap / 2 slots); segments are processed in row order like blocks.
        const int cap = gp.cap_pairs, half = cap >> 1;
        bool listed = total <= cap;
        int n_seg = 1;
        if (!listed && !__syncthreads_or(S > half ? 1 : 0)) {
            listed = true;
            n_seg = (total + half - 1) / half;
        }
        uint32_t b0 = KEY32_ABSENT, b1 = KEY32_ABSENT; // register top-2 of the re-walk form
        auto take = [&](int i2, int d) {
            const uint32_t k2 = (static_cast<uint32_t>(d) << GRID_KEY_BITS) | static_cast<uint32_t>(i2);
            if (k2 != b0 && k2 != b1) top2_insert(b0, b1, k2);
        };
        auto fold = [&](int round) { // per column: this round's record becomes the bound of the next one
            for (int i2 = tid; i2 < n2; i2 += NT) {
                const uint32_t k = K[i2];
                if (k != KEY32_ABSENT) {
                    if (round == 0) { // the best pair of the column in this block / segment: new threshold, m21 candidate
                        Tnew[i2] = static_cast<uint16_t>(k >> 16);
                        atomicMin(&gp.m21key[i2], make_key64(k >> 16, static_cast<uint32_t>(job.i1_base + base + (k & 0xFFFFu))));
                    }
                    B[i2] = static_cast<uint16_t>(k & 0xFFFFu);
                    K[i2] = KEY32_ABSENT;
                }
            }
        };
        auto next_thresholds = [&]() {
            for (int i2 = tid; i2 < n2; i2 += NT) {
                T[i2] = Tnew[i2];
                B[i2] = D_INF;
            }
        };
        bool und = false;
        if (listed) {
            auto entry_distance = [&](uint32_t v, uint32_t &row, uint32_t &i2) -> uint32_t { // 0xFFFF = rejected
                if (v == PAIR_INVALID) return 0xFFFFu;
                row = v >> 16;
                i2 = v & 0xFFFFu;
                if (thresholds && T[i2] == 0) return 0xFFFFu; // column without a live pair in this CTA (grid_scan_kernel)
                if (job.is_lines) {
                    const double2 q = rowdir[row];
                    const double2 t2 = dirp[i2];
                    const double dp = __dadd_rn(__dmul_rn(q.x, t2.x), __dmul_rn(q.y, t2.y));
                    if (fabs(dp) < gp.line_sim_th) return 0xFFFFu; // matching.cpp:221, NaN passes
                }
                Desc a;
                a.lo = d1s[2 * row];
                a.hi = d1s[2 * row + 1];
                return static_cast<uint32_t>(hamming256(a, d2p[2 * i2], d2p[2 * i2 + 1]));
            };
            auto entry_apply = [&](uint32_t d, uint32_t row, uint32_t i2) -> uint32_t { // survivor encoding or INVALID
                if (d == 0xFFFFu) return PAIR_INVALID;
                if (PASS == 0) {
                    if (d < K[i2]) atomicMin(&K[i2], d);
                } else if (!gp.best_lr) {
                    top2_insert_atomic(&rb0[row], &rb1[row], (d << GRID_KEY_BITS) | i2);
                } else if (d < T[i2]) {
                    atomicMin(&K[i2], (d << 16) | row);
                    return (i2 << 17) | (row << 9) | d;
                }
                return PAIR_INVALID;
            };
            for (int seg = 0; seg < n_seg; ++seg) {
                bool mine = has_row;
                int p = off, seg_total = total;
                if (n_seg > 1) {
                    for (int e = tid; e < half; e += NT) ent[e] = PAIR_INVALID; // slots below the segment's first row
                    if (tid == 0) s_seg_end = 0;
                    __syncthreads();
                    mine = has_row && S > 0 && off / half == seg;
                    p = off - seg * half;
                    if (mine) atomicMax(&s_seg_end, p + S);
                }
                if (mine) {
                    auto emit = [&](const int32_t *items) {
#pragma unroll
                        for (int k = 0; k < 2; ++k)
                            for (int x = r.rw.min_x[k]; x < r.rw.min_x[k] + r.rw.nx[k]; ++x) {
                                const int lo = cs[x * gp.grid_rows + r.rw.min_y[k]], hi = cs[x * gp.grid_rows + r.rw.max_y[k]];
                                for (int t = lo; t < hi; ++t) {
                                    const int i2 = items[t];
                                    ent[p++] = (i2 >= 0 && i2 < n2) ? ((utid << 16) | static_cast<uint32_t>(i2)) : PAIR_INVALID;
                                }
                            }
                    };
                    if (STAGED && items_staged) emit(s_items);
                    else emit(job.cell_items);
                }
                __syncthreads();
                if (PASS == 0) PLM_TL(3);
                if (n_seg > 1) seg_total = s_seg_end;
                // pass 0: room for this segment's pairs in the CTA's hand-over region (uniform decision)
                uint32_t *rec_dst = nullptr;
                if (rec_smem) {
                    if (n_seg == 1) { // in place: every thread rewrites the entries it has just read
                        rec_dst = ent;
                        smem_rec_total = seg_total;
                    }
                } else if (recording && rec_ok) {
                    if (rec_n < GRID_SEG_TAB && rec_used + seg_total <= gp.ent_per_cta) {
                        rec_dst = rec_base + rec_used;
                        if (tid == 0) rec_tab[rec_n] = make_int4(rec_used, seg_total, blk, 0);
                        rec_used += seg_total;
                        ++rec_n;
                    } else {
                        rec_ok = false;
                    }
                }
                // ---- phase B: one thread per entry, two independent entries in flight per trip ----
                for (int e = tid; e < seg_total; e += 2 * NT) {
                    const int e2 = e + NT;
                    const uint32_t va = ent[e], vb = (e2 < seg_total) ? ent[e2] : PAIR_INVALID;
                    uint32_t ra = 0, ia = 0, rb = 0, ib = 0;
                    const uint32_t da = entry_distance(va, ra, ia), db = entry_distance(vb, rb, ib);
                    const uint32_t oa = entry_apply(da, ra, ia), ob = entry_apply(db, rb, ib);
                    if (thresholds) {
                        ent[e] = oa;
                        if (e2 < seg_total) ent[e2] = ob;
                    }
                    if (rec_dst) {
                        rec_dst[e] = (da == 0xFFFFu) ? PAIR_INVALID : ((ia << 17) | (ra << 9) | da);
                        if (e2 < seg_total) rec_dst[e2] = (db == 0xFFFFu) ? PAIR_INVALID : ((ib << 17) | (rb << 9) | db);
                    }
                }
                __syncthreads();
                if (PASS == 0) PLM_TL(4);
                if (!thresholds) continue;
                list_rounds(seg_total, base);
                if (seg + 1 < n_seg || base + NT < row_end) next_thresholds(); // for the rows that follow
            }
        } else {
            // a single row has more slots than half the list: every thread walks its own row
            rec_ok = false;
            if (has_row) {
                if (PASS == 0) {
                    row_walk(job, gp, r, [&](int i2, int d) {
                        if (static_cast<uint32_t>(d) < K[i2]) atomicMin(&K[i2], static_cast<uint32_t>(d));
                    });
                } else if (!gp.best_lr) {
                    row_walk(job, gp, r, take);
                } else {
                    row_walk(job, gp, r, [&](int i2, int d) {
                        if (d < T[i2]) {
                            und = true;
                            atomicMin(&K[i2], (static_cast<uint32_t>(d) << 16) | utid);
                        }
                    });
                }
            }
            __syncthreads();
            if (thresholds) {
                for (int round = 0;; ++round) {
                    if (und) {
                        bool again = false;
                        row_walk(job, gp, r, [&](int i2, int d) {
                            if (d < T[i2] && utid < B[i2]) {
                                const uint32_t key = (static_cast<uint32_t>(d) << 16) | utid, k = K[i2];
                                if (key == k) take(i2, d);
                                else if (utid < (k & 0xFFFFu)) again = true;
                            }
                        });
                        und = again;
                    }
                    __syncthreads();
                    fold(round);
                    if (!__syncthreads_or(und ? 1 : 0)) break;
                    if (und)
                        row_walk(job, gp, r, [&](int i2, int d) {
                            if (d < T[i2] && utid < B[i2]) atomicMin(&K[i2], (static_cast<uint32_t>(d) << 16) | utid);
                        });
                    __syncthreads();
                }
                if (base + NT < row_end) next_thresholds();
            }
        }
        if (PASS == 0) continue; // every path above ends with a block barrier; K accumulates over the blocks of the CTA
        if (listed) {
            b0 = rb0[tid];
            b1 = rb1[tid];
        }
        __syncthreads(); // rb / ent / T are rewritten by the next block
        if (has_row && b0 != KEY32_ABSENT) {
            // matching.cpp:160 / :241 -- int -> double, one double multiply, strict compare
            const int best_d = static_cast<int>(b0 >> GRID_KEY_BITS);
            const int best_d2 = (b1 == KEY32_ABSENT) ? 0x7FFFFFFF : static_cast<int>(b1 >> GRID_KEY_BITS);
            if (static_cast<double>(best_d) < __dmul_rn(static_cast<double>(best_d2), gp.ratio)) {
                job.m12[i1] = static_cast<int32_t>(b0 & ((1u << GRID_KEY_BITS) - 1));
                ++accepted;
            }
        }
    }
    if (PASS == 0) {
        __syncthreads();
        for (int i = tid; i < n2; i += NT) cmin_out[i] = static_cast<uint16_t>(min(K[i], 0xFFFFu));
        if (recording && !rec_smem && tid == 0) gp.seg_cnt[cta] = rec_ok ? rec_n : -1;
        return;
    }
    if (accepted) atomicAdd(job.count, accepted);
    }; // run_pass

    if (MODE == 0) {
        run_pass(std::integral_constant<int, 0>{}, cta_min, nullptr);
        PLM_TL(5);
    } else if (MODE == 1) {
        run_pass(std::integral_constant<int, 1>{}, nullptr, cta_min);
        PLM_TL(10);
    } else {
        cg::cluster_group cluster = cg::this_cluster();
        const int rank = static_cast<int>(cluster.block_rank());
        if (gp.best_lr) {
            if (rank == 0)
                for (int i = tid; i < n2; i += NT) gp.m21key[i] = KEY64_ABSENT;
            run_pass(std::integral_constant<int, 0>{}, cmin16, nullptr);
            PLM_TL(5);
            cluster.sync();
            PLM_TL(6);
            // exclusive prefix-min over the lower-ranked CTAs: all remote loads of a column in flight together
            const uint16_t *rem[GRID_CLUSTER_MAX_CTAS];
#pragma unroll
            for (int r = 0; r < GRID_CLUSTER_MAX_CTAS; ++r) rem[r] = cluster.map_shared_rank(cmin16, min(r, max(rank - 1, 0)));
            for (int i = tid; i < n2; i += NT) {
                uint16_t v[GRID_CLUSTER_MAX_CTAS];
#pragma unroll
                for (int r = 0; r < GRID_CLUSTER_MAX_CTAS - 1; ++r) v[r] = (r < rank) ? rem[r][i] : D_INF;
                uint16_t t = D_INF;
#pragma unroll
                for (int r = 0; r < GRID_CLUSTER_MAX_CTAS - 1; ++r) t = min(t, v[r]);
                Tnew[i] = t;
            }
            __syncthreads();
            PLM_TL(7);
        }
        run_pass(std::integral_constant<int, 1>{}, nullptr, Tnew);
        PLM_TL(10);
        if (gp.best_lr) {
            cluster.sync(); // every CTA's m21 keys are in; nobody's cmin16 is read any more
            PLM_TL(11);
            // mutual check (matching.cpp:166-174) over this CTA's rows, stale entries included
            int culled = 0;
            for (long long i1 = cta_row0 + tid; i1 < row_end; i1 += NT) {
                const int32_t i2 = job.m12[i1];
                if (i2 >= 0) {
                    const unsigned long long k = (i2 < n2) ? gp.m21key[i2] : KEY64_ABSENT;
                    const long long back = (k == KEY64_ABSENT) ? -1 : static_cast<long long>(k & 0xFFFFFFFFull);
                    if (back != job.i1_base + i1) {
                        job.m12[i1] = -1;
                        ++culled;
                    }
                }
            }
            if (culled) atomicSub(job.count, culled);
            PLM_TL(12);
        }
    }
}

template <int PASS, int STAGED>
__global__ void __launch_bounds__(GRID_ROW_THREADS, 3)
grid_rows_kernel(GridJob job, GridParams gp) { grid_rows_device<STAGED, PASS>(job, gp, static_cast<int>(blockIdx.x)); }

// One launch, one cluster of gridDim.x <= 8 CTAs (MODE 2 above): the single-call path for frame-sized jobs.
template <int STAGED>
__global__ void __launch_bounds__(GRID_ROW_THREADS, 3)
grid_rows_cluster_kernel(GridJob job, GridParams gp) { grid_rows_device<STAGED, 2>(job, gp, static_cast<int>(blockIdx.x)); }

// cta_min[c][i2] <- min(seed[i2], min over c' < c of cta_min[c'][i2]) when CTA c's own minimum beats it, else 0;
// col_min[i2] = overall minimum.
// seed (may be null) carries the minima of lower-ranked database shards (multi-GPU).
// One WARP per column: lanes take 32 consecutive CTAs at a time and scan them with shuffles, so the
// dependent chain is n_cta / 32 steps long instead of n_cta (the map-sized launch has ~400 CTAs).
__global__ void grid_scan_kernel(uint16_t *__restrict__ cta_min, int n_cta, int n2,
                                 const uint16_t *__restrict__ seed, uint16_t *__restrict__ col_min) {
    const int lane = threadIdx.x & 31;
    const int i2 = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i2 >= n2) return;
    uint32_t run = seed ? seed[i2] : D_INF;
    // 8 x 32 CTAs per trip: the loads of a trip are issued together (they alias the stores of the scan, so the
    // compiler cannot hoist them itself) and the dependent chain is n_cta / 256 L2 round trips long
    constexpr int TRIP = 8;
    for (int c0 = 0; c0 < n_cta; c0 += 32 * TRIP) {
        uint32_t t[TRIP];
#pragma unroll
        for (int k = 0; k < TRIP; ++k) {
            const int c = c0 + 32 * k + lane;
            t[k] = (c < n_cta) ? cta_min[static_cast<size_t>(c) * n2 + i2] : D_INF;
        }
#pragma unroll
        for (int k = 0; k < TRIP; ++k) {
            const int c = c0 + 32 * k + lane;
            uint32_t incl = t[k];
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, s);
                if (lane >= s) incl = min(incl, v);
            }
            uint32_t excl = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
            if (lane == 0) excl = D_INF;
            // a CTA whose own minimum does not beat what it inherits has no live pair in this column: threshold 0 tells
            // pass 1 to skip the column's pairs without computing their distances
            const uint32_t inherit = min(run, excl);
            if (c < n_cta) cta_min[static_cast<size_t>(c) * n2 + i2] = static_cast<uint16_t>(t[k] < inherit ? inherit : 0u);
            run = min(run, __shfl_sync(0xFFFFFFFFu, incl, 31));
        }
    }
    if (col_min && lane == 0) col_min[i2] = static_cast<uint16_t>(run);
}

// cta_min[c][i2] <- min(cta_min[c][i2], seed[i2]): the thresholds of a shard that were scanned without the minima of
// the lower-ranked shards receive them afterwards (sharded matchGrid with ONE minima pass).
__global__ void grid_seed_kernel(uint16_t *__restrict__ cta_min, int n_cta, int n2, const uint16_t *__restrict__ seed) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(n_cta) * n2) return;
    const uint16_t s = seed[i % n2];
    if (s < cta_min[i]) cta_min[i] = s;
}

// m21[i2] = row of the best live pair (or -1), from the 64-bit keys of the chunked launch.
__global__ void m21_from_keys_kernel(const unsigned long long *__restrict__ m21key, int n2,
                                     int32_t *__restrict__ m21) {
    const int i2 = blockIdx.x * blockDim.x + threadIdx.x;
    if (i2 >= n2) return;
    const unsigned long long k = m21key[i2];
    m21[i2] = (k == KEY64_ABSENT) ? -1 : static_cast<int32_t>(k & 0xFFFFFFFFull);
}

} // namespace plm
