// Brute-force Hamming top-2 (replaces cv::BFMatcher::knnMatch(k=2) behind StVO::matchNNR,
// call site stvo-pl/src/matching.cpp:47-48) and the small epilogues of matchNNR / match.
//
// Layout: one thread owns one query descriptor in 8 registers; the CTA streams its slice of the
// train set ("database") through shared memory in 8 KB stages (cp.async, double buffered); all
// lanes read the same train row, so each row costs two broadcast LDS.128 per warp.  Per pair:
// 8 LOP3(xor) + 8 POPC + IADD3 tree (or the 5-POPC carry-save form), then a 32-bit packed key
// (dist << 22 | slice-local row) goes through a 3-instruction min/max top-2 update.  Slices are
// merged lexicographically afterwards, so ties always resolve to the lowest train index.
#pragma once
#include "plm_common.cuh"

namespace plm {

constexpr int KNN_STAGE_ROWS = 256;               // train rows per shared-memory stage (8 KB)
constexpr int KNN_IDX_BITS = 22;                  // slice-local row index bits in the 32-bit key
constexpr int KNN_MAX_SLICE_ROWS = 1 << KNN_IDX_BITS;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ unsigned long long expand_key(uint32_t k, uint32_t row_base) {
    if (k == KEY32_ABSENT) return KEY64_ABSENT;
    return make_key64(k >> KNN_IDX_BITS, row_base + (k & (KNN_MAX_SLICE_ROWS - 1)));
}

// One brute-force direction: queries q[0..n1) against train rows db[0..n2).  The train set is cut
// into n_units units of unit_rows <= KNN_STAGE_ROWS rows (one shared-memory stage each); CTA (qblock,
// worker) scans units worker, worker + n_workers, ... in increasing row order through a continuous
// double-buffered cp.async pipeline and keeps a running top-2 of packed 64-bit keys in registers.  A
// launch is therefore a single wave of co-resident CTAs whatever the database size, work is balanced
// to one 256-row unit, and only n_workers partial results per query exist.
// part[worker][query] = (best, second).
struct KnnTask {
    const uint4 *q;
    const uint4 *db;
    ulonglong2 *part;   // [n_workers][n1]
    ulonglong2 *top2;   // [n1] merged result (may be null when only the acceptance is wanted)
    int32_t *m;         // in/out match vector written by the NNR acceptance (may be null)
    int32_t *count;     // accepted-row counter (may be null)
    unsigned long long idx_base;
    long long n2;
    int n1, unit_rows, n_workers, n_units;
    // the first extra_qb query blocks (of `threads` queries each) have n_workers + 1 workers, so that qblocks x workers
    // can equal the number of resident CTA slots exactly; 0 = every block has n_workers
    int extra_qb, threads;
    // optional per-query bound shared by the workers of a launch (long scans only): the smallest second-best DISTANCE
    // any worker has found so far.  A row farther than that cannot be among the two nearest of the whole train set,
    // whichever worker scans it, so workers skip it -- the merged result is unchanged (rows AT the bound are kept:
    // their index may be lower).
    uint32_t *gthr;
};
__device__ __forceinline__ int knn_workers_of(const KnnTask &t, int qblock) { return t.n_workers + (qblock < t.extra_qb ? 1 : 0); }
struct KnnTaskPair {
    KnnTask t[2];
    int cta_split; // CTAs [0, cta_split) belong to t[0], the rest to t[1]
    int pad_;
};

// VARIANT 0: plain 8-POPC distance, per-pair top-2 update
// VARIANT 1: 5-POPC carry-save distance, per-pair top-2 update (frame-sized train sets)
// VARIANT 2: 4-POPC carry-save distance, rows handled 8 at a time: the eight distances only go
//            through the packed-key top-2 update when their minimum beats the thread's current
//            second-best distance.  Train rows are scanned in increasing index order, so a row that
//            merely ties the second best can never displace it (its key is larger) and the strict
//            test is exact.  For long scans the update path is almost never taken.
// VARIANT 3: as 2 with the 13-LOP3 distance (plm_common.cuh): the query is transformed once, every stage of
//            train rows is transformed in place in shared memory before it is scanned.
// VARIANT 4: as 3, but the first KNN_B_ROWS rows of every group of 8 carry the parity transform and go through the
//            16-LOP3 / 3-POPC lower bound: the blocked test min(...) < thr stays exact (a lower bound can only make the
//            update path run more often, and that path recomputes those rows exactly), while the XU pipe (POPC), the
//            binding one in variant 3, does 29 instead of 32 instructions per 8 pairs and the ALU pipe 118 instead of 109.
// VARIANT 5: the 13-LOP3 distance of variant 3 with the per-pair top-2 update of variant 1 -- for short train sets
//            (frame-sized calls, per-keyframe-pair batches), where the blocked threshold never gets tight.
constexpr int KNN_B_ROWS = 3;

template <int VARIANT>
__device__ __forceinline__ int knn_dist(const Desc &a, const uint4 &blo, const uint4 &bhi) {
    if (VARIANT >= 3) return hamming256_t13(a, blo, bhi);
    if (VARIANT == 2) return hamming256_csa4(a, blo, bhi);
    if (VARIANT == 1) return hamming256_csa(a, blo, bhi);
    return hamming256(a, blo, bhi);
}

template <int THREADS, int VARIANT>
__device__ __forceinline__ void knn2_slice_body(const KnnTask &t, int qblock, int worker, int n_workers, uint4 (*stage)[KNN_STAGE_ROWS * 2]) {
    const int tid = threadIdx.x;
    const int n1 = t.n1;
    const int qi = qblock * THREADS + tid;
    Desc a;
    if (qi < n1) {
        a = load_desc(t.q, qi);
    } else {
        a.lo = make_uint4(0, 0, 0, 0);
        a.hi = a.lo;
    }
    if (VARIANT >= 3) desc_transform13(a.lo, a.hi);
    Desc ab = a; // variant 4: the query in the parity-transformed domain
    if (VARIANT == 4) ab.hi.w ^= ab.hi.z;
    const uint4 *db = t.db;
    const int unit_rows = t.unit_rows, n_units = t.n_units, step = n_workers;
    const uint32_t idx_base = static_cast<uint32_t>(t.idx_base); // idx_base + n2 <= 2^32 (checked by the host)
    unsigned long long B0 = KEY64_ABSENT, B1 = KEY64_ABSENT;     // running top-2 over all units of this worker
    int thr = 1023;                                              // variant >= 2: distance a row must beat (> 256: none yet)
    int gcap = 1023;                                             // ... and the cap taken from the launch-wide bound
    uint32_t *gthr = (VARIANT >= 2 && t.gthr && qi < n1) ? t.gthr + qi : nullptr;
    auto lower_thr = [&]() { // own second best changed: new strict threshold, published to the other workers
        const int own = (B1 == KEY64_ABSENT) ? 1023 : static_cast<int>(B1 >> 32);
        if (gthr && own < 1023) atomicMin(gthr, static_cast<uint32_t>(own));
        thr = min(own, gcap);
    };

    auto rows_of = [&](int u) { return static_cast<int>(min(static_cast<long long>(unit_rows), t.n2 - static_cast<long long>(u) * unit_rows)); };
    auto issue = [&](int u, int buf) {
        const int rows = rows_of(u);
        const uint4 *src = db + 2 * (static_cast<long long>(u) * unit_rows);
        uint4 *dst = stage[buf];
        for (int i = tid; i < rows * 2; i += THREADS) cp_async16(dst + i, src + i);
        cp_async_commit();
    };

    int buf = 0;
    if (worker < n_units) issue(worker, 0);
    for (int u = worker; u < n_units; u += step, buf ^= 1) {
        if (u + step < n_units) {
            issue(u + step, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const int rows = rows_of(u);
        if (VARIANT >= 3) {
            uint4 *w = stage[buf];
            for (int r = tid; r < rows; r += THREADS) {
                uint4 lo = w[2 * r], hi = w[2 * r + 1];
                desc_transform13(lo, hi);
                if (VARIANT == 4 && (r & 7) < KNN_B_ROWS) hi.w ^= hi.z;
                w[2 * r] = lo;
                w[2 * r + 1] = hi;
            }
            __syncthreads();
        }
        if (VARIANT >= 2 && gthr) { // once per stage: the bound the other workers have reached (rows AT the bound stay in)
            const uint32_t g = *reinterpret_cast<volatile uint32_t *>(gthr);
            gcap = static_cast<int>(min(g, 1022u)) + 1;
            thr = min(thr, gcap);
        }
        const uint4 *sb = stage[buf];
        const uint32_t gbase = idx_base + static_cast<uint32_t>(u) * static_cast<uint32_t>(unit_rows);
        if (VARIANT >= 2 && VARIANT <= 4) {
            int j = 0;
            for (; j + 8 <= rows; j += 8) {
                int d[8];
#pragma unroll
                for (int v = 0; v < 8; ++v)
                    d[v] = (VARIANT == 4 && v < KNN_B_ROWS) ? hamming256_t16<false>(ab, sb[2 * (j + v)], sb[2 * (j + v) + 1])
                                                            : knn_dist<VARIANT>(a, sb[2 * (j + v)], sb[2 * (j + v) + 1]);
                const int m = min(min(min(d[0], d[1]), min(d[2], d[3])), min(min(d[4], d[5]), min(d[6], d[7])));
                if (m < thr) {
                    if (VARIANT == 4) {
#pragma unroll
                        for (int v = 0; v < KNN_B_ROWS; ++v) d[v] = hamming256_t16<true>(ab, sb[2 * (j + v)], sb[2 * (j + v) + 1]);
                    }
#pragma unroll
                    for (int v = 0; v < 8; ++v) top2_insert(B0, B1, make_key64(d[v], gbase + j + v));
                    lower_thr();
                }
            }
            for (; j < rows; ++j) {
                const int d = (VARIANT == 4 && (j & 7) < KNN_B_ROWS) ? hamming256_t16<true>(ab, sb[2 * j], sb[2 * j + 1])
                                                                     : knn_dist<VARIANT>(a, sb[2 * j], sb[2 * j + 1]);
                if (d < thr) {
                    top2_insert(B0, B1, make_key64(d, gbase + j));
                    lower_thr();
                }
            }
        } else {
            uint32_t b0 = KEY32_ABSENT, b1 = KEY32_ABSENT; // unit-local keys (d << 22 | row in unit)
#pragma unroll 8
            for (int j = 0; j < rows; ++j) {
                const uint4 blo = sb[2 * j], bhi = sb[2 * j + 1];
                const int d = knn_dist<VARIANT>(a, blo, bhi);
                top2_insert(b0, b1, (static_cast<uint32_t>(d) << KNN_IDX_BITS) + static_cast<uint32_t>(j));
            }
            top2_insert(B0, B1, expand_key(b0, gbase));
            top2_insert(B0, B1, expand_key(b1, gbase));
        }
        __syncthreads();
    }
    if (qi < n1) t.part[static_cast<size_t>(worker) * n1 + qi] = make_ulonglong2(B0, B1);
}


// VARIANT 6: variant 3 with TWO queries per thread (a CTA of 128 threads owns 256 queries: thread t holds queries
// t and t + 128 of its block).  Every train row fetched from shared memory (2 LDS.128) now serves two pairs, and the
// loop overhead is shared: the MIO queue, through which both the LDS and the POPC instructions issue, sees 16 LDS per
// 16 pairs instead of per 8.  Same blocked threshold logic per query.
template <int THREADS>
__device__ __forceinline__ void knn2_slice_body2(const KnnTask &t, int qblock, int worker, int n_workers, uint4 (*stage)[KNN_STAGE_ROWS * 2]) {
    const int tid = threadIdx.x;
    const int n1 = t.n1;
    int qi[2];
    Desc a[2];
    unsigned long long B0[2], B1[2];
    int thr[2], gcap[2];
    uint32_t *gthr[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        qi[k] = qblock * (2 * THREADS) + k * THREADS + tid;
        if (qi[k] < n1) {
            a[k] = load_desc(t.q, qi[k]);
        } else {
            a[k].lo = make_uint4(0, 0, 0, 0);
            a[k].hi = a[k].lo;
        }
        desc_transform13(a[k].lo, a[k].hi);
        B0[k] = KEY64_ABSENT;
        B1[k] = KEY64_ABSENT;
        thr[k] = 1023;
        gcap[k] = 1023;
        gthr[k] = (t.gthr && qi[k] < n1) ? t.gthr + qi[k] : nullptr;
    }
    const uint4 *db = t.db;
    const int unit_rows = t.unit_rows, n_units = t.n_units, step = n_workers;
    const uint32_t idx_base = static_cast<uint32_t>(t.idx_base);
    auto rows_of = [&](int u) { return static_cast<int>(min(static_cast<long long>(unit_rows), t.n2 - static_cast<long long>(u) * unit_rows)); };
    auto issue = [&](int u, int buf) {
        const int rows = rows_of(u);
        const uint4 *src = db + 2 * (static_cast<long long>(u) * unit_rows);
        uint4 *dst = stage[buf];
        for (int i = tid; i < rows * 2; i += THREADS) cp_async16(dst + i, src + i);
        cp_async_commit();
    };
    int buf = 0;
    if (worker < n_units) issue(worker, 0);
    for (int u = worker; u < n_units; u += step, buf ^= 1) {
        if (u + step < n_units) {
            issue(u + step, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const int rows = rows_of(u);
        {
            uint4 *w = stage[buf];
            for (int r = tid; r < rows; r += THREADS) {
                uint4 lo = w[2 * r], hi = w[2 * r + 1];
                desc_transform13(lo, hi);
                w[2 * r] = lo;
                w[2 * r + 1] = hi;
            }
            __syncthreads();
        }
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (gthr[k]) {
                const uint32_t g = *reinterpret_cast<volatile uint32_t *>(gthr[k]);
                gcap[k] = static_cast<int>(min(g, 1022u)) + 1;
                thr[k] = min(thr[k], gcap[k]);
            }
        const uint4 *sb = stage[buf];
        const uint32_t gbase = idx_base + static_cast<uint32_t>(u) * static_cast<uint32_t>(unit_rows);
        auto update = [&](int k, int d, uint32_t row) {
            top2_insert(B0[k], B1[k], make_key64(d, row));
        };
        auto lower = [&](int k) {
            const int own = (B1[k] == KEY64_ABSENT) ? 1023 : static_cast<int>(B1[k] >> 32);
            if (gthr[k] && own < 1023) atomicMin(gthr[k], static_cast<uint32_t>(own));
            thr[k] = min(own, gcap[k]);
        };
        int j = 0;
        for (; j + 8 <= rows; j += 8) {
            int d0[8], d1[8];
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                const uint4 blo = sb[2 * (j + v)], bhi = sb[2 * (j + v) + 1];
                d0[v] = hamming256_t13(a[0], blo, bhi);
                d1[v] = hamming256_t13(a[1], blo, bhi);
            }
            const int m0 = min(min(min(d0[0], d0[1]), min(d0[2], d0[3])), min(min(d0[4], d0[5]), min(d0[6], d0[7])));
            const int m1 = min(min(min(d1[0], d1[1]), min(d1[2], d1[3])), min(min(d1[4], d1[5]), min(d1[6], d1[7])));
            if (m0 < thr[0]) {
#pragma unroll
                for (int v = 0; v < 8; ++v) update(0, d0[v], gbase + j + v);
                lower(0);
            }
            if (m1 < thr[1]) {
#pragma unroll
                for (int v = 0; v < 8; ++v) update(1, d1[v], gbase + j + v);
                lower(1);
            }
        }
        for (; j < rows; ++j) {
            const uint4 blo = sb[2 * j], bhi = sb[2 * j + 1];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int d = hamming256_t13(a[k], blo, bhi);
                if (d < thr[k]) {
                    update(k, d, gbase + j);
                    lower(k);
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < 2; ++k)
        if (qi[k] < n1) t.part[static_cast<size_t>(worker) * n1 + qi[k]] = make_ulonglong2(B0[k], B1[k]);
}

// 1-D grid over (task, query block, worker): both directions of StVO::match in one launch.
template <int THREADS, int VARIANT>
__global__ void __launch_bounds__(THREADS, (VARIANT == 6) ? 6 : ((THREADS == 128) ? 9 : 12)) knn2_slice_kernel(const KnnTaskPair tasks) {
    __shared__ __align__(16) uint4 stage[2][KNN_STAGE_ROWS * 2];
    const int which = static_cast<int>(blockIdx.x) >= tasks.cta_split ? 1 : 0;
    const KnnTask &t = tasks.t[which];
    int local = static_cast<int>(blockIdx.x) - (which ? tasks.cta_split : 0);
    int nw = t.n_workers, qb0 = 0;
    const int wide = t.extra_qb * (t.n_workers + 1); // CTAs of the query blocks that have one more worker
    if (local < wide) {
        nw = t.n_workers + 1;
    } else {
        local -= wide;
        qb0 = t.extra_qb;
    }
    if (VARIANT == 6) knn2_slice_body2<THREADS>(t, qb0 + local / nw, local % nw, nw, stage);
    else knn2_slice_body<THREADS, VARIANT == 6 ? 3 : VARIANT>(t, qb0 + local / nw, local % nw, nw, stage);
}

// Batched form: cta_map[cta] = (task, qblock, worker); tasks live in device memory.
template <int THREADS, int VARIANT>
__global__ void __launch_bounds__(THREADS) knn2_slice_list_kernel(const KnnTask *__restrict__ tasks, const int4 *__restrict__ cta_map) {
    __shared__ __align__(16) uint4 stage[2][KNN_STAGE_ROWS * 2];
    const int4 m = __ldg(cta_map + blockIdx.x);
    const KnnTask t = tasks[m.x];
    knn2_slice_body<THREADS, VARIANT>(t, m.y, m.z, t.n_workers, stage);
}

// Slice merge + (optionally) the matchNNR acceptance (stvo-pl/src/matching.cpp:53-58).
// top2[q] = two smallest of part[p][q].{x,y}, p < n_workers.  Keys with different train indices are
// distinct, so a plain lexicographic min-2 reproduces lowest-index tie-breaking across slices,
// database shards and GPUs alike.  Acceptance is a float compare of a float product (no FMA); only
// accepted rows are written because m is the reference's in/out matches_12.
__device__ __forceinline__ void knn2_merge_body(const KnnTask &t, int q, float nnr, bool do_accept) {
    bool acc = false;
    if (q < t.n1) {
        unsigned long long b0 = KEY64_ABSENT, b1 = KEY64_ABSENT;
        const int nw = t.extra_qb ? knn_workers_of(t, q / t.threads) : t.n_workers;
        for (int p = 0; p < nw; ++p) {
            const ulonglong2 v = t.part[static_cast<size_t>(p) * t.n1 + q];
            top2_insert(b0, b1, v.x);
            top2_insert(b0, b1, v.y);
        }
        if (t.top2) t.top2[q] = make_ulonglong2(b0, b1);
        if (do_accept && b1 != KEY64_ABSENT) {
            const float d0 = static_cast<float>(static_cast<int>(b0 >> 32));
            const float d1 = static_cast<float>(static_cast<int>(b1 >> 32));
            if (d0 < __fmul_rn(d1, nnr)) {
                t.m[q] = static_cast<int32_t>(b0 & 0xFFFFFFFFull);
                acc = true;
            }
        }
    }
    if (do_accept) {
        const unsigned msk = __ballot_sync(0xFFFFFFFFu, acc);
        if ((threadIdx.x & 31) == 0 && msk && t.count) atomicAdd(t.count, __popc(msk));
    }
}

__global__ void knn2_merge_kernel(const KnnTaskPair tasks, int task0, float nnr, int do_accept) {
    const KnnTask &t = tasks.t[task0 + blockIdx.y];
    knn2_merge_body(t, blockIdx.x * blockDim.x + threadIdx.x, nnr, do_accept != 0);
}

// Wide form for many partials per query (a short query side against a long train side, e.g. the 21
// direction of the map fallback: 600 queries x ~180 workers): one WARP per query, lanes stride over the
// workers and finish with a shuffle butterfly.  Same lexicographic min-2, same acceptance.
__global__ void knn2_merge_wide_kernel(const KnnTaskPair tasks, int task0, float nnr, int do_accept) {
    const KnnTask &t = tasks.t[task0 + blockIdx.y];
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= t.n1) return;
    unsigned long long b0 = KEY64_ABSENT, b1 = KEY64_ABSENT;
    const int nw = t.extra_qb ? knn_workers_of(t, q / t.threads) : t.n_workers;
    for (int p = lane; p < nw; p += 32) {
        const ulonglong2 v = t.part[static_cast<size_t>(p) * t.n1 + q];
        top2_insert(b0, b1, v.x);
        top2_insert(b0, b1, v.y);
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const unsigned long long o0 = __shfl_xor_sync(0xFFFFFFFFu, b0, s), o1 = __shfl_xor_sync(0xFFFFFFFFu, b1, s);
        top2_insert(b0, b1, o0);
        top2_insert(b0, b1, o1);
    }
    if (lane != 0) return;
    if (t.top2) t.top2[q] = make_ulonglong2(b0, b1);
    if (do_accept && b1 != KEY64_ABSENT) {
        const float d0 = static_cast<float>(static_cast<int>(b0 >> 32));
        const float d1 = static_cast<float>(static_cast<int>(b1 >> 32));
        if (d0 < __fmul_rn(d1, nnr)) {
            t.m[q] = static_cast<int32_t>(b0 & 0xFFFFFFFFull);
            if (t.count) atomicAdd(t.count, 1);
        }
    }
}

// Batched form: merge_map[cta] = (task, first query of this CTA).
__global__ void knn2_merge_list_kernel(const KnnTask *__restrict__ tasks, const int2 *__restrict__ merge_map, float nnr, int do_accept) {
    const int2 m = __ldg(merge_map + blockIdx.x);
    const KnnTask t = tasks[m.x];
    knn2_merge_body(t, m.y + threadIdx.x, nnr, do_accept != 0);
}

// Stand-alone merge of n_parts x n1 x 2 keys (database shards gathered from other GPUs).
__global__ void top2_merge_kernel(const ulonglong2 *__restrict__ parts, int n_parts, int n1,
                                  ulonglong2 *__restrict__ out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n1) return;
    unsigned long long b0 = KEY64_ABSENT, b1 = KEY64_ABSENT;
    for (int p = 0; p < n_parts; ++p) {
        const ulonglong2 v = parts[static_cast<size_t>(p) * n1 + q];
        top2_insert(b0, b1, v.x);
        top2_insert(b0, b1, v.y);
    }
    out[q] = make_ulonglong2(b0, b1);
}

// matchNNR acceptance from already merged keys (multi-GPU path).
__global__ void nnr_accept_kernel(const ulonglong2 *__restrict__ top2, int n1, float nnr,
                                  int32_t *__restrict__ m12, int32_t *__restrict__ count) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    bool acc = false;
    if (q < n1) {
        const ulonglong2 v = top2[q];
        if (v.y != KEY64_ABSENT) {
            const float d0 = static_cast<float>(static_cast<int>(v.x >> 32));
            const float d1 = static_cast<float>(static_cast<int>(v.y >> 32));
            if (d0 < __fmul_rn(d1, nnr)) {
                m12[q] = static_cast<int32_t>(v.x & 0xFFFFFFFFull);
                acc = true;
            }
        }
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, acc);
    if ((threadIdx.x & 31) == 0 && m && count) atomicAdd(count, __popc(m));
}

// Mutual check (matching.cpp:80-86 / :166-174): every m12 entry >= 0 -- stale ones included --
// whose reverse match is not i1 is culled and decrements the count.
__global__ void cross_check_kernel(int32_t *__restrict__ m12, int n1, long long i1_base,
                                   const int32_t *__restrict__ m21, long long n2,
                                   int32_t *__restrict__ count) {
    const int i1 = blockIdx.x * blockDim.x + threadIdx.x;
    bool cull = false;
    if (i1 < n1) {
        const int32_t i2 = m12[i1];
        if (i2 >= 0 && (i2 >= n2 || static_cast<long long>(m21[i2]) != i1_base + i1)) {
            m12[i1] = -1;
            cull = true;
        }
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, cull);
    if ((threadIdx.x & 31) == 0 && m) atomicSub(count, __popc(m));
}

// The same mutual check straight from the per-column keys of matchGrid (distance << 32 | global row; all ones = no
// live pair): saves the separate key -> row conversion launch.
__global__ void cross_check_keys_kernel(int32_t *__restrict__ m12, int n1, long long i1_base,
                                        const unsigned long long *__restrict__ m21key, long long n2,
                                        int32_t *__restrict__ count) {
    const int i1 = blockIdx.x * blockDim.x + threadIdx.x;
    bool cull = false;
    if (i1 < n1) {
        const int32_t i2 = m12[i1];
        if (i2 >= 0) {
            const unsigned long long k = (i2 < n2) ? m21key[i2] : KEY64_ABSENT;
            const long long back = (k == KEY64_ABSENT) ? -1 : static_cast<long long>(static_cast<int32_t>(k & 0xFFFFFFFFull));
            if (back != i1_base + i1) {
                m12[i1] = -1;
                cull = true;
            }
        }
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, cull);
    if ((threadIdx.x & 31) == 0 && m) atomicSub(count, __popc(m));
}

// StVO::distance for n independent pairs.
__global__ void hamming_pairs_kernel(const uint4 *__restrict__ a, const uint4 *__restrict__ b, int n,
                                     int32_t *__restrict__ dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dist[i] = hamming256(load_desc(a, i), load_desc(b, i));
}

} // namespace plm

namespace plm {

// Batched mutual check: xmap[cta] = (job, first row of this CTA).
struct XJob {
    int32_t *m12;
    const int32_t *m21;
    int32_t *count;
    int32_t n1, n2;
};

__global__ void cross_check_list_kernel(const XJob *__restrict__ jobs, const int2 *__restrict__ xmap) {
    const int2 m = __ldg(xmap + blockIdx.x);
    const XJob j = jobs[m.x];
    const int i1 = m.y + threadIdx.x;
    bool cull = false;
    if (i1 < j.n1) {
        const int32_t i2 = j.m12[i1];
        if (i2 >= 0 && (i2 >= j.n2 || j.m21[i2] != i1)) {
            j.m12[i1] = -1;
            cull = true;
        }
    }
    const unsigned msk = __ballot_sync(0xFFFFFFFFu, cull);
    if ((threadIdx.x & 31) == 0 && msk) atomicSub(j.count, __popc(msk));
}

} // namespace plm
