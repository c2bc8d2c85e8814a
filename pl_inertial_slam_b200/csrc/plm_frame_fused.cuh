// One frame, one launch: the matcher calls of a live frame -- StereoFrame::matchStereoPoints / matchStereoLines
// (matchGrid, stvo-pl/src/stereoFrame.cpp:157, :356) and StereoFrameHandler::matchF2FPoints / matchF2FLines
// (StVO::match, stvo-pl/src/stereoFrameHandler.cpp:168, :191) -- executed by ONE kernel.  The reference runs the
// point and line halves on two std::async threads (stereoFrame.cpp:75-76, stereoFrameHandler.cpp:142-143); here every
// call of the frame session owns one or more thread-block CLUSTERS of 8 CTAs, all of them in one grid:
//
//   matchGrid jobs   grid_rows_device<STAGED, 2> (plm_grid.cuh): pass 0, cluster barrier, thresholds of the
//                    lower-ranked CTAs through distributed shared memory, pass 1, cluster barrier, mutual check.
//   match jobs       match_job_device below, on one or more clusters whose CTAs work independently: a CTA owns a block
//                    of query rows of ONE direction (train side staged in shared memory), a row is split over S
//                    threads (column slices) whose packed top-2 keys meet in shared memory; direction 21 writes
//                    matches_21 to global scratch; the last CTA of the job to finish runs the mutual check.
//
// The job table travels in the same host -> device copy as the descriptors and the kernel stores the final match
// vectors and counts straight into the pinned host block, so a frame is: one copy in, one launch, one synchronisation
// (plm_frame_end in plmatch.cu).
#pragma once
#include "plm_grid.cuh"

namespace plm {

constexpr int FRAME_CLUSTER = 8;          // CTAs per job
constexpr int FRAME_MATCH_MAX_ROWS = 2048; // per side: a CTA owns <= 256 rows, keys carry a 16-bit index

struct MatchJob {
    const uint4 *d1, *d2;
    int32_t *m12;   // in/out, n1 (accepted rows are written, the mutual check culls -- stale entries included)
    int32_t *count; // accepted - culled
    int32_t *done;  // CTAs of this job that have finished their rows (zero on entry)
    int32_t *m21;   // scratch, n2 (best_lr only)
    int32_t n1, n2, best_lr;
    int32_t n_cta, c12; // CTAs of the job; the first c12 take direction 12, the others direction 21
    float nnr;
};

struct alignas(16) FrameJobRec {
    int32_t kind;      // 0 = match / matchNNR, 1 = matchGrid
    int32_t cta_begin; // first CTA of the job in the grid (a multiple of FRAME_CLUSTER)
    int32_t *h_io;     // the job's [m12 | count] in the PINNED HOST block: the kernel stores the results there itself
    GridJob gj;
    GridParams gp;
    MatchJob mj;
};

// The job table is a kernel PARAMETER (constant bank): no dependent global loads before a CTA knows its job.
constexpr int FRAME_MAX_JOBS = 12;
struct FrameTable {
    int32_t n_jobs, n_cta, pad_[2];
    FrameJobRec job[FRAME_MAX_JOBS];
};
static_assert(sizeof(FrameTable) <= 4000, "the job table must fit the kernel parameter space");

__host__ __device__ inline size_t match_cta_smem(int n1, int n2, int best_lr) {
    const int nt = best_lr ? (n1 > n2 ? n1 : n2) : n2;
    return static_cast<size_t>(nt) * 32 + 2 * GRID_ROW_THREADS * 4;
}

// One direction of StVO::matchNNR (matching.cpp:41-61) for rows [row0, row0 + nr) of q against the nt train rows t
// (shared memory).  A row is split over S threads (column slices) whose packed top-2 keys meet in shared memory.  Keys
// are (distance << 16 | train index): unsigned min = lowest index among equal distances, the K-slot insertion order of
// cv::BFMatcher::knnMatch.  FORWARD: accepted rows are written to out (the in/out vector) and counted; else out
// receives the full matches_21 slice (-1 where rejected).
template <bool FORWARD>
__device__ __forceinline__ void match_rows_direction(const uint4 *__restrict__ q, int row0, int nr, const uint4 *t, int nt, float nnr,
                                                     int32_t *out, int32_t *count, uint32_t *part0, uint32_t *part1) {
    const int tid = threadIdx.x, NT = GRID_ROW_THREADS;
    bool acc = false;
    if (nr > 0) {
        const int S = min(NT / nr, 32);       // threads per row
        const int cps = (nt + S - 1) / S;      // columns per slice
        uint32_t b0 = KEY32_ABSENT, b1 = KEY32_ABSENT;
        if (tid < nr * S) {
            const int row = row0 + tid % nr, s = tid / nr;
            const Desc a = load_desc(q, row);
            const int j1 = min(nt, (s + 1) * cps);
            int j = s * cps;
            for (; j + 4 <= j1; j += 4) { // four independent distances in flight
                uint32_t k[4];
#pragma unroll
                for (int v = 0; v < 4; ++v)
                    k[v] = (static_cast<uint32_t>(hamming256_csa4(a, t[2 * (j + v)], t[2 * (j + v) + 1])) << 16) | static_cast<uint32_t>(j + v);
                const uint32_t m = min(min(k[0], k[1]), min(k[2], k[3]));
                if (m < b1) { // keys are distinct: a block whose minimum does not beat b1 cannot change the top-2
#pragma unroll
                    for (int v = 0; v < 4; ++v) top2_insert(b0, b1, k[v]);
                }
            }
            for (; j < j1; ++j)
                top2_insert(b0, b1, (static_cast<uint32_t>(hamming256_csa4(a, t[2 * j], t[2 * j + 1])) << 16) | static_cast<uint32_t>(j));
        }
        part0[tid] = b0;
        part1[tid] = b1;
        __syncthreads();
        if (tid < nr) {
            b0 = KEY32_ABSENT;
            b1 = KEY32_ABSENT;
            for (int s = 0; s < S; ++s) {
                top2_insert(b0, b1, part0[s * nr + tid]);
                top2_insert(b0, b1, part1[s * nr + tid]);
            }
            // matching.cpp:54 -- float arithmetic; a row without a second neighbour is not accepted
            acc = b1 != KEY32_ABSENT &&
                  static_cast<float>(static_cast<int>(b0 >> 16)) < __fmul_rn(static_cast<float>(static_cast<int>(b1 >> 16)), nnr);
            const int32_t idx = static_cast<int32_t>(b0 & 0xFFFFu);
            if (FORWARD) {
                if (acc) out[row0 + tid] = idx;
            } else {
                out[row0 + tid] = acc ? idx : -1;
            }
        }
    }
    if (FORWARD) {
        const unsigned m = __ballot_sync(0xFFFFFFFFu, acc);
        if ((tid & 31) == 0 && m) atomicAdd(count, __popc(m));
    }
}

// StVO::match (matching.cpp:63-91) of one frame-sized job.  The job's CTAs are independent: CTA c takes a block of
// query rows of one direction (the train side of that direction staged in shared memory).  The LAST CTA of the job to
// finish (a counter in global memory, no waiting) runs the mutual check over all rows.
__device__ __forceinline__ void match_job_device(const MatchJob &j, unsigned char *smem, int c, int32_t *h_io) {
    __shared__ int s_last;
    const int tid = threadIdx.x, NT = GRID_ROW_THREADS;
    const bool fwd = c < j.c12;
    const int nq = fwd ? j.n1 : j.n2, nt = fwd ? j.n2 : j.n1;
    const int cd = fwd ? j.c12 : j.n_cta - j.c12, rank = fwd ? c : c - j.c12;
    const int rpc = (nq + cd - 1) / cd;
    const int row0 = rank * rpc, nr = max(0, min(rpc, nq - row0));
    uint4 *st = reinterpret_cast<uint4 *>(smem);
    uint32_t *part0 = reinterpret_cast<uint32_t *>(smem + static_cast<size_t>(nt) * 32);
    uint32_t *part1 = part0 + NT;
    if (nr > 0) {
        stage_bytes_async(reinterpret_cast<unsigned char *>(st), fwd ? j.d2 : j.d1, static_cast<size_t>(nt) * 32);
        prefetch_l2((fwd ? j.d1 : j.d2) + 2 * (row0 + tid % nr)); // the query row this thread loads next
        stage_wait();
        __syncthreads();
    }
    if (fwd) match_rows_direction<true>(j.d1, row0, nr, st, nt, j.nnr, j.m12, j.count, part0, part1);
    else match_rows_direction<false>(j.d2, row0, nr, st, nt, j.nnr, j.m21, nullptr, part0, part1);
    __threadfence(); // this CTA's slice of matches_12 / matches_21 and its count before the arrival
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(j.done, 1) == j.n_cta - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (j.best_lr) {
        // mutual check (matching.cpp:80-86): every entry >= 0, stale ones included
        int culled = 0;
        for (int i1 = tid; i1 < j.n1; i1 += NT) {
            const int32_t i2 = __ldcg(j.m12 + i1);
            if (i2 >= 0 && (i2 >= j.n2 || __ldcg(j.m21 + i2) != i1)) {
                j.m12[i1] = -1;
                ++culled;
            }
        }
        if (culled) atomicSub(j.count, culled);
        __syncthreads();
    }
    // results straight into the pinned host block (no device -> host copy after the kernel)
    for (int i1 = tid; i1 < j.n1; i1 += NT) h_io[i1] = __ldcg(j.m12 + i1);
    if (tid == 0) h_io[j.n1] = atomicAdd(j.count, 0);
}

// Cluster dimension FRAME_CLUSTER; job k owns the CTAs [job[k].cta_begin, job[k + 1].cta_begin): exactly one cluster
// for a matchGrid job, one or more for a match job.
__global__ void __launch_bounds__(GRID_ROW_THREADS, 1)
frame_fused_kernel(const __grid_constant__ FrameTable tab) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PLM_TL(0);
    const int cta = static_cast<int>(blockIdx.x);
    int job = 0;
    while (job + 1 < tab.n_jobs && cta >= tab.job[job + 1].cta_begin) ++job; // uniform over the CTA (and its cluster)
    const FrameJobRec *r = &tab.job[job];
    const int c = cta - r->cta_begin;
    int32_t *h_io = r->h_io;
    if (r->kind == 1) {
        if (r->gp.staged) grid_rows_device<1, 2>(r->gj, r->gp, c);
        else grid_rows_device<0, 2>(r->gj, r->gp, c);
        // Results straight into the pinned host block: every CTA its own rows (their culls are its own), the count by
        // the last CTA of the cluster to arrive (the word after the count is the job's arrival counter, zeroed by the host).
        const int n1 = r->gj.n1, rpc = r->gp.rows_per_cta;
        const int32_t *m12 = r->gj.m12;
        __threadfence(); // this CTA's count updates before its arrival
        __syncthreads();
        if (threadIdx.x == 0) {
            int32_t *arrive = r->gj.count + 1;
            if (atomicAdd(arrive, 1) == FRAME_CLUSTER - 1) {
                __threadfence();
                h_io[n1] = atomicAdd(r->gj.count, 0);
            }
        }
        const int i1 = c * rpc + static_cast<int>(threadIdx.x);
        if (static_cast<int>(threadIdx.x) < rpc && i1 < n1) h_io[i1] = m12[i1];
        PLM_TL(14);
    } else {
        const MatchJob mj = r->mj;
        match_job_device(mj, smem_raw, c, h_io);
        PLM_TL(15);
    }
}

} // namespace plm
