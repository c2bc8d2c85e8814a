// Multi-GPU top-2 merge over NVLink peer memory ("K5"): the per-query top-2 exchange of the
// row-sharded keyframe database / local map done by ONE kernel instead of a gather collective plus a merge.
//
// Every rank owns one exchange buffer (cudaMalloc + CUDA IPC, mapped into every other rank's address
// space) laid out as   [parity 2][rank G][q_cap] ulonglong2 keys  |  [parity 2][rank G][blocks] uint32 flags.
// A CTA handles 128 queries:
//   push   its packed keys go straight into slot [parity][my rank][q] of EVERY rank's buffer
//          (16-byte stores over NVLink / NVSwitch; the local copy is an ordinary store);
//   signal CTA barrier, then flag [parity][my rank][block] = epoch on every rank with st.release.sys (cumulative);
//   wait   until the G flags [parity][r][block] of the OWN buffer read `epoch` (bounded spin);
//   merge  the G x 2 keys per query with unsigned min (== the reference's lowest-index tie-breaking),
//          optionally followed by the matchNNR ratio test.
// A CTA only ever waits for the SAME block index of the other ranks, which depends on nothing but that
// rank's own progress, so there is no intra-grid dependency and no deadlock for any grid size.  Buffers are
// double-buffered by epoch parity: a rank can run at most one exchange ahead of the slowest one (it cannot
// pass the wait of epoch e+1 before every rank has pushed e+1, i.e. has finished reading e).
#pragma once
#include "plm_common.cuh"

namespace plm {

constexpr int PEER_MAX_RANKS = 16;
constexpr int PEER_THREADS = 128;

struct PeerExchangeArgs {
    unsigned char *peer[PEER_MAX_RANKS]; // base of every rank's exchange buffer as mapped HERE (own included)
    int rank, world;
    int q_cap, blocks_cap;
    uint32_t epoch;                // > 0, the same on every rank, increases by one per exchange
    const ulonglong2 *local;       // n1 x 2 packed keys of this rank
    int n1;
    ulonglong2 *out;               // merged keys (may be null)
    float nnr;
    int32_t *m12;                  // optional matchNNR acceptance on the merged keys
    int32_t *count;
    int32_t *error;                // set to 1 when a wait timed out
    long long spin_limit;          // clock64 ticks
};

__host__ __device__ inline size_t peer_keys_bytes(int world, int q_cap) { return size_t(2) * world * q_cap * 16; }
__host__ __device__ inline size_t peer_buffer_bytes(int world, int q_cap, int blocks_cap) {
    return peer_keys_bytes(world, q_cap) + size_t(2) * world * blocks_cap * 4;
}

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t atom_add_acq_rel_gpu(uint32_t *p, uint32_t v) {
    uint32_t old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ ulonglong2 ld_volatile_u64x2(const ulonglong2 *p) {
    ulonglong2 v;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}

// push + signal + wait of one CTA: element q of this rank (16 bytes) goes to slot [parity][my rank][q] of every
// rank's buffer; returns false when a peer did not arrive in time.  After a true return the G slots
// [parity][r][q] of the OWN buffer hold every rank's element (read them with ld_volatile_u64x2).
__device__ __forceinline__ bool peer_push_wait(const PeerExchangeArgs &a, int bx, int q, bool active, ulonglong2 v) {
    const int par = a.epoch & 1;
    const size_t key_slot = (size_t(par) * a.world + a.rank) * a.q_cap; // [par][my rank][.]
    const size_t keys_bytes = peer_keys_bytes(a.world, a.q_cap);
    if (active) {
        for (int p = 0; p < a.world; ++p) {
            const int dst = (a.rank + p) % a.world; // start with the local copy, spread the remote stores
            reinterpret_cast<ulonglong2 *>(a.peer[dst])[key_slot + q] = v;
        }
    }
    // Release pattern of the PTX memory model: the CTA barrier orders every thread's stores before the signalling
    // threads, whose st.release.sys is cumulative over that order -- no membar.sys (it alone costs ~10 us here).
    __syncthreads();
    if (threadIdx.x < a.world) { // signal
        const int dst = (a.rank + threadIdx.x) % a.world;
        uint32_t *flags = reinterpret_cast<uint32_t *>(a.peer[dst] + keys_bytes);
        st_release_sys(flags + (size_t(par) * a.world + a.rank) * a.blocks_cap + bx, a.epoch);
    }
    __shared__ int s_fail;
    if (threadIdx.x == 0) s_fail = 0;
    __syncthreads();
    if (threadIdx.x < a.world) { // wait for block `blockIdx.x` of every rank
        const uint32_t *flags = reinterpret_cast<const uint32_t *>(a.peer[a.rank] + keys_bytes);
        const uint32_t *f = flags + (size_t(par) * a.world + threadIdx.x) * a.blocks_cap + bx;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) != a.epoch) {
            if (clock64() - t0 > a.spin_limit) {
                s_fail = 1;
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
    if (s_fail) {
        if (threadIdx.x == 0) atomicExch(a.error, 1);
        return false;
    }
    return true; // acquire pattern: ld.acquire.sys by the waiting threads, then the CTA barrier above
}

__device__ __forceinline__ const ulonglong2 *peer_slots(const PeerExchangeArgs &a) {
    return reinterpret_cast<const ulonglong2 *>(a.peer[a.rank]) + size_t(a.epoch & 1) * a.world * a.q_cap;
}

__device__ __forceinline__ void top2_exchange_merge_body(const PeerExchangeArgs &a, int bx) {
    const int q = bx * PEER_THREADS + threadIdx.x;
    const bool active = q < a.n1;
    if (!peer_push_wait(a, bx, q, active, active ? a.local[q] : make_ulonglong2(0, 0))) return;
    if (!active) return;
    const ulonglong2 *keys = peer_slots(a);
    unsigned long long b0 = KEY64_ABSENT, b1 = KEY64_ABSENT;
    for (int r = 0; r < a.world; ++r) {
        const ulonglong2 v = ld_volatile_u64x2(keys + size_t(r) * a.q_cap + q);
        top2_insert(b0, b1, v.x);
        top2_insert(b0, b1, v.y);
    }
    if (a.out) a.out[q] = make_ulonglong2(b0, b1);
    if (a.m12 && b1 != KEY64_ABSENT) {
        const float d0 = static_cast<float>(static_cast<int>(b0 >> 32));
        const float d1 = static_cast<float>(static_cast<int>(b1 >> 32));
        if (d0 < __fmul_rn(d1, a.nnr)) { // matchNNR ratio test in fp32 (matching.cpp:54)
            a.m12[q] = static_cast<int32_t>(b0 & 0xFFFFFFFFull);
            if (a.count) atomicAdd(a.count, 1);
        }
    }
}

__global__ void __launch_bounds__(PEER_THREADS) top2_exchange_merge_kernel(PeerExchangeArgs a) { top2_exchange_merge_body(a, blockIdx.x); }

// Element-wise reductions over the ranks with the same push / signal / wait: the two exchanges of the row-sharded
// matchGrid (database.py ShardedMap.match_grid).  The payload travels in 16-byte chunks (a.local /
// a.out are arrays of n1 chunks):
//   OP 0  min over ALL ranks of 2 x uint64 per chunk  -- the per-column best pairs (distance << 32 | global row)
//   OP 1  min over the LOWER ranks (r < rank) of 8 x uint16 per chunk, 0xFFFF where there is none -- the running
//         column minima that seed a shard's thresholds
template <int OP> __device__ __forceinline__ void peer_reduce_body(const PeerExchangeArgs &a, int bx) {
    const int q = bx * PEER_THREADS + threadIdx.x;
    const bool active = q < a.n1;
    if (!peer_push_wait(a, bx, q, active, active ? a.local[q] : make_ulonglong2(0, 0))) return;
    if (!active) return;
    const ulonglong2 *slots = peer_slots(a);
    if (OP == 0) {
        unsigned long long m0 = KEY64_ABSENT, m1 = KEY64_ABSENT;
        for (int r = 0; r < a.world; ++r) {
            const ulonglong2 v = ld_volatile_u64x2(slots + size_t(r) * a.q_cap + q);
            m0 = min(m0, v.x);
            m1 = min(m1, v.y);
        }
        a.out[q] = make_ulonglong2(m0, m1);
    } else {
        unsigned long long m0 = KEY64_ABSENT, m1 = KEY64_ABSENT; // 8 x 0xFFFF
        for (int r = 0; r < a.rank; ++r) {
            const ulonglong2 v = ld_volatile_u64x2(slots + size_t(r) * a.q_cap + q);
            // lane-wise unsigned 16-bit minimum of two 64-bit words (4 lanes each)
            auto min16x4 = [](unsigned long long x, unsigned long long y) {
                unsigned long long out = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned long long xa = (x >> (16 * k)) & 0xFFFFull, ya = (y >> (16 * k)) & 0xFFFFull;
                    out |= (xa < ya ? xa : ya) << (16 * k);
                }
                return out;
            };
            m0 = min16x4(m0, v.x);
            m1 = min16x4(m1, v.y);
        }
        a.out[q] = make_ulonglong2(m0, m1);
    }
}

template <int OP> __global__ void __launch_bounds__(PEER_THREADS) peer_reduce_kernel(PeerExchangeArgs a) { peer_reduce_body<OP>(a, blockIdx.x); }

// All-gather of the per-shard match vectors + sum of the per-shard counts (the last step of the row-sharded
// match / matchGrid: every rank ends with the global matches_12 vector and the global count) as one kernel.
// Every rank owns a second peer-mapped buffer:
//   [parity 2][n_rows_cap] int32 rows | [parity 2][G] int32 counts | [parity 2][G] uint32 flags | [2] uint32 done
// push   grid-stride: this rank's rows go to [parity][row_lo + i] of every rank's buffer, its count to slot
//        [parity][my rank];
// signal the LAST block of the grid to finish pushing (device-scope counter) raises flag [parity][my rank] = epoch on
//        every rank;
// wait   every block waits for the G flags of the own buffer (bounded spin).  G - 1 of them depend only on the other
//        ranks' pushes; the flag of the OWN rank is raised by the last block of THIS grid, so every block of the grid
//        must be resident at the same time: the host launches at most one block per SM (plm_dev_peer_allgather_i32),
//        which holds as long as nothing else occupies whole SMs for longer than the spin limit (then: PLM_E_PEER, never
//        a silent result);
// copy   grid-stride copy of the assembled vector into the caller's output, block 0 sums the G counts.
struct PeerGatherArgs {
    unsigned char *peer[PEER_MAX_RANKS];
    int rank, world;
    long long n_rows_cap;
    uint32_t epoch;
    const int32_t *local;       // n_local rows of this rank
    long long row_lo, n_local, n_rows;
    const int32_t *local_count; // 1 (may be null: 0)
    int32_t *out;               // n_rows
    int32_t *out_count;         // 1 (may be null)
    int32_t *error;
    long long spin_limit;
};

__host__ __device__ inline size_t peer_gather_bytes(int world, long long n_rows_cap) {
    return size_t(2) * size_t(n_rows_cap) * 4 + size_t(2) * world * 4 + size_t(2) * world * 4 + 16;
}

__device__ __forceinline__ void peer_allgather_body(const PeerGatherArgs &a, int bx, int nbx) {
    const int par = a.epoch & 1;
    const size_t rows_bytes = size_t(2) * size_t(a.n_rows_cap) * 4;
    const long long tid = static_cast<long long>(bx) * blockDim.x + threadIdx.x;
    const long long nthreads = static_cast<long long>(nbx) * blockDim.x;
    // push
    for (int p = 0; p < a.world; ++p) {
        const int dst = (a.rank + p) % a.world;
        int32_t *rows = reinterpret_cast<int32_t *>(a.peer[dst]) + size_t(par) * a.n_rows_cap + a.row_lo;
        for (long long i = tid; i < a.n_local; i += nthreads) rows[i] = a.local[i];
        if (tid == 0) {
            int32_t *counts = reinterpret_cast<int32_t *>(a.peer[dst] + rows_bytes);
            counts[par * a.world + a.rank] = a.local_count ? *a.local_count : 0;
        }
    }
    __syncthreads();
    // signal: the last block of this grid to get here (acq_rel counter: the other blocks' stores happen before the
    // last block's st.release.sys, which is cumulative)
    uint32_t *own_flags = reinterpret_cast<uint32_t *>(a.peer[a.rank] + rows_bytes + size_t(2) * a.world * 4);
    uint32_t *done = own_flags + 2 * a.world + par;
    __shared__ int s_last, s_fail;
    if (threadIdx.x == 0) {
        s_fail = 0;
        s_last = (atom_add_acq_rel_gpu(done, 1u) == static_cast<uint32_t>(nbx) - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_last) {
        if (threadIdx.x == 0) *done = 0; // ready for the next use of this parity
        if (threadIdx.x < a.world) {
            const int dst = (a.rank + threadIdx.x) % a.world;
            uint32_t *flags = reinterpret_cast<uint32_t *>(a.peer[dst] + rows_bytes + size_t(2) * a.world * 4);
            st_release_sys(flags + par * a.world + a.rank, a.epoch);
        }
    }
    // wait
    if (threadIdx.x < a.world) {
        const uint32_t *f = own_flags + par * a.world + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) != a.epoch) {
            if (clock64() - t0 > a.spin_limit) {
                s_fail = 1;
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
    if (s_fail) {
        if (threadIdx.x == 0) atomicExch(a.error, 1);
        return;
    }
    // copy out
    const volatile int32_t *rows = reinterpret_cast<const volatile int32_t *>(a.peer[a.rank]) + size_t(par) * a.n_rows_cap;
    for (long long i = tid; i < a.n_rows; i += nthreads) a.out[i] = rows[i];
    if (tid == 0 && a.out_count) {
        const volatile int32_t *counts = reinterpret_cast<const volatile int32_t *>(a.peer[a.rank] + rows_bytes);
        int32_t sum = 0;
        for (int r = 0; r < a.world; ++r) sum += counts[par * a.world + r];
        *a.out_count = sum;
    }
}
__global__ void __launch_bounds__(256) peer_allgather_kernel(PeerGatherArgs a) { peer_allgather_body(a, blockIdx.x, gridDim.x); }

// ---- all ranks of an exchange inside ONE cooperative kernel (blockIdx.y = rank) ---------------------------------
// For boxes with fewer GPUs than ranks (tests): kernels that wait for one another must not be separate launches on one
// GPU -- nothing guarantees that they run at the same time.  A cooperative launch makes every CTA co-resident, so the
// CTAs of "rank" y can wait for the flags the CTAs of the other ranks raise.  The ranks' buffers are plain device
// allocations on the same GPU; the code path is the one the real multi-GPU launches run.
constexpr int PEER_EMU_MAX_RANKS = 4;
struct PeerEmuExchange { PeerExchangeArgs a[PEER_EMU_MAX_RANKS]; };
struct PeerEmuGather { PeerGatherArgs a[PEER_EMU_MAX_RANKS]; };
// KIND 0: top-2 exchange + merge, 1: min over all ranks (u64), 2: prefix-min over the lower ranks (u16)
template <int KIND> __global__ void __launch_bounds__(PEER_THREADS) peer_emu_exchange_kernel(PeerEmuExchange e) {
    const PeerExchangeArgs &a = e.a[blockIdx.y];
    if (KIND == 0) top2_exchange_merge_body(a, blockIdx.x);
    else if (KIND == 1) peer_reduce_body<0>(a, blockIdx.x);
    else peer_reduce_body<1>(a, blockIdx.x);
}
__global__ void __launch_bounds__(256) peer_emu_gather_kernel(PeerEmuGather e) { peer_allgather_body(e.a[blockIdx.y], blockIdx.x, gridDim.x); }

} // namespace plm
