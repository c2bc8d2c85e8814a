// Bag-of-words loop-candidate scoring: the reference's vendored DBoW2 on the device.
//
//   bow_transform_kernel  TemplatedVocabulary::transform(features, BowVector)
//                         (3rdparty/DBoW2/include/DBoW2/TemplatedVocabulary.h:1045-1101, descent :1196-1238,
//                         FORB::distance src/DBoW2/FORB.cpp:78-100, BowVector::addWeight / addIfNotExist /
//                         normalize src/DBoW2/BowVector.cpp:31-81)
//   bow_score_kernel      L1Scoring::score (src/DBoW2/ScoringObject.cpp:25-69)
// as called from MapHandler::insertKFBowVectorP / L / PL (src/mapHandler.cpp:3116-3237).
//
// Everything that is floating point is evaluated in the reference's ORDER, because fp64 addition is not
// associative and the scores feed comparisons (lookForLoopCandidates, mapHandler.cpp:3239-3299):
//   * a word's value is its weight added once per occurrence (a loop of `count` additions, not count * w);
//   * the L1 norm is one sequential sum over the words in ascending word id (std::map order);
//   * a score is one sequential sum over the common words in ascending word id.
// The parallel parts are the tree descent (one thread per feature; the node table stays L2-resident), the
// per-set sort of (word, leaf) keys in shared memory, the per-word folds, and -- for scoring -- one warp per
// database vector: 128 coalesced entries at a time are tested against a shared-memory bitmap of the query's word
// ids, the rare candidates find the query value through a shared-memory hash table, and the hit lanes' terms are
// then added in lane order, which is word order.
#pragma once
#include "plm_common.cuh"

namespace plm {

constexpr int BOW_THREADS = 256;

struct VocDev {
    const int32_t *child_start; // n_nodes + 1
    const int32_t *child_ids;
    const uint4 *node_desc;     // n_nodes x 32 B
    const double *node_weight;
    const int32_t *node_word;   // >= 0 on leaves
    int n_nodes;
    int weighting;              // DBoW2::WeightingType: 0 TF_IDF, 1 TF, 2 IDF, 3 BINARY
};

// TemplatedVocabulary::transform(feature, id, weight): first child with the smallest distance, down to a leaf.
__device__ __forceinline__ int bow_descend(const VocDev &v, const Desc &q) {
    int node = 0;
    int c0 = __ldg(v.child_start), c1 = __ldg(v.child_start + 1);
    while (c1 > c0) {
        int best = __ldg(v.child_ids + c0);
        int best_d = hamming256(q, load_desc(v.node_desc, best));
        for (int c = c0 + 1; c < c1; ++c) {
            const int id = __ldg(v.child_ids + c);
            const int d = hamming256(q, load_desc(v.node_desc, id));
            if (d < best_d) {
                best_d = d;
                best = id;
            }
        }
        node = best;
        c0 = __ldg(v.child_start + node);
        c1 = __ldg(v.child_start + node + 1);
    }
    return node;
}

struct BowTransformArgs {
    VocDev voc;
    const uint4 *desc;        // n_rows x 32 B
    const int32_t *set_start; // n_sets + 1: features of set s are rows set_start[s] .. set_start[s+1]-1
    int n_sets;
    int cap;                  // power of two >= the largest set (shared-memory sort width)
    uint32_t *bow_ids;        // entries of set s go to slots set_start[s] ..
    double *bow_vals;
    int32_t *bow_len;         // n_sets
};

inline size_t bow_transform_smem(int cap) { return size_t(cap) * (8 + 8 + 4) + 64; }

// One CTA per descriptor set.
__global__ void __launch_bounds__(BOW_THREADS) bow_transform_kernel(BowTransformArgs a) {
    extern __shared__ __align__(16) unsigned char bow_smem[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(bow_smem);    // cap: (word << 32 | leaf), sorted
    double *vals = reinterpret_cast<double *>(bow_smem + size_t(a.cap) * 8);        // cap: values in word order
    uint32_t *ids = reinterpret_cast<uint32_t *>(bow_smem + size_t(a.cap) * 16);    // cap
    __shared__ int s_scan[BOW_THREADS];
    __shared__ double s_norm;
    const int tid = threadIdx.x;
    for (int s = blockIdx.x; s < a.n_sets; s += gridDim.x) {
        const int lo = __ldg(a.set_start + s);
        const int n = min(max(__ldg(a.set_start + s + 1) - lo, 0), a.cap);
        // 1. tree descent, one feature per thread
        for (int f = tid; f < a.cap; f += BOW_THREADS) {
            unsigned long long key = KEY64_ABSENT;
            if (f < n && a.voc.n_nodes > 1) {
                const int leaf = bow_descend(a.voc, load_desc(a.desc, static_cast<long long>(lo) + f));
                if (__ldg(a.voc.node_weight + leaf) > 0.0) // stopped words are dropped (:1074)
                    key = make_key64(static_cast<uint32_t>(__ldg(a.voc.node_word + leaf)), static_cast<uint32_t>(leaf));
            }
            keys[f] = key;
        }
        __syncthreads();
        // 2. bitonic sort of the keys (ascending word id == std::map order; absent keys sink to the end)
        for (int k = 2; k <= a.cap; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < a.cap; i += BOW_THREADS) {
                    const int p = i ^ j;
                    if (p > i) {
                        const unsigned long long x = keys[i], y = keys[p];
                        const bool up = (i & k) == 0;
                        if ((x > y) == up) {
                            keys[i] = y;
                            keys[p] = x;
                        }
                    }
                }
                __syncthreads();
            }
        }
        // 3. one entry per distinct word: every thread owns a contiguous slice of the sorted keys
        const int per = (a.cap + BOW_THREADS - 1) / BOW_THREADS;
        const int i0 = tid * per, i1 = min(i0 + per, a.cap);
        int heads = 0;
        for (int i = i0; i < i1; ++i) {
            const unsigned long long k = keys[i];
            if (k != KEY64_ABSENT && (i == 0 || (keys[i - 1] >> 32) != (k >> 32))) ++heads;
        }
        s_scan[tid] = heads;
        __syncthreads();
        for (int off = 1; off < BOW_THREADS; off <<= 1) { // inclusive Hillis-Steele scan of the head counts
            const int v = tid >= off ? s_scan[tid - off] : 0;
            __syncthreads();
            s_scan[tid] += v;
            __syncthreads();
        }
        const int len = s_scan[BOW_THREADS - 1];
        int pos = s_scan[tid] - heads;
        for (int i = i0; i < i1; ++i) {
            const unsigned long long k = keys[i];
            if (k == KEY64_ABSENT || (i > 0 && (keys[i - 1] >> 32) == (k >> 32))) continue;
            const double w = __ldg(a.voc.node_weight + static_cast<uint32_t>(k));
            double v = w; // insert(id, w)
            if (a.voc.weighting <= 1) // TF_IDF / TF: addWeight once per further occurrence
                for (int j = i + 1; j < a.cap && (keys[j] >> 32) == (k >> 32); ++j) v = __dadd_rn(v, w);
            ids[pos] = static_cast<uint32_t>(k >> 32);
            vals[pos] = v;
            ++pos;
        }
        __syncthreads();
        // 4. BowVector::normalize(L1): sequential sum in word order
        if (tid == 0) {
            double norm = 0.0;
            for (int i = 0; i < len; ++i) norm = __dadd_rn(norm, fabs(vals[i]));
            s_norm = norm;
        }
        __syncthreads();
        const double norm = s_norm;
        for (int i = tid; i < len; i += BOW_THREADS) {
            a.bow_ids[lo + i] = ids[i];
            a.bow_vals[lo + i] = norm > 0.0 ? __ddiv_rn(vals[i], norm) : vals[i];
        }
        if (tid == 0) a.bow_len[s] = len;
        __syncthreads();
    }
}

struct BowScoreArgs {
    const uint32_t *q_ids; // query vectors: entries of query q at q_start[q] .. + q_len[q]
    const double *q_vals;
    const long long *q_start;
    const int32_t *q_len;
    int n_q;
    const uint32_t *db_ids; // database vectors, same layout
    const double *db_vals;
    const long long *db_start;
    const int32_t *db_len;
    int n_db;
    int q_cap;        // >= the longest query (shared-memory staging)
    int table_slots;  // power of two >= 2 * q_cap: open-addressing hash table word id -> position in the query
    int bitmap_words; // 32-bit words of the query-membership bitmap in shared memory (0: none, hash table only)
    double *scores;   // n_q x n_db: scores[q * n_db + j] = score(query q, db j)
};

inline size_t bow_score_smem(int q_cap, int table_slots, int bitmap_words) {
    return size_t(q_cap) * 8 + size_t(table_slots) * 8 + size_t(bitmap_words) * 4 + 16;
}

__device__ __forceinline__ uint32_t bow_hash(uint32_t id, int slots) { return (id * 2654435761u) & static_cast<uint32_t>(slots - 1); }

// grid = (CTAs over the database, queries); a warp per database vector, 128 coalesced entries in flight.  Two
// keyframes share only a handful of the vocabulary's 10^5 - 10^6 words, so membership of a database entry in the
// query is first tested against a BITMAP of the query's word ids in shared memory (one LDS); the rare candidates
// find the query value through a shared-memory hash table (1-2 probes) and only then load the database value.
// The kernel is then a coalesced stream over the database ids.  The hit lanes' terms are accumulated in lane
// order = ascending word id = the reference's summation order.
__global__ void __launch_bounds__(1024) bow_score_kernel(BowScoreArgs a) {
    extern __shared__ __align__(16) unsigned char bow_smem[];
    double *qv = reinterpret_cast<double *>(bow_smem);
    uint2 *table = reinterpret_cast<uint2 *>(bow_smem + size_t(a.q_cap) * 8);
    uint32_t *bitmap = reinterpret_cast<uint32_t *>(table + a.table_slots);
    const int q = blockIdx.y;
    const long long qs = __ldg(a.q_start + q);
    const int qn = min(__ldg(a.q_len + q), a.q_cap);
    for (int i = threadIdx.x; i < a.bitmap_words; i += blockDim.x) bitmap[i] = 0u;
    for (int i = threadIdx.x; i < a.table_slots; i += blockDim.x) table[i] = make_uint2(0xFFFFFFFFu, 0u);
    __syncthreads();
    const uint32_t bitmap_bits = static_cast<uint32_t>(a.bitmap_words) * 32u;
    for (int i = threadIdx.x; i < qn; i += blockDim.x) {
        const uint32_t id = __ldg(a.q_ids + qs + i);
        qv[i] = __ldg(a.q_vals + qs + i);
        if (id < bitmap_bits) atomicOr(bitmap + (id >> 5), 1u << (id & 31u));
        uint32_t h = bow_hash(id, a.table_slots); // ids of a BowVector are distinct: every insert claims an empty slot
        while (atomicCAS(&table[h].x, 0xFFFFFFFFu, id) != 0xFFFFFFFFu) h = (h + 1) & static_cast<uint32_t>(a.table_slots - 1);
        table[h].y = static_cast<uint32_t>(i);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps = blockDim.x >> 5;
    for (int j = blockIdx.x * warps + (threadIdx.x >> 5); j < a.n_db; j += gridDim.x * warps) {
        const long long ds = __ldg(a.db_start + j);
        const int dn = __ldg(a.db_len + j);
        double score = 0.0;
        for (int base = 0; base < dn; base += 128) {
            uint32_t id[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int e = base + 32 * c + lane;
                id[c] = e < dn ? __ldg(a.db_ids + ds + e) : 0xFFFFFFFFu; // no word has this id
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int e = base + 32 * c + lane;
                bool maybe = id[c] != 0xFFFFFFFFu;
                if (maybe && id[c] < bitmap_bits) maybe = (bitmap[id[c] >> 5] >> (id[c] & 31u)) & 1u;
                double term = 0.0;
                bool hit = false;
                if (maybe) {
                    uint32_t h = bow_hash(id[c], a.table_slots);
                    uint2 t = table[h];
                    while (t.x != id[c] && t.x != 0xFFFFFFFFu) {
                        h = (h + 1) & static_cast<uint32_t>(a.table_slots - 1);
                        t = table[h];
                    }
                    if (t.x == id[c]) {
                        const double vi = qv[t.y], wi = __ldg(a.db_vals + ds + e);
                        term = __dsub_rn(__dsub_rn(fabs(__dsub_rn(vi, wi)), fabs(vi)), fabs(wi));
                        hit = true;
                    }
                }
                unsigned mask = __ballot_sync(0xFFFFFFFFu, hit);
                while (mask) { // common words in ascending id: the reference's summation order
                    const int src = __ffs(mask) - 1;
                    score = __dadd_rn(score, __shfl_sync(0xFFFFFFFFu, term, src));
                    mask &= mask - 1;
                }
            }
        }
        if (lane == 0) a.scores[static_cast<size_t>(q) * a.n_db + j] = __ddiv_rn(-score, 2.0);
    }
}

} // namespace plm
