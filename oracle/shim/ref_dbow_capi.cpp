// extern "C" entry points over the reference's vendored DBoW2 (/root/reference/3rdparty/DBoW2, compiled
// UNMODIFIED): TemplatedVocabulary<FORB::TDescriptor, FORB> -- the `Vocabulary` of include/mapHandler.h:70 --
// with its own create() (hierarchical k-means++), transform() and score().  A subclass reaches the
// protected tree so that it can be exported to / filled from the flat arrays of include/plmatch.h.
// Test infrastructure only.
#include <opencv2/core.hpp>

#include <cstdint>
#include <vector>

#include <DBoW2/BowVector.h>
#include <DBoW2/FORB.h>
#include <DBoW2/TemplatedVocabulary.h>

#define PLREF_API extern "C" __attribute__((visibility("default")))

#ifdef PLREF_DBOW_GPU
// the same harness over the product's drop-in vocabulary type (pl_inertial_slam_b200/csrc/dbow_vocabulary_gpu.h)
#include "dbow_vocabulary_gpu.h"
#endif

namespace {

#ifdef PLREF_DBOW_GPU
typedef PLM::GpuVocabulary Vocabulary;
#else
typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> Vocabulary; // mapHandler.h:70
#endif

class Voc : public Vocabulary {
public:
    Voc(int k, int L, DBoW2::WeightingType w, DBoW2::ScoringType s) : Vocabulary(k, L, w, s) {}

    int n_nodes() const { return static_cast<int>(m_nodes.size()); }
    int n_children() const {
        size_t n = 0;
        for (const Node &nd : m_nodes) n += nd.children.size();
        return static_cast<int>(n);
    }
    void to_flat(int32_t *child_start, int32_t *child_ids, uint8_t *desc, double *weight, int32_t *word) const {
        int pos = 0;
        for (size_t i = 0; i < m_nodes.size(); i++) {
            const Node &nd = m_nodes[i];
            child_start[i] = pos;
            for (DBoW2::NodeId c : nd.children) child_ids[pos++] = static_cast<int32_t>(c);
            if (!nd.descriptor.empty()) std::memcpy(desc + 32 * i, nd.descriptor.ptr<unsigned char>(), 32);
            else std::memset(desc + 32 * i, 0, 32);
            weight[i] = nd.weight;
            word[i] = nd.isLeaf() && i > 0 ? static_cast<int32_t>(nd.word_id) : -1;
        }
        child_start[m_nodes.size()] = pos;
    }
    void from_flat(int n, const int32_t *child_start, const int32_t *child_ids, const uint8_t *desc, const double *weight,
                   const int32_t *word) {
        m_nodes.clear();
        m_words.clear();
        m_nodes.resize(n);
        int n_words = 0;
        for (int i = 0; i < n; i++) {
            Node &nd = m_nodes[i];
            nd.id = i;
            nd.weight = weight[i];
            for (int c = child_start[i]; c < child_start[i + 1]; c++) {
                nd.children.push_back(child_ids[c]);
                m_nodes[child_ids[c]].parent = i;
            }
            nd.descriptor = cv::Mat(1, 32, CV_8U, const_cast<uint8_t *>(desc) + 32 * static_cast<size_t>(i)).clone();
            if (word[i] >= 0) {
                nd.word_id = word[i];
                n_words = std::max(n_words, word[i] + 1);
            }
        }
        m_words.resize(n_words, nullptr);
        for (int i = 0; i < n; i++)
            if (word[i] >= 0) m_words[word[i]] = &m_nodes[i];
#ifdef PLREF_DBOW_GPU
        sync();
#endif
    }
#ifdef PLREF_DBOW_GPU
    using Vocabulary::scoreAll;
#endif
};

std::vector<cv::Mat> rows_of(const uint8_t *desc, int n, size_t step) {
    std::vector<cv::Mat> v;
    v.reserve(n);
    for (int i = 0; i < n; i++) v.push_back(cv::Mat(1, 32, CV_8U, const_cast<uint8_t *>(desc) + static_cast<size_t>(i) * step, step));
    return v;
}

DBoW2::BowVector bow_of(const uint32_t *ids, const double *vals, int n) {
    DBoW2::BowVector v;
    for (int i = 0; i < n; i++) v.insert(v.end(), DBoW2::BowVector::value_type(ids[i], vals[i]));
    return v;
}

} // namespace

// Vocabulary::create on training sets (set s = rows set_start[s] .. set_start[s+1]-1): the reference's own
// hierarchical k-means++ and idf weights.  `seed` pins DUtils::Random (create() seeds from the clock otherwise).
PLREF_API void *plref_voc_create(const uint8_t *desc, const int32_t *set_start, int n_sets, int k, int L, int weighting,
                                 int scoring, int seed) {
    DUtils::Random::SeedRandOnce(seed);
    DUtils::Random::SeedRand(seed);
    std::vector<std::vector<cv::Mat>> training(n_sets);
    for (int s = 0; s < n_sets; s++)
        training[s] = rows_of(desc + 32 * static_cast<size_t>(set_start[s]), set_start[s + 1] - set_start[s], 32);
    Voc *v = new Voc(k, L, static_cast<DBoW2::WeightingType>(weighting), static_cast<DBoW2::ScoringType>(scoring));
    v->create(training);
    return v;
}

PLREF_API void *plref_voc_from_flat(int n_nodes, const int32_t *child_start, const int32_t *child_ids, const uint8_t *desc,
                                    const double *weight, const int32_t *word, int k, int L, int weighting, int scoring) {
    Voc *v = new Voc(k, L, static_cast<DBoW2::WeightingType>(weighting), static_cast<DBoW2::ScoringType>(scoring));
    v->from_flat(n_nodes, child_start, child_ids, desc, weight, word);
    return v;
}

PLREF_API void plref_voc_destroy(void *h) { delete static_cast<Voc *>(h); }
PLREF_API int plref_voc_n_nodes(void *h) { return static_cast<Voc *>(h)->n_nodes(); }
PLREF_API int plref_voc_n_children(void *h) { return static_cast<Voc *>(h)->n_children(); }
PLREF_API int plref_voc_n_words(void *h) { return static_cast<int>(static_cast<Voc *>(h)->size()); }
PLREF_API void plref_voc_export(void *h, int32_t *child_start, int32_t *child_ids, uint8_t *desc, double *weight, int32_t *word) {
    static_cast<Voc *>(h)->to_flat(child_start, child_ids, desc, weight, word);
}

// Vocabulary::transform(features, BowVector) as called at src/mapHandler.cpp:3125,3150,3176,3198 -> number of
// entries written to ids / vals (word order; capacity n).
PLREF_API int plref_voc_transform(void *h, const uint8_t *desc, int n, size_t step, uint32_t *ids, double *vals) {
    DBoW2::BowVector v;
    static_cast<Voc *>(h)->transform(rows_of(desc, n, step), v);
    int i = 0;
    for (DBoW2::BowVector::const_iterator it = v.begin(); it != v.end(); ++it, ++i) {
        ids[i] = it->first;
        vals[i] = it->second;
    }
    return i;
}

#ifdef PLREF_DBOW_GPU
// GpuVocabulary::scoreAll: query (ids1, vals1) against n_db vectors laid out back to back in (ids2, vals2), lengths len2
PLREF_API void plref_voc_score_all(void *h, const uint32_t *ids1, const double *vals1, int n1, const uint32_t *ids2,
                                   const double *vals2, const int32_t *len2, int n_db, double *out) {
    std::vector<DBoW2::BowVector> db(n_db);
    std::vector<const DBoW2::BowVector *> ptrs(n_db);
    size_t pos = 0;
    for (int j = 0; j < n_db; j++) {
        db[j] = bow_of(ids2 + pos, vals2 + pos, len2[j]);
        ptrs[j] = &db[j];
        pos += len2[j];
    }
    std::vector<double> res;
    static_cast<Voc *>(h)->scoreAll(bow_of(ids1, vals1, n1), ptrs, res);
    for (int j = 0; j < n_db; j++) out[j] = res[j];
}
#endif

// Vocabulary::score(v1, v2) (src/mapHandler.cpp:3133,3158,3218-3219)
PLREF_API double plref_voc_score(void *h, const uint32_t *ids1, const double *vals1, int n1, const uint32_t *ids2,
                                 const double *vals2, int n2) {
    return static_cast<Voc *>(h)->score(bow_of(ids1, vals1, n1), bow_of(ids2, vals2, n2));
}
