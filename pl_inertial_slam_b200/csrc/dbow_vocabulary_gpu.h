// Drop-in for the reference's bag-of-words vocabulary type.  include/mapHandler.h:70 declares
//     typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> Vocabulary;
// and MapHandler calls dbow_voc_p / dbow_voc_l .load(), .transform(features, BowVector) and .score(v1, v2)
// (src/mapHandler.cpp:50-52, :3125, :3133, :3150, :3158, :3176, :3198, :3218-3219, :3230-3231).
// PLM::GpuVocabulary derives from that very class (header-only, like DBoW2's template), so replacing the typedef by
//     typedef PLM::GpuVocabulary Vocabulary;
// is the whole integration: load() / create() keep DBoW2's code and then mirror the tree to the device,
// transform() (virtual in DBoW2) runs plm_bow_transform, and score() -- non-virtual, but it only forwards to the
// protected m_scoring_object -- reaches plm_bow_score through a GeneralScoring subclass.  scoreAll() is the batch
// form for the loop of insertKFBowVectorP / L / PL (one launch for all earlier keyframes).  Results are
// bit-identical to DBoW2's (same fp64 summation order; include/plmatch.h).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include <DBoW2/BowVector.h>
#include <DBoW2/FORB.h>
#include <DBoW2/ScoringObject.h>
#include <DBoW2/TemplatedVocabulary.h>

#include "plmatch.h"

namespace PLM {

inline void bow_throw(int st, const char *where) {
    if (st != PLM_OK) throw std::runtime_error(std::string("[plmatch] ") + where + ": " + plm_status_string(st) + " -- " + plm_last_error());
}

// BowVector (std::map, ascending word id) <-> the flat (ids, vals) arrays of the C ABI
inline void bow_flatten(const DBoW2::BowVector &v, std::vector<uint32_t> &ids, std::vector<double> &vals) {
    for (DBoW2::BowVector::const_iterator it = v.begin(); it != v.end(); ++it) {
        ids.push_back(it->first);
        vals.push_back(it->second);
    }
}

class GpuL1Scoring : public DBoW2::GeneralScoring {
public:
    virtual double score(const DBoW2::BowVector &v, const DBoW2::BowVector &w) const {
        std::vector<uint32_t> qi, di;
        std::vector<double> qv, dv;
        bow_flatten(v, qi, qv);
        bow_flatten(w, di, dv);
        const int64_t zero = 0;
        const int32_t ql = static_cast<int32_t>(qi.size()), dl = static_cast<int32_t>(di.size());
        double out = 0.0;
        bow_throw(plm_bow_score(NULL, qi.data(), qv.data(), &zero, &ql, 1, di.data(), dv.data(), &zero, &dl, 1, &out), "plm_bow_score");
        return out;
    }
    virtual bool mustNormalize(DBoW2::LNorm &norm) const {
        norm = DBoW2::L1;
        return true;
    }
};

class GpuVocabulary : public DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> {
    typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> Base;
    plm_voc *m_dev;

    GpuVocabulary(const GpuVocabulary &);            // the device copy is not shared
    GpuVocabulary &operator=(const GpuVocabulary &);

public:
    GpuVocabulary(int k = 10, int L = 5, DBoW2::WeightingType weighting = DBoW2::TF_IDF, DBoW2::ScoringType scoring = DBoW2::L1_NORM)
        : Base(k, L, weighting, scoring), m_dev(NULL) {}
    virtual ~GpuVocabulary() { plm_voc_destroy(m_dev); }

    // DBoW2 builds / reads the tree, then it is mirrored to the device
    virtual void load(const cv::FileStorage &fs, const std::string &name = "vocabulary") {
        Base::load(fs, name);
        sync();
    }
    using Base::load;
    virtual void create(const std::vector<std::vector<cv::Mat> > &training_features) {
        Base::create(training_features);
        sync();
    }
    using Base::create;

    // Flatten m_nodes (TemplatedVocabulary.h:275-307, :383-408) into the arrays of plm_voc_create.  Call again after
    // anything that edits the tree (stopWords() changes weights).
    void sync() {
        plm_voc_destroy(m_dev);
        m_dev = NULL;
        const size_t n = m_nodes.size();
        if (n == 0) return;
        std::vector<int32_t> child_start(n + 1, 0), child_ids, word(n, -1);
        std::vector<uint8_t> desc(32 * n, 0);
        std::vector<double> weight(n, 0.0);
        for (size_t i = 0; i < n; ++i) {
            const Node &nd = m_nodes[i];
            for (size_t c = 0; c < nd.children.size(); ++c) child_ids.push_back(static_cast<int32_t>(nd.children[c]));
            child_start[i + 1] = static_cast<int32_t>(child_ids.size());
            if (!nd.descriptor.empty()) std::memcpy(&desc[32 * i], nd.descriptor.ptr<unsigned char>(), 32);
            weight[i] = nd.weight;
            if (i > 0 && nd.isLeaf()) word[i] = static_cast<int32_t>(nd.word_id);
        }
        if (child_ids.empty()) child_ids.push_back(0);
        bow_throw(plm_voc_create(NULL, static_cast<int>(n), child_start.data(), child_ids.data(), desc.data(), weight.data(),
                                 word.data(), static_cast<int>(m_weighting), static_cast<int>(m_scoring), &m_dev),
                  "plm_voc_create");
        delete m_scoring_object; // score() forwards to this object (TemplatedVocabulary.h:1179-1183)
        m_scoring_object = new GpuL1Scoring();
    }

    // TemplatedVocabulary::transform(features, BowVector) (:1045-1101)
    virtual void transform(const std::vector<cv::Mat> &features, DBoW2::BowVector &v) const {
        v.clear();
        if (empty() || features.empty() || !m_dev) return;
        const int n = static_cast<int>(features.size());
        std::vector<uint8_t> desc(32 * static_cast<size_t>(n));
        for (int i = 0; i < n; ++i) std::memcpy(&desc[32 * static_cast<size_t>(i)], features[i].ptr<unsigned char>(), 32);
        const int32_t set_start[2] = {0, n};
        std::vector<uint32_t> ids(n);
        std::vector<double> vals(n);
        int32_t len = 0;
        bow_throw(plm_bow_transform(m_dev, desc.data(), n, 32, set_start, 1, ids.data(), vals.data(), &len), "plm_bow_transform");
        for (int i = 0; i < len; ++i) v.insert(v.end(), DBoW2::BowVector::value_type(ids[i], vals[i]));
    }
    using Base::transform;

    // score(query, *database[j]) for every j in one launch: the loop of insertKFBowVectorP / L (mapHandler.cpp:3129-3137)
    void scoreAll(const DBoW2::BowVector &query, const std::vector<const DBoW2::BowVector *> &database, std::vector<double> &out) const {
        out.assign(database.size(), 0.0);
        if (database.empty()) return;
        std::vector<uint32_t> qi, di;
        std::vector<double> qv, dv;
        bow_flatten(query, qi, qv);
        std::vector<int64_t> ds(database.size());
        std::vector<int32_t> dl(database.size());
        for (size_t j = 0; j < database.size(); ++j) {
            ds[j] = static_cast<int64_t>(di.size());
            bow_flatten(*database[j], di, dv);
            dl[j] = static_cast<int32_t>(di.size() - static_cast<size_t>(ds[j]));
        }
        const int64_t zero = 0;
        const int32_t ql = static_cast<int32_t>(qi.size());
        bow_throw(plm_bow_score(NULL, qi.data(), qv.data(), &zero, &ql, 1, di.data(), dv.data(), ds.data(), dl.data(),
                                static_cast<int>(database.size()), out.data()),
                  "plm_bow_score");
    }
};

} // namespace PLM
