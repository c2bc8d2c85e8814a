// Drop-in replacement for the reference's src/mapFeatures.cpp (PLSLAM::MapPoint / PLSLAM::MapLine,
// include/mapFeatures.h:40-101): same constructors and member functions (mangled names unchanged), with
// updateAverageDescDir() -- the n x n Hamming medoid of the observation descriptors and the mean observation
// direction (src/mapFeatures.cpp:51-93, :121-163) -- computed by plm_med_desc (include/plmatch.h) instead of the
// host loops.  Build it INSIDE the reference tree in place of mapFeatures.cpp and link libplmatch.so
// (INTEGRATION.md 4.1).
//
// One landmark per call is the reference's granularity (addMap*Observation recomputes after every append).  The
// batch entry point below, PLSLAM::updateAverageDescDirBatch, does any number of landmarks in one launch and is
// what a keyframe insertion should use after appending its observations with the `defer` helpers.
#include "mapFeatures.h"

#include <cstring>
#include <stdexcept>
#include <string>

#include "plmatch.h"

namespace {

// Observation lists of `count` landmarks flattened into the arenas plm_med_desc takes, results written back.
template <class Landmark> void recompute(Landmark *const *lms, int count) {
    std::vector<int32_t> obs_start(static_cast<size_t>(count) + 1, 0);
    for (int l = 0; l < count; ++l) obs_start[l + 1] = obs_start[l] + static_cast<int32_t>(lms[l]->desc_list.size());
    const int n_obs = obs_start[count];
    std::vector<uint8_t> desc(static_cast<size_t>(n_obs) * 32);
    std::vector<double> dirs(static_cast<size_t>(n_obs) * 3);
    for (int l = 0; l < count; ++l) {
        const Landmark &lm = *lms[l];
        for (size_t i = 0; i < lm.desc_list.size(); ++i) {
            const size_t r = static_cast<size_t>(obs_start[l]) + i;
            std::memcpy(&desc[32 * r], lm.desc_list[i].template ptr<unsigned char>(), 32);
            for (int c = 0; c < 3; ++c) dirs[3 * r + c] = lm.dir_list[i](c);
        }
    }
    std::vector<int32_t> med_idx(count);
    std::vector<double> med_dir(static_cast<size_t>(count) * 3);
    const int st = plm_med_desc(NULL, desc.data(), n_obs, 32, dirs.data(), obs_start.data(), count, med_idx.data(), NULL,
                                med_dir.data());
    if (st != PLM_OK)
        throw std::runtime_error(std::string("[plmatch] updateAverageDescDir: ") + plm_status_string(st) + " -- " + plm_last_error());
    for (int l = 0; l < count; ++l) {
        if (med_idx[l] < 0) continue; // empty list: nothing to pick
        lms[l]->med_desc = lms[l]->desc_list[med_idx[l]];
        for (int c = 0; c < 3; ++c) lms[l]->med_obs_dir(c) = med_dir[3 * static_cast<size_t>(l) + c];
    }
}

} // namespace

namespace PLSLAM {

// ---- MapPoint (mapFeatures.h:40-68) ---------------------------------------------------------------------------
MapPoint::MapPoint(int idx_, Vector3d point3D_, Mat desc_, int kf_obs_, Vector2d obs_, Vector3d dir_, double sigma2_)
    : idx(idx_), inlier(true), point3D(point3D_), med_obs_dir(dir_), med_desc(desc_) {
    desc_list.push_back(desc_);
    obs_list.push_back(obs_);
    dir_list.push_back(dir_);
    kf_obs_list.push_back(kf_obs_);
    sigma_list.push_back(sigma2_);
}

void MapPoint::addMapPointObservation(Mat desc_, int kf_obs_, Vector2d obs_, Vector3d dir_, double sigma2_) {
    desc_list.push_back(desc_);
    obs_list.push_back(obs_);
    dir_list.push_back(dir_);
    kf_obs_list.push_back(kf_obs_);
    sigma_list.push_back(sigma2_);
    updateAverageDescDir();
}

void MapPoint::updateAverageDescDir() {
    MapPoint *self = this;
    recompute(&self, 1);
}

// ---- MapLine (mapFeatures.h:70-101) ---------------------------------------------------------------------------
MapLine::MapLine(int idx_, Vector6d line3D_, Mat desc_, int kf_obs_, Vector3d obs_, Vector3d dir_, Vector4d pts_, double sigma2_)
    : idx(idx_), inlier(true), line3D(line3D_), med_obs_dir(dir_), med_desc(desc_) {
    desc_list.push_back(desc_);
    obs_list.push_back(obs_);
    pts_list.push_back(pts_);
    dir_list.push_back(dir_);
    kf_obs_list.push_back(kf_obs_);
    sigma_list.push_back(sigma2_);
}

void MapLine::addMapLineObservation(Mat desc_, int kf_obs_, Vector3d obs_, Vector3d dir_, Vector4d pts_, double sigma2_) {
    desc_list.push_back(desc_);
    obs_list.push_back(obs_);
    pts_list.push_back(pts_);
    dir_list.push_back(dir_);
    kf_obs_list.push_back(kf_obs_);
    sigma_list.push_back(sigma2_);
    updateAverageDescDir();
}

void MapLine::updateAverageDescDir() {
    MapLine *self = this;
    recompute(&self, 1);
}

// ---- batch form (not in the reference) --------------------------------------------------------------------------
void updateAverageDescDirBatch(MapPoint *const *points, int n_points, MapLine *const *lines, int n_lines) {
    if (n_points > 0) recompute(points, n_points);
    if (n_lines > 0) recompute(lines, n_lines);
}

} // namespace PLSLAM
