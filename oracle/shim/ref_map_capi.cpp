// extern "C" entry points over the UNMODIFIED reference source /root/reference/src/mapFeatures.cpp:
// a landmark is built exactly as the SLAM back end does it -- constructor with the first observation,
// then addMapPointObservation / addMapLineObservation per further observation (each runs
// updateAverageDescDir) -- and its med_desc / med_obs_dir are read back.  Test infrastructure only.
#include "mapFeatures.h"

#include <cstdint>
#include <vector>

#define PLREF_API extern "C" __attribute__((visibility("default")))

namespace {

cv::Mat row_of(const uint8_t *desc, size_t step, int r) {
    return cv::Mat(1, 32, CV_8U, const_cast<uint8_t *>(desc) + static_cast<size_t>(r) * step, step);
}

Eigen::Vector3d dir_of(const double *dirs, int r) {
    Eigen::Vector3d v;
    if (dirs)
        for (int c = 0; c < 3; c++) v(c) = dirs[3 * static_cast<size_t>(r) + c];
    return v;
}

template <typename LM> void read_back(const LM &lm, const uint8_t *desc, size_t step, int lo, int32_t *med_idx, double *med_dir) {
    *med_idx = static_cast<int32_t>((lm.med_desc.data - (desc + static_cast<size_t>(lo) * step)) / static_cast<std::ptrdiff_t>(step));
    if (med_dir)
        for (int c = 0; c < 3; c++) med_dir[c] = lm.med_obs_dir(c);
}

} // namespace

#ifdef PLREF_MAP_BATCH
// Only the product's C++ drop-in (map_features_gpu.cpp) has the batch entry point: every landmark is built with
// its first observation, the others are appended WITHOUT the per-append recomputation, then one batch call.
namespace PLSLAM {
void updateAverageDescDirBatch(MapPoint *const *points, int n_points, MapLine *const *lines, int n_lines);
}
PLREF_API void plref_med_desc_batch(int is_line, const uint8_t *desc, size_t step, const double *dirs, const int32_t *obs_start,
                                    int n_lm, int32_t *med_idx, double *med_dir) {
    std::vector<PLSLAM::MapPoint *> pts;
    std::vector<PLSLAM::MapLine *> lns;
    for (int l = 0; l < n_lm; l++) {
        const int lo = obs_start[l], n = obs_start[l + 1] - lo;
        if (n <= 0) continue;
        if (!is_line) {
            PLSLAM::MapPoint *p = new PLSLAM::MapPoint(l, Eigen::Vector3d(), row_of(desc, step, lo), 0, Eigen::Vector2d(), dir_of(dirs, lo));
            for (int i = 1; i < n; i++) {
                p->desc_list.push_back(row_of(desc, step, lo + i));
                p->dir_list.push_back(dir_of(dirs, lo + i));
            }
            pts.push_back(p);
        } else {
            PLSLAM::MapLine *q = new PLSLAM::MapLine(l, Vector6d(), row_of(desc, step, lo), 0, Eigen::Vector3d(), dir_of(dirs, lo), Eigen::Vector4d());
            for (int i = 1; i < n; i++) {
                q->desc_list.push_back(row_of(desc, step, lo + i));
                q->dir_list.push_back(dir_of(dirs, lo + i));
            }
            lns.push_back(q);
        }
    }
    PLSLAM::updateAverageDescDirBatch(pts.data(), static_cast<int>(pts.size()), lns.data(), static_cast<int>(lns.size()));
    size_t k = 0;
    for (int l = 0; l < n_lm; l++) {
        const int lo = obs_start[l], n = obs_start[l + 1] - lo;
        double *md = med_dir ? med_dir + 3 * static_cast<size_t>(l) : nullptr;
        if (n <= 0) {
            med_idx[l] = -1;
            if (md) md[0] = md[1] = md[2] = 0.0;
            continue;
        }
        if (!is_line) {
            read_back(*pts[k], desc, step, lo, med_idx + l, md);
            delete pts[k];
        } else {
            read_back(*lns[k], desc, step, lo, med_idx + l, md);
            delete lns[k];
        }
        k++;
    }
}
#endif

// is_line = 0: PLSLAM::MapPoint, 1: PLSLAM::MapLine.  Layouts as plo_med_desc (oracle/plm_oracle.c).
PLREF_API void plref_med_desc(int is_line, const uint8_t *desc, size_t step, const double *dirs, const int32_t *obs_start,
                              int n_lm, int32_t *med_idx, double *med_dir) {
    for (int l = 0; l < n_lm; l++) {
        const int lo = obs_start[l], n = obs_start[l + 1] - lo;
        double *md = med_dir ? med_dir + 3 * static_cast<size_t>(l) : nullptr;
        if (n <= 0) {
            med_idx[l] = -1;
            if (md) md[0] = md[1] = md[2] = 0.0;
            continue;
        }
        if (!is_line) {
            PLSLAM::MapPoint p(l, Eigen::Vector3d(), row_of(desc, step, lo), 0, Eigen::Vector2d(), dir_of(dirs, lo));
            for (int i = 1; i < n; i++)
                p.addMapPointObservation(row_of(desc, step, lo + i), i, Eigen::Vector2d(), dir_of(dirs, lo + i));
            read_back(p, desc, step, lo, med_idx + l, md);
        } else {
            PLSLAM::MapLine q(l, Vector6d(), row_of(desc, step, lo), 0, Eigen::Vector3d(), dir_of(dirs, lo), Eigen::Vector4d());
            for (int i = 1; i < n; i++)
                q.addMapLineObservation(row_of(desc, step, lo + i), i, Eigen::Vector3d(), dir_of(dirs, lo + i), Eigen::Vector4d());
            read_back(q, desc, step, lo, med_idx + l, md);
        }
    }
}
