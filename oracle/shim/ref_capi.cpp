// extern "C" entry points over the UNMODIFIED reference sources
//   /root/reference/stvo-pl/src/{matching,gridStructure,lineIterator}.cpp
// so that tests / bench can drive StVO::matchNNR / match / distance / matchGrid x2 /
// getLineCoords through ctypes.  Buffers use the same flat layouts as include/plmatch.h.
// Test infrastructure only.
#include <opencv2/core.hpp>

#include <list>
#include <stdexcept>
#include <utility>
#include <vector>

#include "config.h"
#include "gridStructure.h"
#include "matching.h"

#define PLREF_API extern "C" __attribute__((visibility("default")))

namespace {

cv::Mat wrap(const uint8_t *d, int n, size_t step) {
    return cv::Mat(n, 32, CV_8U, const_cast<uint8_t *>(d), step);
}

void fill_grid(StVO::GridStructure &grid, const int32_t *cell_start, const int32_t *cell_items) {
    for (int x = 0; x < grid.cols; x++)
        for (int y = 0; y < grid.rows; y++) {
            const int c = x * grid.rows + y;
            for (int k = cell_start[c]; k < cell_start[c + 1]; k++) grid.at(x, y).push_back(cell_items[k]);
        }
}

} // namespace

PLREF_API void plref_set_config(int best_lr, int lr_parallel, double min_ratio_12p, double line_sim_th) {
    Config::bestLRMatches() = best_lr != 0;
    Config::lrInParallel() = lr_parallel != 0;
    Config::minRatio12P() = min_ratio_12p;
    Config::lineSimTh() = line_sim_th;
}

PLREF_API int plref_distance(const uint8_t *a, const uint8_t *b) {
    return StVO::distance(wrap(a, 1, 32), wrap(b, 1, 32));
}

// returns 0 and *n_matches, or -1 when the reference threw std::runtime_error
PLREF_API int plref_match_nnr(const uint8_t *d1, int n1, size_t step1, const uint8_t *d2, int n2,
                              size_t step2, float nnr, int32_t *m12, int *n_matches) {
    std::vector<int> v(m12, m12 + n1);
    try {
        *n_matches = StVO::matchNNR(wrap(d1, n1, step1), wrap(d2, n2, step2), nnr, v);
    } catch (const std::runtime_error &) {
        return -1;
    }
    for (int i = 0; i < n1; i++) m12[i] = v[i];
    return 0;
}

PLREF_API int plref_match(const uint8_t *d1, int n1, size_t step1, const uint8_t *d2, int n2,
                          size_t step2, float nnr, int32_t *m12, int *n_matches) {
    std::vector<int> v(m12, m12 + n1);
    try {
        *n_matches = StVO::match(wrap(d1, n1, step1), wrap(d2, n2, step2), nnr, v);
    } catch (const std::runtime_error &) {
        return -1;
    }
    for (int i = 0; i < n1; i++) m12[i] = v[i];
    return 0;
}

PLREF_API int plref_match_grid_points(const int32_t *xy, const uint8_t *d1, int n1, size_t step1,
                                      const int32_t *cell_start, const int32_t *cell_items, int rows,
                                      int cols, const uint8_t *d2, int n2, size_t step2,
                                      const int32_t *win, int32_t *m12, int *n_matches) {
    try {
        StVO::GridStructure grid(rows, cols);
        fill_grid(grid, cell_start, cell_items);
        std::vector<StVO::point_2d> pts(n1);
        for (int i = 0; i < n1; i++) pts[i] = std::make_pair(xy[2 * i], xy[2 * i + 1]);
        StVO::GridWindow w;
        w.width = std::make_pair(win[0], win[1]);
        w.height = std::make_pair(win[2], win[3]);
        std::vector<int> v(m12, m12 + n1);
        *n_matches = StVO::matchGrid(pts, wrap(d1, n1, step1), grid, wrap(d2, n2, step2), w, v);
        for (int i = 0; i < n1; i++) m12[i] = v[i];
    } catch (const std::runtime_error &) {
        return -1;
    }
    return 0;
}

PLREF_API int plref_match_grid_lines(const int32_t *xyxy, const uint8_t *d1, int n1, size_t step1,
                                     const int32_t *cell_start, const int32_t *cell_items, int rows,
                                     int cols, const uint8_t *d2, int n2, size_t step2,
                                     const double *dirs2, const int32_t *win, int32_t *m12,
                                     int *n_matches) {
    try {
        StVO::GridStructure grid(rows, cols);
        fill_grid(grid, cell_start, cell_items);
        std::vector<StVO::line_2d> lines(n1);
        for (int i = 0; i < n1; i++)
            lines[i] = std::make_pair(std::make_pair(xyxy[4 * i], xyxy[4 * i + 1]),
                                      std::make_pair(xyxy[4 * i + 2], xyxy[4 * i + 3]));
        std::vector<std::pair<double, double>> dirs(n2);
        for (int i = 0; i < n2; i++) dirs[i] = std::make_pair(dirs2[2 * i], dirs2[2 * i + 1]);
        StVO::GridWindow w;
        w.width = std::make_pair(win[0], win[1]);
        w.height = std::make_pair(win[2], win[3]);
        std::vector<int> v(m12, m12 + n1);
        *n_matches = StVO::matchGrid(lines, wrap(d1, n1, step1), grid, wrap(d2, n2, step2), dirs, w, v);
        for (int i = 0; i < n1; i++) m12[i] = v[i];
    } catch (const std::runtime_error &) {
        return -1;
    }
    return 0;
}

// ---- timing at the StVO:: signature level (the same harness is linked against the reference's matching.cpp and
// against the GPU drop-in, so the two numbers compare what a caller of StVO::matchGrid / match sees: for the drop-in
// the GridStructure -> CSR flattening, staging, copies and the synchronisation are inside) ----
#include <algorithm>
#include <chrono>

namespace {
template <class F> double median_us(F &&f, int reps) {
    std::vector<double> t;
    for (int i = 0; i < 3; i++) f();
    for (int i = 0; i < reps; i++) {
        auto a = std::chrono::steady_clock::now();
        f();
        t.push_back(std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - a).count());
    }
    std::sort(t.begin(), t.end());
    return t.empty() ? 0.0 : t[t.size() / 2];
}
} // namespace

PLREF_API double plref_time_match(const uint8_t *d1, int n1, size_t step1, const uint8_t *d2, int n2, size_t step2, float nnr, int reps) {
    const cv::Mat a = wrap(d1, n1, step1), b = wrap(d2, n2, step2);
    return median_us([&] {
        std::vector<int> v;
        StVO::match(a, b, nnr, v);
    }, reps);
}

PLREF_API double plref_time_match_grid(int is_lines, const int32_t *coords, const uint8_t *d1, int n1, size_t step1, const int32_t *cell_start,
                                       const int32_t *cell_items, int rows, int cols, const uint8_t *d2, int n2, size_t step2,
                                       const double *dirs2, const int32_t *win, int reps) {
    StVO::GridStructure grid(rows, cols);
    fill_grid(grid, cell_start, cell_items);
    StVO::GridWindow w;
    w.width = std::make_pair(win[0], win[1]);
    w.height = std::make_pair(win[2], win[3]);
    const cv::Mat a = wrap(d1, n1, step1), b = wrap(d2, n2, step2);
    if (!is_lines) {
        std::vector<StVO::point_2d> pts(n1);
        for (int i = 0; i < n1; i++) pts[i] = std::make_pair(coords[2 * i], coords[2 * i + 1]);
        return median_us([&] {
            std::vector<int> v;
            StVO::matchGrid(pts, a, grid, b, w, v);
        }, reps);
    }
    std::vector<StVO::line_2d> lines(n1);
    for (int i = 0; i < n1; i++)
        lines[i] = std::make_pair(std::make_pair(coords[4 * i], coords[4 * i + 1]), std::make_pair(coords[4 * i + 2], coords[4 * i + 3]));
    std::vector<std::pair<double, double>> dirs(n2);
    for (int i = 0; i < n2; i++) dirs[i] = std::make_pair(dirs2[2 * i], dirs2[2 * i + 1]);
    return median_us([&] {
        std::vector<int> v;
        StVO::matchGrid(lines, a, grid, b, dirs, w, v);
    }, reps);
}

// getLineCoords (gridStructure.cpp:33-41): writes up to max_cells (x, y) pairs, returns the count
PLREF_API int plref_line_coords(double x1, double y1, double x2, double y2, int32_t *cells,
                                int max_cells) {
    std::list<std::pair<int, int>> lc;
    StVO::getLineCoords(x1, y1, x2, y2, lc);
    int n = 0;
    for (const auto &p : lc) {
        if (n < max_cells) {
            cells[2 * n] = p.first;
            cells[2 * n + 1] = p.second;
        }
        n++;
    }
    return n;
}

// matching.h:43-48 normalize, as the callers use it to build directions2 (stereoFrame.cpp:342-344)
PLREF_API void plref_normalize(double *v) {
    std::pair<double, double> p(v[0], v[1]);
    StVO::normalize(p);
    v[0] = p.first;
    v[1] = p.second;
}
