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

#ifdef PLM_STVO_GPU
// The product's one-launch frame (StVO::GpuFrame, pl_inertial_slam_b200/csrc/stvo_gpu_frame.h) behind the same harness:
// stereo matchGrid for points and lines + temporal match for points and lines of ONE frame, recorded with the
// reference's own argument types (GridStructure, GridWindow, cv::Mat, std::vector) and executed by run().  Descriptor
// rows are 32 bytes apart.  reps > 0 also returns the median wall time of the whole frame -- GridStructure
// flattening, staging, the copy in, the launch and the synchronisation inside the number.
#include "stvo_gpu_frame.h"
PLREF_API int plref_gpu_frame(const int32_t *xy, const uint8_t *pd1, int np1, const int32_t *p_cs, const int32_t *p_ci, const uint8_t *pd2,
                              int np2, const int32_t *p_win, const int32_t *xyxy, const uint8_t *ld1, int nl1, const int32_t *l_cs,
                              const int32_t *l_ci, const uint8_t *ld2, int nl2, const double *dirs2, const int32_t *l_win, int rows,
                              int cols, const uint8_t *tp1, int ntp1, const uint8_t *tp2, int ntp2, const uint8_t *tl1, int ntl1,
                              const uint8_t *tl2, int ntl2, float nnr, int32_t *m_sp, int32_t *m_sl, int32_t *m_tp, int32_t *m_tl,
                              int32_t *counts, int reps, double *median_out) {
    StVO::GridStructure grid_p(rows, cols), grid_l(rows, cols);
    fill_grid(grid_p, p_cs, p_ci);
    fill_grid(grid_l, l_cs, l_ci);
    StVO::GridWindow wp, wl;
    wp.width = std::make_pair(p_win[0], p_win[1]);
    wp.height = std::make_pair(p_win[2], p_win[3]);
    wl.width = std::make_pair(l_win[0], l_win[1]);
    wl.height = std::make_pair(l_win[2], l_win[3]);
    const cv::Mat a_p = wrap(pd1, np1, 32), b_p = wrap(pd2, np2, 32), a_l = wrap(ld1, nl1, 32), b_l = wrap(ld2, nl2, 32);
    const cv::Mat t_p1 = wrap(tp1, ntp1, 32), t_p2 = wrap(tp2, ntp2, 32), t_l1 = wrap(tl1, ntl1, 32), t_l2 = wrap(tl2, ntl2, 32);
    std::vector<StVO::point_2d> pts(np1);
    for (int i = 0; i < np1; i++) pts[i] = std::make_pair(xy[2 * i], xy[2 * i + 1]);
    std::vector<StVO::line_2d> lines(nl1);
    for (int i = 0; i < nl1; i++)
        lines[i] = std::make_pair(std::make_pair(xyxy[4 * i], xyxy[4 * i + 1]), std::make_pair(xyxy[4 * i + 2], xyxy[4 * i + 3]));
    std::vector<std::pair<double, double>> dirs(nl2);
    for (int i = 0; i < nl2; i++) dirs[i] = std::make_pair(dirs2[2 * i], dirs2[2 * i + 1]);
    std::vector<int> v_sp, v_sl, v_tp, v_tl;
    int n_sp = 0, n_sl = 0, n_tp = 0, n_tl = 0;
    auto frame = [&] {
        v_sp.assign(m_sp, m_sp + np1); // in/out vectors as the caller handed them over
        v_sl.assign(m_sl, m_sl + nl1);
        v_tp.assign(m_tp, m_tp + ntp1);
        v_tl.assign(m_tl, m_tl + ntl1);
        StVO::GpuFrame f;
        f.matchGrid(pts, a_p, grid_p, b_p, wp, v_sp, n_sp);
        f.matchGrid(lines, a_l, grid_l, b_l, dirs, wl, v_sl, n_sl);
        f.match(t_p1, t_p2, nnr, v_tp, n_tp);
        f.match(t_l1, t_l2, nnr, v_tl, n_tl);
        f.run();
    };
    try {
        if (reps > 0 && median_out) *median_out = median_us(frame, reps);
        frame();
    } catch (const std::exception &) {
        return -1;
    }
    std::copy(v_sp.begin(), v_sp.end(), m_sp);
    std::copy(v_sl.begin(), v_sl.end(), m_sl);
    std::copy(v_tp.begin(), v_tp.end(), m_tp);
    std::copy(v_tl.begin(), v_tl.end(), m_tl);
    counts[0] = n_sp; counts[1] = n_sl; counts[2] = n_tp; counts[3] = n_tl;
    return 0;
}
#endif
