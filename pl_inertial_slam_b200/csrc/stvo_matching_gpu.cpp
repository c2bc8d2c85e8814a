// Drop-in replacement for the reference's stvo-pl/src/matching.cpp: the same five StVO:: free
// functions (stvo-pl/include/matching.h:50-60, mangled names unchanged), implemented over the C ABI
// of include/plmatch.h.  Build it INSIDE the reference tree in place of matching.cpp and link
// libplmatch.so; nothing else in StVO / PLSLAM changes.
//
// Duties kept from the reference:
//  * reads Config::bestLRMatches / minRatio12P / lineSimTh exactly where matching.cpp does
//    (:65, :122, :160, :193, :221, :241); Config::lrInParallel is irrelevant -- both directions of
//    match() are one kernel launch, no std::async;
//  * matches_12.resize(n, -1) and in/out semantics of the vector (stale entries survive, :44);
//  * honours cv::Mat::step;
//  * rethrows the reference's std::runtime_error messages (:51, :114, :185);
//  * flattens GridStructure to CSR through its public const get() only (`grid` is private,
//    gridStructure.h:54-57).  Define PLM_GRIDSTRUCTURE_FRIEND and add
//    `friend struct PlmGridAccess;` to GridStructure to skip the per-cell set round trip.
#include "matching.h"

#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <vector>

#include <opencv2/core.hpp>

#include "config.h"
#include "gridStructure.h"
#include "plmatch.h"
#include "stvo_gpu_frame.h"

#ifdef PLM_GRIDSTRUCTURE_FRIEND
namespace StVO {
struct PlmGridAccess {
    static const std::vector<std::vector<std::list<int>>> &cells(const GridStructure &g) { return g.grid; }
};
} // namespace StVO
#endif

namespace {

const unsigned char *rows_of(const cv::Mat &m) { return m.ptr<unsigned char>(); }

void check_desc(const cv::Mat &m) {
    if (m.rows > 0 && m.cols != 32) throw std::runtime_error("[plmatch] descriptors must be 32 bytes (256 bit) wide");
}

// The free functions return their count by value: inside an open frame session the library would record the address of
// a local -- use StVO::GpuFrame (stvo_gpu_frame.h) there.
void no_open_frame(const char *where) {
    if (plm_frame_active(nullptr))
        throw std::logic_error(std::string("[plmatch] StVO::") + where + " called inside an open frame session; record the call with StVO::GpuFrame");
}

void throw_status(int st, const char *where) {
    if (st == PLM_OK) return;
    if (st == PLM_E_TRAIN) throw std::runtime_error("[matchNNR] Different size for matches and descriptors!");
    if (st == PLM_E_GRID) throw std::runtime_error("[GridStructure] invalid dimension");
    throw std::runtime_error(std::string("[plmatch] ") + where + ": " + plm_status_string(st) + " -- " + plm_last_error());
}

// CSR copy of the bucket grid: cell (x, y) -> id x * rows + y (x outermost, like grid[x][y]).
struct GridCsr {
    std::vector<int32_t> cell_start, cell_items;
};

void flatten(const StVO::GridStructure &grid, GridCsr &csr) {
    const int rows = grid.rows, cols = grid.cols;
    csr.cell_start.assign(static_cast<size_t>(rows) * cols + 1, 0);
    csr.cell_items.clear();
#ifdef PLM_GRIDSTRUCTURE_FRIEND
    const auto &cells = StVO::PlmGridAccess::cells(grid);
    for (int x = 0; x < cols; ++x)
        for (int y = 0; y < rows; ++y) {
            csr.cell_items.insert(csr.cell_items.end(), cells[x][y].begin(), cells[x][y].end());
            csr.cell_start[static_cast<size_t>(x) * rows + y + 1] = static_cast<int32_t>(csr.cell_items.size());
        }
#else
    // GridStructure::at is the reference's own public accessor (gridStructure.h:48); it is not const-qualified although
    // it only returns a reference to the cell's list, hence the const_cast.  The cells of one grid column are the
    // elements of ONE std::vector<std::list<int>> (gridStructure.h:54, grid[x][y]), so &at(x, 0) + y is cell (x, y):
    // one out-of-line call per column instead of one per cell (4.4 us instead of 10-13 us for the 64 x 48 grid).  The
    // portable alternative, get(x, y, {0,0}x{0,0}, unordered_set), costs a hash-set build per cell.
    StVO::GridStructure &g = const_cast<StVO::GridStructure &>(grid);
    int32_t *cs = csr.cell_start.data();
    int32_t n = 0;
    for (int x = 0; x < cols; ++x) {
        const std::list<int> *column = &g.at(x, 0);
        for (int y = 0; y < rows; ++y) {
            const std::list<int> &cell = column[y];
            if (!cell.empty()) {
                for (std::list<int>::const_iterator it = cell.begin(); it != cell.end(); ++it) csr.cell_items.push_back(*it);
                n = static_cast<int32_t>(csr.cell_items.size());
            }
            cs[static_cast<size_t>(x) * rows + y + 1] = n;
        }
    }
#endif
}

} // namespace

namespace StVO {

int matchNNR(const cv::Mat &desc1, const cv::Mat &desc2, float nnr, std::vector<int> &matches_12) {
    no_open_frame("matchNNR");
    check_desc(desc1);
    check_desc(desc2);
    matches_12.resize(desc1.rows, -1);
    int matches = 0;
    const int st = plm_match_nnr(nullptr, rows_of(desc1), desc1.rows, static_cast<size_t>(desc1.step), rows_of(desc2),
                                 desc2.rows, static_cast<size_t>(desc2.step), nnr, matches_12.data(), &matches);
    throw_status(st, "matchNNR");
    return matches;
}

int match(const cv::Mat &desc1, const cv::Mat &desc2, float nnr, std::vector<int> &matches_12) {
    no_open_frame("match");
    check_desc(desc1);
    check_desc(desc2);
    matches_12.resize(desc1.rows, -1);
    int matches = 0;
    const int st = plm_match(nullptr, rows_of(desc1), desc1.rows, static_cast<size_t>(desc1.step), rows_of(desc2),
                             desc2.rows, static_cast<size_t>(desc2.step), nnr, Config::bestLRMatches() ? 1 : 0,
                             matches_12.data(), &matches);
    throw_status(st, "match");
    return matches;
}

// One 256-bit pair is not worth a host <-> device round trip: the same 8 x (xor, popcount) as matching.cpp:93-109 on
// the host (the batched device form is plm_hamming256).
int distance(const cv::Mat &a, const cv::Mat &b) {
    const int32_t *pa = a.ptr<int32_t>();
    const int32_t *pb = b.ptr<int32_t>();
    int dist = 0;
    for (int i = 0; i < 8; ++i) dist += __builtin_popcount(static_cast<unsigned>(pa[i] ^ pb[i]));
    return dist;
}

} // namespace StVO

namespace {

// Flat arguments of one matchGrid call (coordinates, CSR grid, directions, window) built from the reference's types.
struct GridArgs {
    std::vector<int32_t> coords;
    std::vector<double> dirs;
    GridCsr csr;
    int32_t win[4];
};

void grid_args_points(const std::vector<StVO::point_2d> &points1, const cv::Mat &desc1, const StVO::GridStructure &grid,
                      const cv::Mat &desc2, const StVO::GridWindow &w, std::vector<int> &matches_12, GridArgs &a) {
    if (points1.size() != static_cast<size_t>(desc1.rows))
        throw std::runtime_error("[matchGrid] Each point needs a corresponding descriptor!");
    check_desc(desc1);
    check_desc(desc2);
    matches_12.resize(desc1.rows, -1);
    a.coords.resize(points1.size() * 2);
    for (size_t i = 0; i < points1.size(); ++i) {
        a.coords[2 * i] = points1[i].first;
        a.coords[2 * i + 1] = points1[i].second;
    }
    flatten(grid, a.csr);
    a.win[0] = w.width.first; a.win[1] = w.width.second; a.win[2] = w.height.first; a.win[3] = w.height.second;
}

void grid_args_lines(const std::vector<StVO::line_2d> &lines1, const cv::Mat &desc1, const StVO::GridStructure &grid,
                     const cv::Mat &desc2, const std::vector<std::pair<double, double>> &directions2, const StVO::GridWindow &w,
                     std::vector<int> &matches_12, GridArgs &a) {
    if (lines1.size() != static_cast<size_t>(desc1.rows))
        throw std::runtime_error("[matchGrid] Each line needs a corresponding descriptor!");
    check_desc(desc1);
    check_desc(desc2);
    matches_12.resize(desc1.rows, -1);
    a.coords.resize(lines1.size() * 4);
    for (size_t i = 0; i < lines1.size(); ++i) {
        a.coords[4 * i] = lines1[i].first.first;
        a.coords[4 * i + 1] = lines1[i].first.second;
        a.coords[4 * i + 2] = lines1[i].second.first;
        a.coords[4 * i + 3] = lines1[i].second.second;
    }
    // the reference indexes directions2[i2] for every candidate i2 < desc2.rows (:221)
    a.dirs.assign(static_cast<size_t>(desc2.rows) * 2, 0.0);
    for (size_t i = 0; i < directions2.size() && i < static_cast<size_t>(desc2.rows); ++i) {
        a.dirs[2 * i] = directions2[i].first;
        a.dirs[2 * i + 1] = directions2[i].second;
    }
    flatten(grid, a.csr);
    a.win[0] = w.width.first; a.win[1] = w.width.second; a.win[2] = w.height.first; a.win[3] = w.height.second;
}

int call_grid_points(const GridArgs &a, const cv::Mat &desc1, const StVO::GridStructure &grid, const cv::Mat &desc2,
                     std::vector<int> &matches_12, int *n_matches) {
    return plm_match_grid_points(nullptr, a.coords.data(), rows_of(desc1), desc1.rows, static_cast<size_t>(desc1.step),
                                 a.csr.cell_start.data(), a.csr.cell_items.data(), grid.rows, grid.cols, rows_of(desc2), desc2.rows,
                                 static_cast<size_t>(desc2.step), a.win, Config::minRatio12P(),
                                 Config::bestLRMatches() ? 1 : 0, matches_12.data(), n_matches);
}

int call_grid_lines(const GridArgs &a, const cv::Mat &desc1, const StVO::GridStructure &grid, const cv::Mat &desc2,
                    std::vector<int> &matches_12, int *n_matches) {
    return plm_match_grid_lines(nullptr, a.coords.data(), rows_of(desc1), desc1.rows, static_cast<size_t>(desc1.step),
                                a.csr.cell_start.data(), a.csr.cell_items.data(), grid.rows, grid.cols, rows_of(desc2), desc2.rows,
                                static_cast<size_t>(desc2.step), a.dirs.data(), Config::lineSimTh(), a.win,
                                Config::minRatio12P(), Config::bestLRMatches() ? 1 : 0, matches_12.data(), n_matches);
}

} // namespace

namespace StVO {

int matchGrid(const std::vector<point_2d> &points1, const cv::Mat &desc1, const GridStructure &grid, const cv::Mat &desc2,
              const GridWindow &w, std::vector<int> &matches_12) {
    no_open_frame("matchGrid");
    GridArgs a;
    grid_args_points(points1, desc1, grid, desc2, w, matches_12, a);
    int matches = 0;
    throw_status(call_grid_points(a, desc1, grid, desc2, matches_12, &matches), "matchGrid");
    return matches;
}

int matchGrid(const std::vector<line_2d> &lines1, const cv::Mat &desc1, const GridStructure &grid, const cv::Mat &desc2,
              const std::vector<std::pair<double, double>> &directions2, const GridWindow &w, std::vector<int> &matches_12) {
    no_open_frame("matchGrid");
    GridArgs a;
    grid_args_lines(lines1, desc1, grid, desc2, directions2, w, matches_12, a);
    int matches = 0;
    throw_status(call_grid_lines(a, desc1, grid, desc2, matches_12, &matches), "matchGrid");
    return matches;
}

// ---- StVO::GpuFrame (stvo_gpu_frame.h): the calls of one frame as one launch -----------------------------------------
struct GpuFrame::Scratch {
    GridArgs a;
};

GpuFrame::GpuFrame() : open_(false) {
    throw_status(plm_frame_begin(nullptr), "GpuFrame");
    open_ = true;
}

GpuFrame::~GpuFrame() {
    if (open_) plm_frame_end(nullptr);
}

void GpuFrame::matchNNR(const cv::Mat &desc1, const cv::Mat &desc2, float nnr, std::vector<int> &matches_12, int &n_matches) {
    check_desc(desc1);
    check_desc(desc2);
    matches_12.resize(desc1.rows, -1);
    n_matches = 0;
    throw_status(plm_match_nnr(nullptr, rows_of(desc1), desc1.rows, static_cast<size_t>(desc1.step), rows_of(desc2), desc2.rows,
                               static_cast<size_t>(desc2.step), nnr, matches_12.data(), &n_matches), "matchNNR");
}

void GpuFrame::match(const cv::Mat &desc1, const cv::Mat &desc2, float nnr, std::vector<int> &matches_12, int &n_matches) {
    check_desc(desc1);
    check_desc(desc2);
    matches_12.resize(desc1.rows, -1);
    n_matches = 0;
    throw_status(plm_match(nullptr, rows_of(desc1), desc1.rows, static_cast<size_t>(desc1.step), rows_of(desc2), desc2.rows,
                           static_cast<size_t>(desc2.step), nnr, Config::bestLRMatches() ? 1 : 0, matches_12.data(), &n_matches), "match");
}

void GpuFrame::matchGrid(const std::vector<point_2d> &points1, const cv::Mat &desc1, const GridStructure &grid, const cv::Mat &desc2,
                         const GridWindow &w, std::vector<int> &matches_12, int &n_matches) {
    keep_.push_back(std::unique_ptr<Scratch>(new Scratch));
    GridArgs &a = keep_.back()->a;
    grid_args_points(points1, desc1, grid, desc2, w, matches_12, a);
    n_matches = 0;
    throw_status(call_grid_points(a, desc1, grid, desc2, matches_12, &n_matches), "matchGrid");
}

void GpuFrame::matchGrid(const std::vector<line_2d> &lines1, const cv::Mat &desc1, const GridStructure &grid, const cv::Mat &desc2,
                         const std::vector<std::pair<double, double>> &directions2, const GridWindow &w, std::vector<int> &matches_12,
                         int &n_matches) {
    keep_.push_back(std::unique_ptr<Scratch>(new Scratch));
    GridArgs &a = keep_.back()->a;
    grid_args_lines(lines1, desc1, grid, desc2, directions2, w, matches_12, a);
    n_matches = 0;
    throw_status(call_grid_lines(a, desc1, grid, desc2, matches_12, &n_matches), "matchGrid");
}

void GpuFrame::run() {
    if (!open_) throw std::runtime_error("[plmatch] GpuFrame::run: the frame has already run");
    open_ = false;
    const int st = plm_frame_end(nullptr);
    keep_.clear();
    throw_status(st, "GpuFrame::run");
}

} // namespace StVO
