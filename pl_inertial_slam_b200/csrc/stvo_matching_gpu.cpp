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

#include <stdexcept>
#include <string>
#include <unordered_set>
#include <vector>

#include <opencv2/core.hpp>

#include "config.h"
#include "gridStructure.h"
#include "plmatch.h"

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
    // it only returns a reference to the cell's list, hence the const_cast.  One list walk per cell -- the portable
    // alternative, get(x, y, {0,0}x{0,0}, unordered_set), costs a hash-set build per cell (3072 per call).
    StVO::GridStructure &g = const_cast<StVO::GridStructure &>(grid);
    for (int x = 0; x < cols; ++x)
        for (int y = 0; y < rows; ++y) {
            const std::list<int> &cell = g.at(x, y);
            csr.cell_items.insert(csr.cell_items.end(), cell.begin(), cell.end());
            csr.cell_start[static_cast<size_t>(x) * rows + y + 1] = static_cast<int32_t>(csr.cell_items.size());
        }
#endif
}

} // namespace

namespace StVO {

int matchNNR(const cv::Mat &desc1, const cv::Mat &desc2, float nnr, std::vector<int> &matches_12) {
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

int matchGrid(const std::vector<point_2d> &points1, const cv::Mat &desc1, const GridStructure &grid, const cv::Mat &desc2,
              const GridWindow &w, std::vector<int> &matches_12) {
    if (points1.size() != static_cast<size_t>(desc1.rows))
        throw std::runtime_error("[matchGrid] Each point needs a corresponding descriptor!");
    check_desc(desc1);
    check_desc(desc2);
    matches_12.resize(desc1.rows, -1);

    std::vector<int32_t> xy(points1.size() * 2);
    for (size_t i = 0; i < points1.size(); ++i) {
        xy[2 * i] = points1[i].first;
        xy[2 * i + 1] = points1[i].second;
    }
    GridCsr csr;
    flatten(grid, csr);
    const int32_t win[4] = {w.width.first, w.width.second, w.height.first, w.height.second};
    int matches = 0;
    const int st = plm_match_grid_points(nullptr, xy.data(), rows_of(desc1), desc1.rows, static_cast<size_t>(desc1.step),
                                         csr.cell_start.data(), csr.cell_items.data(), grid.rows, grid.cols, rows_of(desc2),
                                         desc2.rows, static_cast<size_t>(desc2.step), win, Config::minRatio12P(),
                                         Config::bestLRMatches() ? 1 : 0, matches_12.data(), &matches);
    throw_status(st, "matchGrid");
    return matches;
}

int matchGrid(const std::vector<line_2d> &lines1, const cv::Mat &desc1, const GridStructure &grid, const cv::Mat &desc2,
              const std::vector<std::pair<double, double>> &directions2, const GridWindow &w, std::vector<int> &matches_12) {
    if (lines1.size() != static_cast<size_t>(desc1.rows))
        throw std::runtime_error("[matchGrid] Each line needs a corresponding descriptor!");
    check_desc(desc1);
    check_desc(desc2);
    matches_12.resize(desc1.rows, -1);

    std::vector<int32_t> xyxy(lines1.size() * 4);
    for (size_t i = 0; i < lines1.size(); ++i) {
        xyxy[4 * i] = lines1[i].first.first;
        xyxy[4 * i + 1] = lines1[i].first.second;
        xyxy[4 * i + 2] = lines1[i].second.first;
        xyxy[4 * i + 3] = lines1[i].second.second;
    }
    // the reference indexes directions2[i2] for every candidate i2 < desc2.rows (:221)
    std::vector<double> dirs(static_cast<size_t>(desc2.rows) * 2, 0.0);
    for (size_t i = 0; i < directions2.size() && i < static_cast<size_t>(desc2.rows); ++i) {
        dirs[2 * i] = directions2[i].first;
        dirs[2 * i + 1] = directions2[i].second;
    }
    GridCsr csr;
    flatten(grid, csr);
    const int32_t win[4] = {w.width.first, w.width.second, w.height.first, w.height.second};
    int matches = 0;
    const int st = plm_match_grid_lines(nullptr, xyxy.data(), rows_of(desc1), desc1.rows, static_cast<size_t>(desc1.step),
                                        csr.cell_start.data(), csr.cell_items.data(), grid.rows, grid.cols, rows_of(desc2),
                                        desc2.rows, static_cast<size_t>(desc2.step), dirs.data(), Config::lineSimTh(), win,
                                        Config::minRatio12P(), Config::bestLRMatches() ? 1 : 0, matches_12.data(), &matches);
    throw_status(st, "matchGrid");
    return matches;
}

} // namespace StVO
