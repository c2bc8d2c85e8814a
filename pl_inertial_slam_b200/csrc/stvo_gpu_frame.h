// Not in the reference: one launch per frame for hosts that use the StVO:: types.
//
// The reference runs the point half and the line half of a frame on two std::async threads
// (stvo-pl/src/stereoFrame.cpp:75-76, stvo-pl/src/stereoFrameHandler.cpp:142-143).  StVO::GpuFrame records the matcher
// calls of one frame -- same argument types as the free functions of matching.h:50-60, the int return value becomes an
// int & that is filled by run() -- and executes them as ONE kernel (plm_frame_begin / plm_frame_end of include/plmatch.h:
// one host -> device copy, frame_fused_kernel, results stored straight into pinned host memory).
//
//     StVO::GpuFrame f;
//     f.matchGrid(points_l, pdesc_l, grid_p, pdesc_r, w, matches_sp, n_sp);
//     f.matchGrid(lines_l, ldesc_l, grid_l, ldesc_r, directions_r, w, matches_sl, n_sl);
//     f.match(prev_pdesc, pdesc_l, nnr, matches_tp, n_tp);
//     f.match(prev_ldesc, ldesc_l, nnr, matches_tl, n_tl);
//     f.run();          // every matches_* vector and n_* is defined from here on
//
// The descriptor matrices, the match vectors and the ints must stay alive (and the vectors must not be resized) until
// run() returns; grids, windows, coordinates and directions are copied when the call is recorded.  Config:: values are
// read when a call is recorded, as matching.cpp reads them when it is called.
#pragma once
#include <memory>
#include <utility>
#include <vector>

#include <opencv2/core.hpp>

#include "gridStructure.h"
#include "matching.h"

namespace StVO {

class __attribute__((visibility("default"))) GpuFrame {
public:
    GpuFrame();
    ~GpuFrame(); // an unfinished frame is executed and its errors dropped (the recorded pointers must not dangle)
    GpuFrame(const GpuFrame &) = delete;
    GpuFrame &operator=(const GpuFrame &) = delete;

    void matchNNR(const cv::Mat &desc1, const cv::Mat &desc2, float nnr, std::vector<int> &matches_12, int &n_matches);
    void match(const cv::Mat &desc1, const cv::Mat &desc2, float nnr, std::vector<int> &matches_12, int &n_matches);
    void matchGrid(const std::vector<point_2d> &points1, const cv::Mat &desc1, const GridStructure &grid, const cv::Mat &desc2,
                   const GridWindow &w, std::vector<int> &matches_12, int &n_matches);
    void matchGrid(const std::vector<line_2d> &lines1, const cv::Mat &desc1, const GridStructure &grid, const cv::Mat &desc2,
                   const std::vector<std::pair<double, double>> &directions2, const GridWindow &w, std::vector<int> &matches_12,
                   int &n_matches);
    void run(); // throws std::runtime_error with the reference's messages where the free functions would

private:
    struct Scratch;
    std::vector<std::unique_ptr<Scratch>> keep_;
    bool open_;
};

} // namespace StVO
