// Minimal stand-in for <opencv2/features2d.hpp>: cv::BFMatcher::{create, knnMatch} only.
// knnMatch is implemented in oracle/shim/bfmatcher_shim.cpp with OpenCV's batchDistance
// K-slot insertion rule.  Test infrastructure only.
#pragma once
#include "core.hpp"

namespace cv {

class BFMatcher {
public:
    static Ptr<BFMatcher> create(int normType = NORM_L2, bool crossCheck = false);
    void knnMatch(const Mat &queryDescriptors, const Mat &trainDescriptors,
                  std::vector<std::vector<DMatch>> &matches, int k) const;

private:
    int normType_ = NORM_HAMMING;
    bool crossCheck_ = false;
};

} // namespace cv
