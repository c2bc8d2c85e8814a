// Stand-in for <opencv2/features2d.hpp>: cv::BFMatcher::{create, knnMatch} (implemented in
// oracle/shim/bfmatcher_shim.cpp with OpenCV's batchDistance K-slot insertion rule) and a declaration-level
// cv::ORB (never called on the matching path).  Test infrastructure only.
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

class ORB {
public:
    static Ptr<ORB> create(int = 500, float = 1.2f, int = 8, int = 31, int = 0, int = 2, int = 0, int = 31, int = 20) {
        standin_unavailable("ORB::create");
        return Ptr<ORB>();
    }
    void detectAndCompute(const Mat &, const Mat &, std::vector<KeyPoint> &, Mat &, bool = false) { standin_unavailable("ORB"); }
};

} // namespace cv
