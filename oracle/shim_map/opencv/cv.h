// Minimal stand-in for <opencv/cv.h>: just enough surface for /root/reference/src/mapFeatures.cpp
// to compile UNMODIFIED (cv::Mat passed by value, cv::norm(.., NORM_HAMMING) on 1 x 32 CV_8U rows).
// OpenCV's NORM_HAMMING on CV_8U is the popcount of the byte-wise xor (hal::normHamming).
// Test infrastructure only.
#pragma once
#include "../../shim/opencv2/core.hpp"

namespace cv {

inline double norm(const Mat &a, const Mat &b, int normType) {
    (void)normType; // the path only uses NORM_HAMMING
    int d = 0;
    for (int i = 0; i < a.cols; i++) d += __builtin_popcount(static_cast<unsigned>(a.data[i] ^ b.data[i]));
    return static_cast<double>(d);
}

} // namespace cv
