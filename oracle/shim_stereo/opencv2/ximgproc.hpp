// Stand-in for <opencv2/ximgproc.hpp>: declaration-level FastLineDetector (never called on the matching path).
#pragma once
#include "core.hpp"
namespace cv { namespace ximgproc {
class FastLineDetector {
public:
    void detect(const Mat &, std::vector<Vec4f> &) { standin_unavailable("FastLineDetector"); }
};
inline Ptr<FastLineDetector> createFastLineDetector(double = 10) {
    standin_unavailable("createFastLineDetector");
    return Ptr<FastLineDetector>();
}
}} // namespace cv::ximgproc
