// Minimal stand-in for <opencv2/core.hpp>: just enough surface for
// /root/reference/stvo-pl/src/matching.cpp to compile UNMODIFIED (OpenCV C++ headers are not
// installed in this image).  Test infrastructure only.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0

namespace cv {

enum NormTypes { NORM_L2 = 4, NORM_HAMMING = 6 };

class Mat {
public:
    int rows = 0, cols = 0;
    unsigned char *data = nullptr;
    size_t step = 0;

    Mat() {}
    Mat(int rows_, int cols_, int /*type*/, void *data_, size_t step_ = 0)
        : rows(rows_), cols(cols_), data(static_cast<unsigned char *>(data_)),
          step(step_ ? step_ : static_cast<size_t>(cols_)) {}

    Mat row(int r) const { return Mat(1, cols, CV_8U, data + static_cast<size_t>(r) * step, step); }
    template <typename T> T *ptr(int r = 0) { return reinterpret_cast<T *>(data + static_cast<size_t>(r) * step); }
    template <typename T> const T *ptr(int r = 0) const { return reinterpret_cast<const T *>(data + static_cast<size_t>(r) * step); }
    bool empty() const { return rows == 0 || cols == 0 || data == nullptr; }
};

struct DMatch {
    int queryIdx = -1, trainIdx = -1, imgIdx = -1;
    float distance = 0.f;
    DMatch() {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(0), distance(d) {}
};

template <typename T> using Ptr = std::shared_ptr<T>;

} // namespace cv
