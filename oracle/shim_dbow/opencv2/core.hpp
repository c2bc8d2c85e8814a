// Minimal stand-in for <opencv2/core.hpp>: just enough surface for the reference's vendored DBoW2
// (/root/reference/3rdparty/DBoW2: TemplatedVocabulary.h, FORB.cpp, BowVector.cpp, ScoringObject.cpp,
// FeatureVector.cpp) to compile UNMODIFIED.  cv::Mat here owns its storage (DBoW2 clones and creates
// descriptors); cv::FileStorage / cv::FileNode exist only so that the vocabulary's virtual save / load
// compile -- they are never opened (the vocabulary blob is absent from the mount; trees are built with
// the reference's own create() or filled from flat arrays, see shim/ref_dbow_capi.cpp).
// Test infrastructure only.
#pragma once
#include <math.h>

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstddef>
#include <iostream>
#include <limits>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_32F 5

namespace cv {

class Mat {
    std::shared_ptr<std::vector<unsigned char>> buf_;

public:
    int rows = 0, cols = 0;
    unsigned char *data = nullptr;
    size_t step = 0;

    Mat() {}
    // non-owning view on caller memory (a descriptor row)
    Mat(int r, int c, int /*type*/, void *d, size_t s = 0)
        : rows(r), cols(c), data(static_cast<unsigned char *>(d)), step(s ? s : static_cast<size_t>(c)) {}

    void create(int r, int c, int type) {
        const size_t es = (type == CV_32F) ? 4 : 1;
        buf_ = std::make_shared<std::vector<unsigned char>>(static_cast<size_t>(r) * c * es, 0);
        rows = r;
        cols = c;
        step = static_cast<size_t>(c) * es;
        data = buf_->data();
    }
    static Mat zeros(int r, int c, int type) {
        Mat m;
        m.create(r, c, type);
        return m;
    }
    Mat clone() const {
        Mat m;
        m.create(rows, cols, CV_8U);
        for (int r = 0; r < rows; r++) std::memcpy(m.data + static_cast<size_t>(r) * m.step, data + static_cast<size_t>(r) * step, cols);
        return m;
    }
    void release() {
        buf_.reset();
        rows = cols = 0;
        data = nullptr;
        step = 0;
    }
    bool empty() const { return rows == 0 || cols == 0 || data == nullptr; }
    template <typename T> T *ptr(int r = 0) { return reinterpret_cast<T *>(data + static_cast<size_t>(r) * step); }
    template <typename T> const T *ptr(int r = 0) const { return reinterpret_cast<const T *>(data + static_cast<size_t>(r) * step); }
    void convertTo(Mat &, int) const { std::abort(); } // FORB::toMat32F: not on the path
};

class FileNode {
public:
    FileNode operator[](const char *) const { return FileNode(); }
    FileNode operator[](const std::string &) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    size_t size() const { return 0; }
    operator int() const { return 0; }
    operator double() const { return 0.0; }
    operator std::string() const { return std::string(); }
};

class FileStorage {
public:
    enum Mode { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const std::string &, int) {}
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const std::string &) const { return FileNode(); }
    FileNode operator[](const char *) const { return FileNode(); }
};
template <typename T> inline FileStorage &operator<<(FileStorage &fs, const T &) { return fs; }

} // namespace cv
