// Stand-in for <opencv2/core.hpp> (OpenCV C++ headers are not installed in this image): an OWNING,
// reference-counted cv::Mat with the handful of operations that the reference's
//   stvo-pl/src/{stereoFrame,pinholeStereoCamera,matching,gridStructure,lineIterator}.cpp
// touch, so that those files compile UNMODIFIED (oracle/Makefile target `stereo`).  Semantics kept from
// OpenCV: copies are shallow, row() is a view, push_back() appends rows (adopting type/cols when empty),
// copyTo() is deep and releases the destination when the source is empty.  Feature detectors are
// declaration-level stand-ins that throw if called (extraction is outside the hot path).
// Test infrastructure only.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#define CV_CN_SHIFT 3
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_16SC2 CV_MAKETYPE(CV_16S, 2)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_RGB2GRAY 7
#define CV_GRAY2BGR 8
#define CV_BGRA2BGR 1

namespace cv {

enum NormTypes { NORM_L2 = 4, NORM_HAMMING = 6 };
enum { INTER_LINEAR = 1, CALIB_ZERO_DISPARITY = 1024 };

inline size_t standin_elem_size(int type) {
    static const size_t depth_bytes[8] = {1, 1, 2, 2, 4, 4, 8, 2};
    return depth_bytes[type & 7] * static_cast<size_t>((type >> CV_CN_SHIFT) + 1);
}

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename U> Point_(const Point_<U> &o) : x(static_cast<T>(o.x)), y(static_cast<T>(o.y)) {}
};
typedef Point_<int> Point;
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
};

template <typename T, int N> struct Vec {
    T val[N];
    Vec() : val() {}
    T &operator[](int i) { return val[i]; }
    const T &operator[](int i) const { return val[i]; }
    T &operator()(int i) { return val[i]; }
    const T &operator()(int i) const { return val[i]; }
};
typedef Vec<float, 4> Vec4f;

struct MatSize {
    int rows, cols;
    MatSize() : rows(0), cols(0) {}
    bool operator==(const MatSize &o) const { return rows == o.rows && cols == o.cols; }
    bool operator!=(const MatSize &o) const { return !(*this == o); }
};

class Mat {
public:
    int rows = 0, cols = 0;
    unsigned char *data = nullptr;
    size_t step = 0;
    MatSize size;

    Mat() {}
    Mat(int rows_, int cols_, int type) { create(rows_, cols_, type); }
    // external (non-owning) data, like cv::Mat(rows, cols, type, void*, step)
    Mat(int rows_, int cols_, int type, void *data_, size_t step_ = 0)
        : rows(rows_), cols(cols_), data(static_cast<unsigned char *>(data_)),
          step(step_ ? step_ : static_cast<size_t>(cols_) * standin_elem_size(type)), type_(type) { sync(); }

    void create(int rows_, int cols_, int type) {
        type_ = type;
        rows = rows_;
        cols = cols_;
        step = static_cast<size_t>(cols_) * standin_elem_size(type);
        buf_ = std::make_shared<std::vector<unsigned char>>(static_cast<size_t>(rows_) * step, 0);
        data = buf_->empty() ? nullptr : buf_->data();
        sync();
    }
    void release() { *this = Mat(); }

    int type() const { return type_; }
    int channels() const { return (type_ >> CV_CN_SHIFT) + 1; }
    bool empty() const { return rows == 0 || cols == 0 || data == nullptr; }
    size_t elemSize() const { return standin_elem_size(type_); }

    Mat row(int r) const {
        Mat m;
        m.rows = 1;
        m.cols = cols;
        m.step = step;
        m.type_ = type_;
        m.data = data + static_cast<size_t>(r) * step;
        m.buf_ = buf_;
        m.sync();
        return m;
    }
    template <typename T> T *ptr(int r = 0) { return reinterpret_cast<T *>(data + static_cast<size_t>(r) * step); }
    template <typename T> const T *ptr(int r = 0) const { return reinterpret_cast<const T *>(data + static_cast<size_t>(r) * step); }
    template <typename T> T &at(int r, int c) { return ptr<T>(r)[c]; }
    template <typename T> const T &at(int r, int c) const { return ptr<T>(r)[c]; }

    // Mat::push_back(const Mat&): rows of m appended; an empty matrix adopts m's type and width
    void push_back(const Mat &m) {
        if (m.empty()) return;
        if (empty()) {
            type_ = m.type_;
            cols = m.cols;
            rows = 0;
        }
        if (m.cols != cols || m.type_ != type_) throw std::runtime_error("[cv::Mat stand-in] push_back: size/type mismatch");
        const size_t row_bytes = static_cast<size_t>(cols) * elemSize();
        std::shared_ptr<std::vector<unsigned char>> nb = std::make_shared<std::vector<unsigned char>>(static_cast<size_t>(rows + m.rows) * row_bytes);
        for (int r = 0; r < rows; r++) std::memcpy(nb->data() + static_cast<size_t>(r) * row_bytes, data + static_cast<size_t>(r) * step, row_bytes);
        for (int r = 0; r < m.rows; r++) std::memcpy(nb->data() + static_cast<size_t>(rows + r) * row_bytes, m.data + static_cast<size_t>(r) * m.step, row_bytes);
        buf_ = nb;
        data = nb->data();
        rows += m.rows;
        step = row_bytes;
        sync();
    }
    void copyTo(Mat &dst) const {
        if (empty()) {
            dst.release();
            return;
        }
        Mat out(rows, cols, type_);
        const size_t row_bytes = static_cast<size_t>(cols) * elemSize();
        for (int r = 0; r < rows; r++) std::memcpy(out.data + static_cast<size_t>(r) * out.step, data + static_cast<size_t>(r) * step, row_bytes);
        dst = out;
    }
    void convertTo(Mat &, int) const { throw std::logic_error("[cv::Mat stand-in] convertTo is outside the matching path"); }

    static Mat eye(int r, int c, int type) {
        Mat m(r, c, type);
        for (int i = 0; i < std::min(r, c); i++) {
            if ((type & 7) == CV_64F) m.at<double>(i, i) = 1.0;
            else if ((type & 7) == CV_32F) m.at<float>(i, i) = 1.f;
        }
        return m;
    }

protected:
    int type_ = CV_8U;
    std::shared_ptr<std::vector<unsigned char>> buf_;
    void sync() {
        size.rows = rows;
        size.cols = cols;
    }
};

template <typename T> struct standin_type;
template <> struct standin_type<float> { enum { value = CV_32F }; };
template <> struct standin_type<double> { enum { value = CV_64F }; };
template <> struct standin_type<unsigned char> { enum { value = CV_8U }; };

template <typename T> class Mat_;
template <typename T> class MatCommaInitializer_ {
public:
    explicit MatCommaInitializer_(Mat_<T> *m) : m_(m), i_(0) {}
    template <typename U> MatCommaInitializer_ &operator,(U v) {
        m_->template at<T>(i_ / m_->cols, i_ % m_->cols) = static_cast<T>(v);
        i_++;
        return *this;
    }
    operator Mat() const { return *m_; }
    operator Mat_<T>() const { return *m_; }

private:
    Mat_<T> *m_;
    int i_;
};
template <typename T> class Mat_ : public Mat {
public:
    Mat_() {}
    Mat_(int r, int c) : Mat(r, c, standin_type<T>::value) {}
    static Mat_ eye(int r, int c) {
        Mat_ m(r, c);
        for (int i = 0; i < std::min(r, c); i++) m.template at<T>(i, i) = T(1);
        return m;
    }
};
// (Mat_<T>(r, c) << a, b, ...): the initializer keeps its own copy of the header; the buffer is shared.
template <typename T, typename U> inline MatCommaInitializer_<T> operator<<(const Mat_<T> &m, U v) {
    static thread_local Mat_<T> holder;
    holder = m;
    MatCommaInitializer_<T> ci(&holder);
    return (ci, v);
}

struct DMatch {
    int queryIdx = -1, trainIdx = -1, imgIdx = -1;
    float distance = 0.f;
    DMatch() {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(0), distance(d) {}
};

struct KeyPoint {
    Point2f pt;
    float size = 0.f, angle = -1.f, response = 0.f;
    int octave = 0, class_id = -1;
};

template <typename T> using Ptr = std::shared_ptr<T>;

// image-processing entry points the non-matching methods of StereoFrame / PinholeStereoCamera name
inline void standin_unavailable(const char *what) { throw std::logic_error(std::string("[OpenCV stand-in] ") + what + " is outside the matching path"); }
inline void cvtColor(const Mat &, Mat &, int, int = 0) { standin_unavailable("cvtColor"); }
inline void circle(Mat &, Point, int, const Scalar &, double = 1) { standin_unavailable("circle"); }
inline void line(Mat &, Point, Point, const Scalar &, double = 1) { standin_unavailable("line"); }
inline void remap(const Mat &, Mat &, const Mat &, const Mat &, int) { standin_unavailable("remap"); }
// rectification maps are built by PinholeStereoCamera's constructors but never read on the matching path: no-op
inline void initUndistortRectifyMap(const Mat &, const Mat &, const Mat &, const Mat &, Size, int, Mat &, Mat &) {}
inline void stereoRectify(const Mat &, const Mat &, const Mat &, const Mat &, Size, const Mat &, const Mat &, Mat &, Mat &, Mat &,
                          Mat &, Mat &, int, double) { standin_unavailable("stereoRectify"); }
namespace fisheye {
inline void initUndistortRectifyMap(const Mat &, const Mat &, const Mat &, const Mat &, Size, int, Mat &, Mat &) {}
} // namespace fisheye

class LineIterator {
public:
    LineIterator(const Mat &, Point2f, Point2f) : count(0) { standin_unavailable("cv::LineIterator"); }
    int count;
};

} // namespace cv
