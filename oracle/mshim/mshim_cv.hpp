// Test infrastructure, never shipped: a stand-in for the few cv:: types the reference's src/ORBmatcher.cc touches, so that
// the file compiles UNMODIFIED from where it lies (oracle/Makefile refmatch).  cv::Mat here is a small dense matrix of
// CV_8U (descriptor rows) or CV_32F (poses, points) elements with views, products and sums.  The bridge builds its scenes
// with identity cameras, so every product and sum the matcher forms with these is exact in any evaluation order and the
// arithmetic of OpenCV's gemm does not enter the comparison.
#pragma once
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <list>
#include <memory>
#include <stdexcept>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_32FC1 5

namespace cv {
typedef unsigned char uchar;

template <typename T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<float> Point2f;
typedef Point_<int> Point;

struct KeyPoint {
    Point2f pt;
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
};

class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;  // bytes per row
    uchar* data = nullptr;

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    void create(int r, int c, int type) {
        type_ = type;
        esz_ = type == CV_32F ? 4 : 1;
        buf_ = std::make_shared<std::vector<uchar>>((size_t)r * c * esz_ + 8, 0);
        rows = r;
        cols = c;
        step = (size_t)c * esz_;
        data = buf_->data();
    }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    static Mat ones(int r, int c, int type) {
        Mat m(r, c, type);
        for (int i = 0; i < r; ++i)
            for (int j = 0; j < c; ++j) m.at<float>(i, j) = 1.0f;
        return m;
    }
    static Mat eye(int r, int c, int type) {
        Mat m(r, c, type);
        for (int i = 0; i < r && i < c; ++i) m.at<float>(i, i) = 1.0f;
        return m;
    }
    bool empty() const { return !data || rows == 0 || cols == 0; }
    int type() const { return type_; }
    template <typename T>
    T& at(int r, int c) { return *(T*)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T>
    const T& at(int r, int c) const { return *(const T*)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T>
    T& at(int i) { return cols == 1 ? at<T>(i, 0) : at<T>(0, i); }
    template <typename T>
    const T& at(int i) const { return cols == 1 ? at<T>(i, 0) : at<T>(0, i); }
    template <typename T>
    T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <typename T>
    const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }

    Mat view(int r0, int r1, int c0, int c1) const {
        if (r0 < 0 || c0 < 0 || r1 > rows || c1 > cols || r0 > r1 || c0 > c1) throw std::out_of_range("Mat view");
        Mat m;
        m.buf_ = buf_;
        m.type_ = type_;
        m.esz_ = esz_;
        m.rows = r1 - r0;
        m.cols = c1 - c0;
        m.step = step;
        m.data = data + (size_t)r0 * step + (size_t)c0 * esz_;
        return m;
    }
    Mat row(int i) const { return view(i, i + 1, 0, cols); }
    Mat col(int j) const { return view(0, rows, j, j + 1); }
    Mat rowRange(int a, int b) const { return view(a, b, 0, cols); }
    Mat colRange(int a, int b) const { return view(0, rows, a, b); }
    Mat clone() const {
        Mat m(rows, cols, type_);
        for (int r = 0; r < rows; ++r) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * esz_);
        return m;
    }
    // 8U -> 32F or a plain copy; the destination may be the source (cv::Mat::convertTo allows it)
    void convertTo(Mat& dst, int type) const {
        Mat m(rows, cols, type);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) {
                if (type == CV_32F)
                    m.at<float>(r, c) = type_ == CV_32F ? at<float>(r, c) : (float)at<uchar>(r, c);
                else
                    m.at<uchar>(r, c) = at<uchar>(r, c);
            }
        dst = m;
    }
    // like cv::Mat::copyTo: a destination of the right shape is written in place (it may be a view), otherwise replaced
    void copyTo(Mat& dst) const {
        if (dst.data && dst.rows == rows && dst.cols == cols && dst.type_ == type_) {
            for (int r = 0; r < rows; ++r) std::memmove(dst.data + (size_t)r * dst.step, data + (size_t)r * step, (size_t)cols * esz_);
        } else {
            dst = clone();
        }
    }
    void copyTo(Mat&& dst) const { copyTo(dst); }
    Mat reshape(int) const { return *this; }  // only on the distortion path, which the bridges never take
    Mat t() const {
        Mat m(cols, rows, CV_32F);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) m.at<float>(c, r) = at<float>(r, c);
        return m;
    }
    double dot(const Mat& o) const {
        double s = 0;
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) s += (double)at<float>(r, c) * (double)o.at<float>(r, c);
        return s;
    }

private:
    std::shared_ptr<std::vector<uchar>> buf_;
    int type_ = CV_8U;
    size_t esz_ = 1;
};

inline Mat operator*(const Mat& a, const Mat& b) {
    if (a.cols != b.rows) throw std::invalid_argument("Mat product");
    Mat m(a.rows, b.cols, CV_32F);
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < b.cols; ++c) {
            double s = 0;
            for (int k = 0; k < a.cols; ++k) s += (double)a.at<float>(r, k) * (double)b.at<float>(k, c);
            m.at<float>(r, c) = (float)s;
        }
    return m;
}
template <typename F>
inline Mat mshim_map(const Mat& a, F f) {
    Mat m(a.rows, a.cols, CV_32F);
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < a.cols; ++c) m.at<float>(r, c) = f(a.at<float>(r, c));
    return m;
}
template <typename F>
inline Mat mshim_zip(const Mat& a, const Mat& b, F f) {
    if (a.rows != b.rows || a.cols != b.cols) throw std::invalid_argument("Mat sizes");
    Mat m(a.rows, a.cols, CV_32F);
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < a.cols; ++c) m.at<float>(r, c) = f(a.at<float>(r, c), b.at<float>(r, c));
    return m;
}
inline Mat operator*(double s, const Mat& a) { return mshim_map(a, [s](float v) { return (float)(s * v); }); }
inline Mat operator*(const Mat& a, double s) { return s * a; }
inline Mat operator/(const Mat& a, double s) { return mshim_map(a, [s](float v) { return (float)(v / s); }); }
inline Mat operator-(const Mat& a) { return mshim_map(a, [](float v) { return -v; }); }
inline Mat operator+(const Mat& a, const Mat& b) { return mshim_zip(a, b, [](float x, float y) { return x + y; }); }
inline Mat operator-(const Mat& a, const Mat& b) { return mshim_zip(a, b, [](float x, float y) { return x - y; }); }
inline double norm(const Mat& a) { return std::sqrt(a.dot(a)); }
enum { NORM_L1 = 2 };
inline double norm(const Mat& a, const Mat& b, int /*NORM_L1*/) {
    double s = 0;
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < a.cols; ++c) s += std::fabs((double)a.at<float>(r, c) - (double)b.at<float>(r, c));
    return s;
}
// (cv::Mat_<float>(r, c) << a, b, c): comma initialiser
template <typename T>
class Mat_ : public Mat {
public:
    Mat_(int r, int c) : Mat(r, c, CV_32F) {}
    struct Init {
        Mat_* m;
        int k;
        Init operator,(T v) {
            m->template at<T>(k / m->cols, k % m->cols) = v;
            return Init{m, k + 1};
        }
        operator Mat() const { return *m; }
    };
    Init operator<<(T v) {
        this->template at<T>(0, 0) = v;
        return Init{this, 1};
    }
};
inline void undistortPoints(const Mat&, Mat&, const Mat&, const Mat&, const Mat&, const Mat&) {
    throw std::logic_error("cv::undistortPoints is not part of the stand-in: the bridges use zero distortion");
}
}  // namespace cv
