// Minimal stand-in for the OpenCV 3.4 C++ API surface that the reference's
// src/ORBextractor.cc + include/ORBextractor.h touch -- TEST INFRASTRUCTURE ONLY.
//
// Purpose: compile the reference extractor UNMODIFIED, from where it lies under
// /root/reference, without OpenCV (absent from this image), so the oracle's
// restatement of its control flow can be checked against the real code
// (oracle/_ref/libref_orb.so, built by oracle/Makefile `ref`).  The image
// primitives (FAST, resize, GaussianBlur, fastAtan2, cvRound) are the oracle's
// own models, which tests pin against cv2 4.13.0; everything else here is
// container plumbing (Mat views, Point/Rect/KeyPoint PODs).
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iterator>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "orb_oracle.h"

typedef unsigned char uchar;

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5  // only named by code paths the oracle never runs (DBoW2 FORB::toMat32F)

inline int cvRound(double v) { return orb_oracle::cv_round_d(v); }
inline int cvRound(float v) { return orb_oracle::cv_round_f(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { return (int)std::floor(v); }
inline int cvFloor(float v) { return (int)std::floor(v); }
inline int cvCeil(double v) { return (int)std::ceil(v); }
inline int cvCeil(float v) { return (int)std::ceil(v); }

namespace cv {

class Exception : public std::runtime_error {
public:
    explicit Exception(const std::string& m) : std::runtime_error(m) {}
};

template <typename T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    template <typename U>
    Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
    Point_& operator*=(float s) {
        x = (T)(x * s);
        y = (T)(y * s);
        return *this;
    }
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

template <typename T>
struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

template <typename T>
struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T _x, T _y, T w, T h) : x(_x), y(_y), width(w), height(h) {}
};
typedef Rect_<int> Rect;

struct KeyPoint {
    Point2f pt;
    float size;
    float angle;
    float response;
    int octave;
    int class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0,
             int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;
    uchar* data = nullptr;

    Mat() {}
    Mat(Size sz, int type) { create(sz.height, sz.width, type); }
    Mat(int r, int c, int type) { create(r, c, type); }

    void create(int r, int c, int /*type*/) {
        if (data && r == rows && c == cols) return;
        buf_ = std::make_shared<std::vector<uchar>>((size_t)r * c);
        rows = r;
        cols = c;
        step = (size_t)c;
        data = buf_->data();
    }
    static Mat zeros(int r, int c, int type) {
        Mat m;
        m.create(r, c, type);
        std::memset(m.data, 0, (size_t)r * c);
        return m;
    }
    void release() { *this = Mat(); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return CV_8UC1; }
    size_t step1() const { return step; }
    template <typename T>
    T& at(int r, int c) { return *(T*)(data + (size_t)r * step + c); }
    template <typename T>
    const T& at(int r, int c) const { return *(const T*)(data + (size_t)r * step + c); }
    template <typename T>
    T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <typename T>
    const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }

    Mat operator()(const Rect& r) const {
        if (r.width < 0 || r.height < 0 || r.x < 0 || r.y < 0 || r.x + r.width > cols || r.y + r.height > rows)
            throw Exception("Mat::operator()(Rect): roi outside the matrix");
        Mat m;
        m.buf_ = buf_;
        m.rows = r.height;
        m.cols = r.width;
        m.step = step;
        m.data = data + (size_t)r.y * step + r.x;
        return m;
    }
    Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
    Mat clone() const {
        Mat m;
        if (!data) return m;
        m.create(rows, cols, CV_8UC1);
        for (int y = 0; y < rows; ++y) memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, cols);
        return m;
    }
    orb_oracle::Image view() const { return orb_oracle::Image{data, rows, cols, (int)step}; }

private:
    std::shared_ptr<std::vector<uchar>> buf_;
};

class _InputArray {
public:
    _InputArray(const Mat& m) : m_(&m) {}
    bool empty() const { return m_->empty(); }
    Mat getMat() const { return *m_; }

private:
    const Mat* m_;
};
class _OutputArray {
public:
    _OutputArray(Mat& m) : m_(&m) {}
    void create(int r, int c, int type) const { m_->create(r, c, type); }
    void create(Size sz, int type) const { m_->create(sz.height, sz.width, type); }
    void release() const { m_->release(); }
    Mat getMat() const { return *m_; }
    Mat& ref() const { return *m_; }

private:
    Mat* m_;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };
enum { INTER_LINEAR = 1 };

inline float fastAtan2(float y, float x) { return orb_oracle::fast_atan2(y, x); }

inline void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nms = true) {
    assert(nms);
    (void)nms;
    Mat img = image.getMat();
    keypoints.clear();
    if (img.empty()) return;
    std::vector<orb_oracle::FastPoint> pts;
    orb_oracle::fast9_16_nms(img.view(), threshold, pts);
    for (const auto& p : pts) keypoints.push_back(KeyPoint((float)p.x, (float)p.y, 7.f, -1, (float)p.score));
}

inline void resize(InputArray src, OutputArray dst, Size dsize, double = 0, double = 0, int = INTER_LINEAR) {
    Mat s = src.getMat();
    dst.create(dsize, CV_8UC1);
    Mat d = dst.getMat();
    orb_oracle::resize_linear_u8(s.view(), d.data, d.rows, d.cols, (int)d.step);
}

inline void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sx, double sy, int border) {
    assert(ksize.width == 7 && ksize.height == 7 && sx == 2 && sy == 2 && border == BORDER_REFLECT_101);
    (void)ksize; (void)sx; (void)sy; (void)border;
    Mat s = src.getMat().clone();
    dst.create(s.rows, s.cols, CV_8UC1);
    Mat d = dst.getMat();
    orb_oracle::gaussian_blur7_u8(s.view(), d.data, (int)d.step);
}

inline void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int type) {
    assert((type & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
    (void)type;
    Mat s = src.getMat().clone();  // src may be an ROI of dst
    dst.create(s.rows + top + bottom, s.cols + left + right, CV_8UC1);
    Mat d = dst.getMat();
    auto refl = [](int p, int n) {
        if (n == 1) return 0;
        while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
        return p;
    };
    for (int y = 0; y < d.rows; ++y) {
        const uchar* S = s.data + (size_t)refl(y - top, s.rows) * s.step;
        uchar* D = d.data + (size_t)y * d.step;
        for (int x = 0; x < d.cols; ++x) D[x] = S[refl(x - left, s.cols)];
    }
}

// cv::FileStorage / cv::FileNode: named by DBoW2's YAML save()/load() (TemplatedVocabulary.h:1456-1610), which the
// oracle never calls (vocabularies travel as the fork's text format, loadFromTextFile).  Syntax only.
class FileNode {
public:
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    size_t size() const { return 0; }
    operator int() const { return 0; }
    operator double() const { return 0; }
    operator std::string() const { return std::string(); }
};
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage(const std::string&, int) {}
    bool isOpened() const { return false; }
    FileNode operator[](const std::string&) const { return FileNode(); }
    template <typename T>
    FileStorage& operator<<(const T&) { return *this; }
};

struct KeyPointsFilter {  // only referenced by the dead ComputeKeyPointsOld (ORBextractor.cc:423)
    static void retainBest(std::vector<KeyPoint>& kps, int n) {
        if ((int)kps.size() > n) {
            std::stable_sort(kps.begin(), kps.end(),
                             [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
            kps.resize(n);
        }
    }
};

}  // namespace cv
