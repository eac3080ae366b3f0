// ORB_SLAM2::ORBextractor over the liborb_b200 C ABI (replaces the reference's src/ORBextractor.cc).
// Error behaviour follows the reference: empty image -> silent return with outputs untouched
// (src/ORBextractor.cc:444-445); zero keypoints -> descriptors released, keypoints not cleared
// (:460-463); anything the CUDA path refuses (a shape on which the reference itself faults, a CUDA
// error) is raised as cv::Exception, the only error channel operator() has.
#include "ORBextractor.h"

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <opencv2/core/core.hpp>
#include <opencv2/imgproc/imgproc.hpp>

#include "orb_b200.h"

namespace ORB_SLAM2 {

static const int EDGE_THRESHOLD = 19;

static int g_device = -1;      // UseDevice
static int g_keep_pyramid = -1;  // KeepImagePyramid; -1: ask the environment

void ORBextractor::UseDevice(int device) { g_device = device; }
void ORBextractor::KeepImagePyramid(bool on) { g_keep_pyramid = on ? 1 : 0; }

static int device_ordinal() {
    if (g_device >= 0) return g_device;
    const char* e = getenv("ORB_B200_DEVICE");
    return e ? atoi(e) : 0;
}
static bool keep_pyramid() {
    if (g_keep_pyramid >= 0) return g_keep_pyramid != 0;
    const char* e = getenv("ORB_B200_IMAGE_PYRAMID");
    return !(e && atoi(e) == 0);
}

[[noreturn]] static void raise(const char* what) {
    throw cv::Exception(std::string("orb_b200: ") + what + ": " + orb_last_error());
}

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST),
      handle_(nullptr) {
    orb_params p;
    p.nfeatures = _nfeatures;
    p.scale_factor = _scaleFactor;
    p.nlevels = _nlevels;
    p.ini_th_fast = _iniThFAST;
    p.min_th_fast = _minThFAST;
    if (orb_extractor_create(&p, 0, 0, 1, device_ordinal(), &handle_) != ORB_OK) raise("orb_extractor_create");
    mvScaleFactor.resize(nlevels);
    mvInvScaleFactor.resize(nlevels);
    mvLevelSigma2.resize(nlevels);
    mvInvLevelSigma2.resize(nlevels);
    mnFeaturesPerLevel.resize(nlevels);
    orb_extractor_tables(handle_, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data(),
                         mnFeaturesPerLevel.data());
    mvImagePyramid.resize(nlevels);
}

ORBextractor::~ORBextractor() {
    orb_extractor_destroy(handle_);
    for (void* p : pin_) orb_host_free(p);
}

void ORBextractor::pinned(size_t which, size_t bytes) {
    if (bytes <= pinBytes_[which]) return;
    orb_host_free(pin_[which]);
    pin_[which] = nullptr;
    pinBytes_[which] = 0;
    if (orb_host_alloc(bytes + bytes / 4, &pin_[which]) != ORB_OK) raise("orb_host_alloc");
    pinBytes_[which] = bytes + bytes / 4;
}

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*mask*/, std::vector<cv::KeyPoint>& _keypoints,
                              cv::OutputArray _descriptors) {
    if (_image.empty()) return;
    cv::Mat image = _image.getMat();
    assert(image.type() == CV_8UC1);

    int bound = 0;
    if (orb_extractor_keypoint_bound(handle_, image.rows, image.cols, &bound) != ORB_OK) raise("unsupported image shape");
    static_assert(sizeof(cv::KeyPoint) == sizeof(orb_keypoint), "cv::KeyPoint must be the 28-byte POD the C ABI writes");
    const size_t rowBytes = (size_t)image.cols;
    pinned(0, rowBytes * image.rows);
    pinned(1, (size_t)bound * sizeof(orb_keypoint));
    pinned(2, (size_t)bound * 32);
    uint8_t* pimg = static_cast<uint8_t*>(pin_[0]);
    for (int y = 0; y < image.rows; ++y) memcpy(pimg + (size_t)y * rowBytes, image.ptr<uint8_t>(y), rowBytes);
    orb_keypoint* pk = static_cast<orb_keypoint*>(pin_[1]);
    uint8_t* pd = static_cast<uint8_t*>(pin_[2]);
    int count = 0;
    if (orb_extract(handle_, pimg, image.rows, image.cols, rowBytes, pk, pd, bound, &count) != ORB_OK) raise("orb_extract");

    // mvImagePyramid: level ROI inside a (w+38) x (h+38) buffer with a reflected border (src/ORBextractor.cc:497-515);
    // all levels land in page-locked staging with one set of copies and one synchronisation
    if (keep_pyramid()) {
        std::vector<cv::Mat> padded(nlevels);
        std::vector<uint8_t*> dst(nlevels);
        std::vector<size_t> stride(nlevels);
        std::vector<int> lr(nlevels), lc(nlevels);
        size_t total = 0;
        for (int level = 0; level < nlevels; ++level) {
            if (orb_get_pyramid_level(handle_, 0, level, nullptr, 0, &lr[level], &lc[level]) != ORB_OK) raise("orb_get_pyramid_level");  // sizes only
            total += (size_t)lr[level] * lc[level];
        }
        pinned(3, total);
        size_t off = 0;
        for (int level = 0; level < nlevels; ++level) {
            dst[level] = static_cast<uint8_t*>(pin_[3]) + off;
            stride[level] = (size_t)lc[level];
            off += (size_t)lr[level] * lc[level];
        }
        if (orb_get_pyramid_levels(handle_, 0, dst.data(), stride.data()) != ORB_OK) raise("orb_get_pyramid_levels");
        for (int level = 0; level < nlevels; ++level) {
            const int r = lr[level], c = lc[level];
            padded[level] = cv::Mat(cv::Size(c + 2 * EDGE_THRESHOLD, r + 2 * EDGE_THRESHOLD), image.type());
            mvImagePyramid[level] = padded[level](cv::Rect(EDGE_THRESHOLD, EDGE_THRESHOLD, c, r));
            for (int y = 0; y < r; ++y) memcpy(mvImagePyramid[level].ptr<uint8_t>(y), dst[level] + (size_t)y * c, (size_t)c);
            cv::copyMakeBorder(mvImagePyramid[level], padded[level], EDGE_THRESHOLD, EDGE_THRESHOLD, EDGE_THRESHOLD, EDGE_THRESHOLD,
                               cv::BORDER_REFLECT_101 + cv::BORDER_ISOLATED);
        }
    }

    if (count == 0) {
        _descriptors.release();
        return;
    }
    _descriptors.create(count, 32, CV_8U);
    cv::Mat out = _descriptors.getMat();
    for (int i = 0; i < count; ++i) memcpy(out.ptr<uint8_t>(i), pd + 32 * (size_t)i, 32);
    const cv::KeyPoint* k = reinterpret_cast<const cv::KeyPoint*>(pk);
    _keypoints.assign(k, k + count);
}

}  // namespace ORB_SLAM2
