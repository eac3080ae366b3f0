// Drop-in replacement for the reference's include/ORBextractor.h: same namespace, class name,
// constructor, operator(), getters and public mvImagePyramid (reference include/ORBextractor.h:26-93),
// so Frame.cc / Tracking.cc compile and behave unchanged.  The work is done by liborb_b200.so
// (include/orb_b200.h); this file only adapts cv:: containers to the C ABI.
//
// Build where OpenCV exists (the reference's own environment): put this directory before the
// reference's include/ in the include path, drop src/ORBextractor.cc from the source list and
// link liborb_b200.so.  In this repository it is syntax-checked against
// oracle/cvshim because the image has no OpenCV C++ headers.
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <list>
#include <vector>

#include <opencv/cv.h>

struct orb_extractor;  // include/orb_b200.h

namespace ORB_SLAM2 {

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    ~ORBextractor();
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // Compute the ORB features and descriptors on an image.  Mask is ignored, as in the reference.
    void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints,
                    cv::OutputArray descriptors);

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return scaleFactor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    // The C-ABI handle behind this instance: what orb_b200::Matcher::ComputeStereoMatches needs to read the pyramids of
    // the last call where they live on the device (an addition; nothing in the reference calls it).
    orb_extractor* handle() const { return handle_; }

    // Two switches the reference does not have (process-wide, read when an instance is constructed / called):
    // the CUDA device of the instances constructed afterwards (default: ORB_B200_DEVICE in the environment, else 0), and
    // whether operator() copies the pyramid back into mvImagePyramid (default on, or ORB_B200_IMAGE_PYRAMID=0).  With
    // Frame::ComputeStereoMatches routed through orb_compute_stereo_matches the pyramid is read on the device and the
    // copy (1.9 MB per KITTI frame) can be switched off.
    static void UseDevice(int device);
    static void KeepImagePyramid(bool on);

    // Refilled after every call (tight level images inside a 19-px reflected border, exactly the
    // layout Frame::ComputeStereoMatches reads, reference src/Frame.cc:453,543-560).
    std::vector<cv::Mat> mvImagePyramid;

protected:
    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;

    std::vector<int> mnFeaturesPerLevel;
    std::vector<float> mvScaleFactor;
    std::vector<float> mvInvScaleFactor;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;

private:
    // page-locked staging of the frame, the results and the pyramid levels (grown on demand): the caller's cv::Mat /
    // std::vector are pageable, and copies from / to pageable memory are several times slower and not asynchronous
    void pinned(size_t which, size_t bytes);
    orb_extractor* handle_;
    void* pin_[4] = {nullptr, nullptr, nullptr, nullptr};  // frame, keypoints, descriptors, pyramid
    size_t pinBytes_[4] = {0, 0, 0, 0};
};

}  // namespace ORB_SLAM2

#endif
