// Test infrastructure, never shipped.  Force-included ahead of the reference's src/Frame.cc (oracle/Makefile refframe): keeps
// the reference's own include/Frame.h and include/ORBmatcher.h, and replaces -- through their include guards -- MapPoint.h,
// KeyFrame.h (stand-ins of orbslam_objects.h), ORBextractor.h, ORBVocabulary.h and Converter.h by the stubs below.  The stub
// extractor hands out keypoints, descriptors and pyramid levels that the bridge loaded into it (they come from the compiled
// reference extractor or from the oracle); everything Frame does with them is the reference's code.
#pragma once
#define MSHIM_REAL_FRAME
#define ORBEXTRACTOR_H
#define ORBVOCABULARY_H
#define CONVERTER_H
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "orbslam_objects.h"

struct orb_extractor;  // include/orb_b200.h (adapter arm only)

namespace ORB_SLAM2 {
class ORBextractor {
public:
    // adapter arm (oracle/Makefile adapterframe): the C-ABI extractor whose last call left this image's pyramid on the
    // device, what adapter/ORBextractor.h's handle() returns in a real build
    orb_extractor* gpu = nullptr;
    orb_extractor* handle() const { return gpu; }
    std::vector<cv::KeyPoint> keys;
    cv::Mat desc;
    std::vector<float> scale, invScale, sigma2, invSigma2;
    float scaleFactor = 1.2f;
    std::vector<cv::Mat> mvImagePyramid;

    void operator()(const cv::Mat&, const cv::Mat&, std::vector<cv::KeyPoint>& k, cv::Mat& d) {
        k = keys;
        d = desc.clone();
    }
    int GetLevels() { return (int)scale.size(); }
    float GetScaleFactor() { return scaleFactor; }
    std::vector<float> GetScaleFactors() { return scale; }
    std::vector<float> GetInverseScaleFactors() { return invScale; }
    std::vector<float> GetScaleSigmaSquares() { return sigma2; }
    std::vector<float> GetInverseScaleSigmaSquares() { return invSigma2; }
};
class ORBVocabulary {
public:
    void transform(const std::vector<cv::Mat>&, DBoW2::BowVector&, DBoW2::FeatureVector&, int) {}
};
class Converter {
public:
    static std::vector<cv::Mat> toDescriptorVector(const cv::Mat&) { return std::vector<cv::Mat>(); }
};
}  // namespace ORB_SLAM2
