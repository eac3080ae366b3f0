// C entry point over the reference's own ORB_SLAM2::KeyFrame, compiled UNMODIFIED from /root/reference/src/KeyFrame.cc (with
// src/Frame.cc behind it) against oracle/mshim/keyframe_objects.h -- TEST INFRASTRUCTURE ONLY.
//   refm_keyframe_features_in_area: a Frame from the reference's monocular constructor (AssignFeaturesToGrid), the KeyFrame the
//   reference builds from it (src/KeyFrame.cc:11-37 copies the grid), then KeyFrame::GetFeaturesInArea (:549-588) per query.
#include "KeyFrame.h"

#include <cstdio>
#include <cstring>
#include <memory>

using namespace ORB_SLAM2;
typedef orb_oracle::KeyPoint OKP;

extern "C" int refm_keyframe_features_in_area(const OKP* keys, int n, int rows, int cols, int nq, const float* x, const float* y,
                                              const float* r, int* offsets, int* cand, int cap) {
  try {
    ORBextractor ex;
    const float scale[8] = {1.f, 1.f, 1.2f, 1.44f, 1.728f, 2.0736f, 2.48832f, 2.985984f};
    for (int i = 0; i < n; ++i) {
        cv::KeyPoint c;
        c.pt = cv::Point2f(keys[i].x, keys[i].y);
        c.size = keys[i].size;
        c.angle = keys[i].angle;
        c.response = keys[i].response;
        c.octave = keys[i].octave;
        c.class_id = keys[i].class_id;
        ex.keys.push_back(c);
    }
    ex.desc = cv::Mat(std::max(n, 1), 32, CV_8U);
    for (int l = 0; l < 8; ++l) {
        ex.scale.push_back(scale[l]);
        ex.invScale.push_back(1.0f / scale[l]);
        ex.sigma2.push_back(scale[l] * scale[l]);
        ex.invSigma2.push_back(1.0f / (scale[l] * scale[l]));
    }
    cv::Mat img(rows, cols, CV_8U);
    cv::Mat K = cv::Mat::eye(3, 3, CV_32F), dist = cv::Mat::zeros(4, 1, CV_32F);
    ORBVocabulary voc;
    Frame::mbInitialComputations = true;
    Frame F(img, 0.0, &ex, &voc, K, dist, 1.0f, 40.0f);
    F.SetPose(cv::Mat::eye(4, 4, CV_32F));  // keyframes are made from tracked frames: the constructor reads the pose
    KeyFrame kf(F, nullptr, nullptr);
    int total = 0;
    offsets[0] = 0;
    for (int i = 0; i < nq; ++i) {
        const std::vector<size_t> v = kf.GetFeaturesInArea(x[i], y[i], r[i]);
        for (size_t idx : v) {
            if (total < cap) cand[total] = (int)idx;
            ++total;
        }
        offsets[i + 1] = total;
    }
    return total;
  } catch (const std::exception& e) {
    fprintf(stderr, "refm_keyframe_features_in_area: %s\n", e.what());
    return -1;
  }
}
