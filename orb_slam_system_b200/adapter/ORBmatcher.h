// Drop-in ORB_SLAM2::ORBmatcher over liborb_b200 (B200 / sm_100a).
//
// Same class name, constructor, static DescriptorDistance and the eleven Search* / Fuse signatures as the reference's
// include/ORBmatcher.h:17-83, so Tracking.cc, LocalMapping.cc, LoopClosing.cc, Frame.cc and MapPoint.cc compile against it
// unchanged.  The bodies (ORBmatcher_b200.cc) keep what needs the object graph -- the per-point gates, the pose arithmetic
// that projects a map point, the write-back into Frame / KeyFrame / MapPoint -- and hand every candidate window and
// every DescriptorDistance loop to the GPU through include/orb_b200.h (orb_search_*).
//
// Replace include/ORBmatcher.h and src/ORBmatcher.cc of the reference by this header and ORBmatcher_b200.cc and link
// liborb_b200.so.  One GPU matcher handle is created per calling thread on first use (the reference
// calls these methods from its Tracking, LocalMapping and LoopClosing threads); the CUDA device is ORB_B200_DEVICE
// (environment, default 0) or whatever UseDevice() was last given.
#ifndef ORBMATCHER_H
#define ORBMATCHER_H

#include <set>
#include <utility>
#include <vector>

#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>

#include "MapPoint.h"
#include "KeyFrame.h"
#include "Frame.h"

namespace ORB_SLAM2 {

class ORBmatcher {
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true);

    // src/ORBmatcher.cc:896-908.  One pair through the GPU scan; loops over many pairs (MapPoint.cc:252) should use
    // orb_match_all on the whole set instead.
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);

    // src/ORBmatcher.cc:19-65 (track local map)
    int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3);
    // :732-818 (track with motion model)
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono);
    // :820-894 (relocalisation)
    int SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist);
    // :121-195 (loop closing)
    int SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, std::vector<MapPoint*>& vpMatched, int th);

    // :88-119: in this fork the method walks the two feature vectors without comparing anything -- kept so
    int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches);
    // :278-366
    int SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12);

    // :197-276
    int SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize = 10);

    // :368-467
    int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, std::vector<std::pair<size_t, size_t> >& vMatchedPairs,
                               const bool bOnlyStereo);

    // :636-730
    int SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12,
                     const cv::Mat& t12, const float th);

    // :504-568 and :570-634
    int Fuse(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const float th = 3.0);
    int Fuse(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, float th, std::vector<MapPoint*>& vpReplacePoint);

    // CUDA device of the matcher handles created after this call (not part of the reference's interface).
    static void UseDevice(int device);

public:
    static const int TH_LOW;
    static const int TH_HIGH;
    static const int HISTO_LENGTH;

protected:
    bool CheckDistEpipolarLine(const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const cv::Mat& F12, const KeyFrame* pKF);
    float RadiusByViewingCos(const float& viewCos);
    void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);

    float mfNNratio;
    bool mbCheckOrientation;
};

}  // namespace ORB_SLAM2

#endif  // ORBMATCHER_H
