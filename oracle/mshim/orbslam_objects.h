// Test infrastructure, never shipped.  Stand-ins for ORB_SLAM2::MapPoint / KeyFrame / Frame with exactly the members the
// reference's src/ORBmatcher.cc touches.  This header is force-included (-include) ahead of the reference's own
// include/ORBmatcher.h and DEFINES THE INCLUDE GUARDS of the reference's MapPoint.h / KeyFrame.h / Frame.h, so those three
// headers (which pull in OpenCV, DBoW2's vocabulary, g2o ...) are skipped by their own guards while ORBmatcher.h and
// ORBmatcher.cc are compiled unmodified from where they lie.  The objects hold plain arrays and canned answers; nothing
// here decides a match -- every candidate loop, threshold, ratio test, histogram and greedy update that runs is the
// reference's.  Frame::GetFeaturesInArea / KeyFrame::GetFeaturesInArea (src/Frame.cc, src/KeyFrame.cc are not compiled)
// come from the oracle's restatement, which tests/test_search_oracle.py pins separately.
#pragma once
#define MAPPOINT_H
#ifndef MSHIM_REAL_KEYFRAME  // keyframe_objects.h keeps the reference's own include/KeyFrame.h
#define KEYFRAME_H
#endif
#ifndef MSHIM_REAL_FRAME  // frame_objects.h keeps the reference's own include/Frame.h
#define FRAME_H
#endif

#include <map>
#include <set>
#include <vector>

#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"
#include "mshim_cv.hpp"
#include "orb_oracle.h"

struct OrcFrame {  // oracle/orb_oracle.cpp
    const orb_oracle::KeyPoint* keysUn;
    const uint8_t* desc;
    int N;
    float mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv;
};
extern "C" int orc_features_in_area(const OrcFrame* F, int nq, const float* x, const float* y, const float* r, const int* minLevel,
                                    const int* maxLevel, int* offsets, int* cand, int cap);

namespace ORB_SLAM2 {
class KeyFrame;
class Frame;

class MapPoint {
public:
    int id = -1;  // bridge bookkeeping
    bool bad = false;
    int nObs = 1;
    int predictedLevel = 0;
    cv::Mat desc, pos, normal;
    std::map<const KeyFrame*, int> indexIn;
    std::vector<std::pair<KeyFrame*, size_t>> addedObservations;
    MapPoint* replacedBy = nullptr;
    // the fields Frame::isInFrustum leaves behind (include/MapPoint.h)
    bool mbTrackInView = false;
    float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0, mTrackViewCos = 0;
    int mnTrackScaleLevel = 0;

    bool isBad() { return bad; }
    int Observations() { return nObs; }
    cv::Mat GetDescriptor() { return desc.clone(); }
    cv::Mat GetWorldPos() { return pos.clone(); }
    cv::Mat GetNormal() { return normal.clone(); }
    float GetMinDistanceInvariance() { return 0.0f; }
    float GetMaxDistanceInvariance() { return 1e30f; }
    int PredictScale(const float&, KeyFrame*) { return predictedLevel; }
    int PredictScale(const float&, Frame*) { return predictedLevel; }
    int GetIndexInKeyFrame(KeyFrame* kf) {
        auto it = indexIn.find(kf);
        return it == indexIn.end() ? -1 : it->second;
    }
    bool IsInKeyFrame(KeyFrame* kf) { return indexIn.count(kf) != 0; }
    std::map<KeyFrame*, size_t> GetObservations() { return std::map<KeyFrame*, size_t>(); }
    void EraseObservation(KeyFrame*) {}
    void AddObservation(KeyFrame* kf, size_t idx) { addedObservations.push_back({kf, idx}); }
    void Replace(MapPoint* p) { replacedBy = p; }
};

// A std::vector<MapPoint*> whose element assignments are logged, so that the bridge can report which query took which
// feature even when a later query (or the rotation check) overwrites the slot.
class LoggedPoints {
public:
    std::vector<MapPoint*> v;
    std::vector<std::pair<size_t, MapPoint*>> log;
    struct Ref {
        LoggedPoints* o;
        size_t i;
        operator MapPoint*() const { return o->v[i]; }
        MapPoint* operator->() const { return o->v[i]; }
        Ref& operator=(MapPoint* p) {
            o->v[i] = p;
            o->log.push_back({i, p});
            return *this;
        }
        Ref& operator=(const Ref& r) { return *this = (MapPoint*)r; }
    };
    Ref operator[](size_t i) { return Ref{this, i}; }
    MapPoint* operator[](size_t i) const { return v[i]; }
    size_t size() const { return v.size(); }
};

struct GridHolder {  // what GetFeaturesInArea needs, in the oracle's layout
    std::vector<orb_oracle::KeyPoint> keys;
    std::vector<uint8_t> desc;
    float mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0, wInv = 0, hInv = 0;
    std::vector<size_t> area(float x, float y, float r, int minLevel, int maxLevel) const {
        OrcFrame F{keys.data(), desc.data(), (int)keys.size(), mnMinX, mnMinY, wInv, hInv};
        int off[2] = {0, 0};
        std::vector<int> cand(keys.size() + 1);
        const int n = orc_features_in_area(&F, 1, &x, &y, &r, &minLevel, &maxLevel, off, cand.data(), (int)cand.size());
        return std::vector<size_t>(cand.begin(), cand.begin() + n);
    }
};

#ifndef MSHIM_REAL_KEYFRAME
class KeyFrame {
public:
    GridHolder grid;
    int N = 0;
    float fx = 1, fy = 1, cx = 0, cy = 0;
    // include/KeyFrame.h:112-113, :165-168 under the reference's own names (read by adapter/ORBmatcher_b200.cc); the bridge
    // fills them from `grid` with grid_members()
    float mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0, mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    void grid_members() {
        mnMinX = grid.mnMinX;
        mnMinY = grid.mnMinY;
        mnMaxX = grid.mnMaxX;
        mnMaxY = grid.mnMaxY;
        mfGridElementWidthInv = grid.wInv;
        mfGridElementHeightInv = grid.hInv;
    }
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight, mvScaleFactors, mvLevelSigma2;
    cv::Mat mDescriptors;
    DBoW2::FeatureVector mFeatVec;
    std::vector<MapPoint*> mapPoints;
    std::vector<std::pair<MapPoint*, size_t>> addedMapPoints;
    cv::Mat Rcw = cv::Mat::eye(3, 3, CV_32F), tcw = cv::Mat::zeros(3, 1, CV_32F), Ow = cv::Mat::zeros(3, 1, CV_32F);

    std::vector<MapPoint*> GetMapPointMatches() { return mapPoints; }
    MapPoint* GetMapPoint(const size_t& idx) { return idx < mapPoints.size() ? mapPoints[idx] : nullptr; }
    void AddMapPoint(MapPoint* p, const size_t& idx) { addedMapPoints.push_back({p, idx}); }
    cv::Mat GetRotation() { return Rcw.clone(); }
    cv::Mat GetTranslation() { return tcw.clone(); }
    cv::Mat GetCameraCenter() { return Ow.clone(); }
    bool IsInImage(const float& x, const float& y) const { return x >= grid.mnMinX && x < grid.mnMaxX && y >= grid.mnMinY && y < grid.mnMaxY; }
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r) const { return grid.area(x, y, r, -1, -1); }
};
#endif

#ifndef MSHIM_REAL_FRAME
class Frame {
public:
    GridHolder grid;
    int N = 0;
    float fx = 1, fy = 1, cx = 0, cy = 0, mb = 1;
    float mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0;
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;  // include/Frame.h:139-140 (static there)
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight, mvScaleFactors;
    std::vector<bool> mvbOutlier;
    cv::Mat mDescriptors;
    cv::Mat mTcw = cv::Mat::eye(4, 4, CV_32F);
    DBoW2::FeatureVector mFeatVec;
    LoggedPoints mvpMapPoints;

    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1) const {
        return grid.area(x, y, r, minLevel, maxLevel);
    }
};
#endif
}  // namespace ORB_SLAM2

// The reference's headers leak `using namespace std` into ORBmatcher.h (it writes `pair<size_t, size_t>` unqualified).
using namespace std;
