// ORB_SLAM2::ORBmatcher over liborb_b200: the file that takes the place of the reference's src/ORBmatcher.cc.
//
// Every method keeps the part of the reference method that needs the object graph and gives the rest to the GPU:
//   1. gate   walk the map points / keypoints in the reference's order, apply its per-point tests (bad, already found,
//             behind the camera, outside the image, distance range, viewing angle) and project with the same cv::Mat
//             expressions, so the numbers the search starts from come out of the same arithmetic;
//   2. search one orb_search_* call: the device rebuilds the 64 x 48 grid, forms every GetFeaturesInArea window in the
//             reference's order and computes every DescriptorDistance; the library replays the method's accept rule and
//             greedy "feature already taken" state in query order (csrc/search_capi.cu);
//   3. write  Frame::mvpMapPoints / vpMatched / vpMatches12 / MapPoint::AddObservation ... exactly where the reference
//             writes them.
// file:line citations are the reference's src/ORBmatcher.cc.
#include "ORBmatcher.h"

#include <climits>
#include <cmath>
#include <cstdlib>

#include "orb_match_b200.h"

namespace ORB_SLAM2 {

const int ORBmatcher::TH_HIGH = 100;
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;

namespace {
int g_device = -1;

// One GPU matcher handle (stream + scratch) per calling thread: ORBmatcher objects are stack-local in the reference and
// its Tracking, LocalMapping and LoopClosing threads search concurrently.
orb_b200::Matcher& gpu() {
    thread_local orb_b200::Matcher m(g_device >= 0 ? g_device : (getenv("ORB_B200_DEVICE") ? atoi(getenv("ORB_B200_DEVICE")) : 0));
    return m;
}

typedef orb_b200::Matcher::Queries Queries;

void push_descriptor(Queries& Q, const cv::Mat& d) { Q.desc.insert(Q.desc.end(), d.ptr<uint8_t>(0), d.ptr<uint8_t>(0) + 32); }

// What GetFeaturesInArea reads (src/Frame.cc:307-360, src/KeyFrame.cc:549-588); the grid itself is rebuilt on the device.
orb_frame_view view_of(const Frame& F) {
    return orb_b200::Matcher::View(F.mvKeysUn, F.mDescriptors, F.mnMinX, F.mnMinY, F.mfGridElementWidthInv, F.mfGridElementHeightInv);
}
orb_frame_view view_of(const KeyFrame& K) {
    return orb_b200::Matcher::View(K.mvKeysUn, K.mDescriptors, (float)K.mnMinX, (float)K.mnMinY, K.mfGridElementWidthInv,
                                   K.mfGridElementHeightInv);
}

// Pinhole projection of a camera-frame point the way the best-only searches write it: u = fx * x * (1 / z) + cx.
struct Pixel {
    float u, v;
};
Pixel project_invz(const cv::Mat& Xc, float fx, float fy, float cx, float cy) {
    const float invz = 1.0f / Xc.at<float>(2);
    Pixel p;
    p.u = fx * Xc.at<float>(0) * invz + cx;
    p.v = fy * Xc.at<float>(1) * invz + cy;
    return p;
}
}  // namespace

void ORBmatcher::UseDevice(int device) { g_device = device; }

ORBmatcher::ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) { return gpu().DescriptorDistance(a, b); }

float ORBmatcher::RadiusByViewingCos(const float& viewCos) { return viewCos > 0.998 ? 2.5f : 4.0f; }  // :67-69

// :71-85.  Not called by the methods below (the epipolar test of SearchForTriangulation runs inside the library); kept
// because the reference's header declares it.
bool ORBmatcher::CheckDistEpipolarLine(const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const cv::Mat& F12, const KeyFrame* pKF2) {
    float l[3];
    for (int j = 0; j < 3; ++j) l[j] = kp1.pt.x * F12.at<float>(0, j) + kp1.pt.y * F12.at<float>(1, j) + F12.at<float>(2, j);
    const float num = l[0] * kp2.pt.x + l[1] * kp2.pt.y + l[2];
    const float den = l[0] * l[0] + l[1] * l[1];
    if (den == 0) return false;
    return num * num / den < 3.84 * pKF2->mvLevelSigma2[kp2.octave];
}

// :469-502.  The three fullest histogram bins, the second / third dropped when below a tenth of the first.
void ORBmatcher::ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3) {
    int idx[3] = {-1, -1, -1}, val[3] = {0, 0, 0};
    for (int i = 0; i < L; ++i) {
        const int s = (int)histo[i].size();
        int j = 0;
        while (j < 3 && s <= val[j]) ++j;
        if (j == 3) continue;
        for (int k = 2; k > j; --k) {
            val[k] = val[k - 1];
            idx[k] = idx[k - 1];
        }
        val[j] = s;
        idx[j] = i;
    }
    ind1 = idx[0];
    ind2 = val[1] < 0.1f * val[0] ? -1 : idx[1];
    ind3 = (val[1] < 0.1f * val[0] || val[2] < 0.1f * val[0]) ? -1 : idx[2];
}

// ---- :19-65 ------------------------------------------------------------------------------------------------------
int ORBmatcher::SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th) {
    Queries Q;
    std::vector<MapPoint*> who;
    for (MapPoint* pMP : vpMapPoints) {
        if (!pMP || pMP->isBad() || !pMP->mbTrackInView) continue;  // :24
        push_descriptor(Q, pMP->GetDescriptor());
        Q.u.push_back(pMP->mTrackProjX);
        Q.v.push_back(pMP->mTrackProjY);
        Q.uR.push_back(pMP->mTrackProjXR);
        Q.level.push_back(pMP->mnTrackScaleLevel);
        Q.viewCos.push_back(pMP->mTrackViewCos);
        Q.observed.push_back(pMP->Observations() > 0 ? 1 : 0);
        who.push_back(pMP);
    }
    if (who.empty()) return 0;
    std::vector<uint8_t> occupied(F.N);
    for (int i = 0; i < F.N; ++i) {  // :38
        MapPoint* p = F.mvpMapPoints[i];
        occupied[i] = p && p->Observations() > 0;
    }
    std::vector<int32_t> hit;
    const int nmatches = gpu().SearchByProjection(view_of(F), F.mvuRight, occupied, F.mvScaleFactors, Q, th, mfNNratio, hit);
    for (size_t q = 0; q < who.size(); ++q)
        if (hit[q] >= 0) F.mvpMapPoints[hit[q]] = who[q];  // :59
    return nmatches;
}

// ---- :88-119: a walk over two feature vectors that compares nothing --------------------------------
int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches) {
    (void)pKF;
    vpMapPointMatches.resize(F.N, nullptr);
    return 0;
}

// ---- :121-195 ----------------------------------------------------------------------------------------------------
int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, std::vector<MapPoint*>& vpMatched, int th) {
    // similarity -> rotation, translation, camera centre (:126-130)
    cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
    const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
    cv::Mat Rcw = sRcw / scw;
    cv::Mat tcw = Scw.rowRange(0, 3).col(3) / scw;
    cv::Mat Ow = -Rcw.t() * tcw;

    std::set<MapPoint*> found(vpMatched.begin(), vpMatched.end());
    found.erase(nullptr);

    Queries Q;
    std::vector<MapPoint*> who;
    for (MapPoint* pMP : vpPoints) {
        if (pMP->isBad() || found.count(pMP)) continue;
        cv::Mat Xw = pMP->GetWorldPos();
        cv::Mat Xc = Rcw * Xw + tcw;
        if (Xc.at<float>(2) < 0.0f) continue;
        const Pixel px = project_invz(Xc, pKF->fx, pKF->fy, pKF->cx, pKF->cy);
        if (!pKF->IsInImage(px.u, px.v)) continue;
        const float maxDistance = pMP->GetMaxDistanceInvariance(), minDistance = pMP->GetMinDistanceInvariance();
        cv::Mat PO = Xw - Ow;
        const float dist = cv::norm(PO);
        if (dist < minDistance || dist > maxDistance) continue;
        cv::Mat Pn = pMP->GetNormal();
        if (PO.dot(Pn) < 0.5 * dist) continue;  // viewing angle below 60 degrees
        const int level = pMP->PredictScale(dist, pKF);
        push_descriptor(Q, pMP->GetDescriptor());
        Q.u.push_back(px.u);
        Q.v.push_back(px.v);
        Q.radius.push_back(th * pKF->mvScaleFactors[level]);
        who.push_back(pMP);
    }
    if (who.empty()) return 0;
    std::vector<uint8_t> claimed(vpMatched.size());
    for (size_t i = 0; i < vpMatched.size(); ++i) claimed[i] = vpMatched[i] != nullptr;  // :177
    std::vector<int32_t> hit;
    const int nmatches = gpu().SearchBest(view_of(*pKF), claimed, Q, false, ORB_ROT_NONE, TH_LOW, hit);
    for (size_t q = 0; q < who.size(); ++q)
        if (hit[q] >= 0) vpMatched[hit[q]] = who[q];  // :189
    return nmatches;
}

// ---- :197-276 ----------------------------------------------------------------------------------------------------
int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize) {
    return gpu().SearchForInitialization(F1.mvKeysUn, F1.mDescriptors, view_of(F2), vbPrevMatched, vnMatches12, windowSize, mfNNratio,
                                         mbCheckOrientation);
}

// ---- :278-366 ----------------------------------------------------------------------------------------------------
namespace {
void usable_points(const std::vector<MapPoint*>& pts, std::vector<uint8_t>& has) {
    has.resize(pts.size());
    for (size_t i = 0; i < pts.size(); ++i) has[i] = pts[i] && !pts[i]->isBad();  // :309, :316
}
void angles_of(const std::vector<cv::KeyPoint>& keys, std::vector<float>& a) {
    a.resize(keys.size());
    for (size_t i = 0; i < keys.size(); ++i) a[i] = keys[i].angle;
}
}  // namespace

int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12) {
    const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
    vpMatches12.resize(vpMapPoints1.size(), nullptr);  // :289 (a caller's earlier entries stay, as in the reference)
    std::vector<uint8_t> has1, has2;
    usable_points(vpMapPoints1, has1);
    usable_points(vpMapPoints2, has2);
    std::vector<float> ang1, ang2;
    angles_of(pKF1->mvKeysUn, ang1);
    angles_of(pKF2->mvKeysUn, ang2);
    const orb_b200::Matcher::FlatFeatureVector fv1(pKF1->mFeatVec), fv2(pKF2->mFeatVec);
    std::vector<int32_t> m12;
    const int nmatches = gpu().SearchByBoW(pKF1->mDescriptors, ang1, has1, fv1, pKF2->mDescriptors, ang2, has2, fv2, mfNNratio,
                                           mbCheckOrientation, m12);
    for (size_t i1 = 0; i1 < m12.size() && i1 < vpMatches12.size(); ++i1)
        if (m12[i1] >= 0) vpMatches12[i1] = vpMapPoints2[m12[i1]];  // :330
    return nmatches;
}

// ---- :368-467 ----------------------------------------------------------------------------------------------------
int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, std::vector<std::pair<size_t, size_t> >& vMatchedPairs,
                                       const bool bOnlyStereo) {
    if (bOnlyStereo) {  // :417 accepts a candidate only when !bOnlyStereo: with the flag set nothing can match
        vMatchedPairs.clear();
        return 0;
    }
    std::vector<uint8_t> has1(pKF1->N), has2(pKF2->N);
    for (int i = 0; i < pKF1->N; ++i) has1[i] = pKF1->GetMapPoint(i) != nullptr;  // :399-400
    for (int i = 0; i < pKF2->N; ++i) has2[i] = pKF2->GetMapPoint(i) != nullptr;  // :409
    float f12[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) f12[3 * r + c] = F12.at<float>(r, c);
    const orb_b200::Matcher::FlatFeatureVector fv1(pKF1->mFeatVec), fv2(pKF2->mFeatVec);
    return gpu().SearchForTriangulation(pKF1->mvKeysUn, pKF1->mDescriptors, has1, fv1, pKF2->mvKeysUn, pKF2->mDescriptors, has2, fv2, f12,
                                        pKF2->mvLevelSigma2, mbCheckOrientation, vMatchedPairs);
}

// ---- :504-568 and :570-634 ----------------------------------------------------------------------------------------
namespace {
// The projection + window + scan part the two Fuse overloads share (:514-551, :582-619): per surviving point the feature
// of pKF with the smallest distance among octaves level - 1 .. level, or -1.
struct FuseSearch {
    std::vector<int> which;  // index into the caller's point vector
    std::vector<int32_t> hit;
};
FuseSearch fuse_search(KeyFrame* pKF, const cv::Mat& Rcw, const cv::Mat& tcw, const cv::Mat& Ow, const std::vector<MapPoint*>& pts,
                       const std::vector<uint8_t>& skip, float th, int maxDist) {
    FuseSearch S;
    Queries Q;
    for (size_t i = 0; i < pts.size(); ++i) {
        MapPoint* pMP = pts[i];
        if (skip[i]) continue;
        cv::Mat Xw = pMP->GetWorldPos();
        cv::Mat Xc = Rcw * Xw + tcw;
        if (Xc.at<float>(2) < 0.0f) continue;
        const float invz = 1.0 / Xc.at<float>(2);
        const float x = Xc.at<float>(0) * invz, y = Xc.at<float>(1) * invz;
        const float u = pKF->fx * x + pKF->cx, v = pKF->fy * y + pKF->cy;
        if (!pKF->IsInImage(u, v)) continue;
        const float dist = cv::norm(Xw - Ow);
        if (dist < pMP->GetMinDistanceInvariance() || dist > pMP->GetMaxDistanceInvariance()) continue;
        const int level = pMP->PredictScale(dist, pKF);
        push_descriptor(Q, pMP->GetDescriptor());
        Q.u.push_back(u);
        Q.v.push_back(v);
        Q.radius.push_back(th * pKF->mvScaleFactors[level]);
        Q.minLevel.push_back(level - 1);  // :542, :610
        Q.maxLevel.push_back(level);
        S.which.push_back((int)i);
    }
    if (!S.which.empty()) {
        std::vector<uint8_t> none;
        gpu().SearchBest(view_of(*pKF), none, Q, true, ORB_ROT_NONE, maxDist, S.hit);
    }
    return S;
}
}  // namespace

int ORBmatcher::Fuse(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const float th) {
    cv::Mat Rcw = pKF->GetRotation(), tcw = pKF->GetTranslation(), Ow = pKF->GetCameraCenter();
    std::vector<uint8_t> skip(vpMapPoints.size());
    for (size_t i = 0; i < vpMapPoints.size(); ++i) {
        MapPoint* p = vpMapPoints[i];
        skip[i] = !p || p->isBad() || p->IsInKeyFrame(pKF);  // :512
    }
    const FuseSearch S = fuse_search(pKF, Rcw, tcw, Ow, vpMapPoints, skip, th, TH_LOW);
    int nFused = 0;
    for (size_t q = 0; q < S.which.size(); ++q) {  // the bookkeeping of :553-565, in the reference's order
        MapPoint* pMP = vpMapPoints[S.which[q]];
        if (pMP->IsInKeyFrame(pKF)) continue;  // an earlier entry of the same point was added in this very call (:512)
        if (S.hit[q] < 0) continue;
        MapPoint* pMPinKF = pKF->GetMapPoint(S.hit[q]);
        if (pMPinKF) {
            if (!pMPinKF->isBad() && pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
        } else {
            pMP->AddObservation(pKF, S.hit[q]);
            pKF->AddMapPoint(pMP, S.hit[q]);
        }
        ++nFused;
    }
    return nFused;
}

int ORBmatcher::Fuse(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, float th, std::vector<MapPoint*>& vpReplacePoint) {
    cv::Mat Rcw = Scw.rowRange(0, 3).colRange(0, 3);
    cv::Mat tcw = Scw.rowRange(0, 3).col(3);
    cv::Mat Ow = -Rcw.t() * tcw;
    std::vector<uint8_t> skip(vpPoints.size());
    for (size_t i = 0; i < vpPoints.size(); ++i) {
        MapPoint* p = vpPoints[i];
        skip[i] = !p || p->isBad() || pKF->GetMapPoint(p->GetIndexInKeyFrame(pKF));  // :580
    }
    const FuseSearch S = fuse_search(pKF, Rcw, tcw, Ow, vpPoints, skip, th, TH_LOW);
    int nFused = 0;
    for (size_t q = 0; q < S.which.size(); ++q) {  // :621-630
        const int iMP = S.which[q];
        MapPoint* pMP = vpPoints[iMP];
        if (pKF->GetMapPoint(pMP->GetIndexInKeyFrame(pKF))) continue;  // became true through an earlier entry of this call (:580)
        if (S.hit[q] < 0) continue;
        MapPoint* pMPinKF = pKF->GetMapPoint(S.hit[q]);
        if (pMPinKF) {
            if (!pMPinKF->isBad()) vpReplacePoint[iMP] = pMPinKF;
        } else {
            pMP->AddObservation(pKF, S.hit[q]);
            pKF->AddMapPoint(pMP, S.hit[q]);
        }
        ++nFused;
    }
    return nFused;
}

// ---- :636-730 ----------------------------------------------------------------------------------------------------
int ORBmatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12,
                             const cv::Mat& t12, const float th) {
    const float fx = pKF1->fx, fy = pKF1->fy, cx = pKF1->cx, cy = pKF1->cy;  // :639-642 (KF1's calibration, as the fork has it)
    cv::Mat R1w = pKF1->GetRotation(), t1w = pKF1->GetTranslation();
    cv::Mat sR21 = (1.0 / s12) * R12.t();
    cv::Mat t21 = -sR21 * t12;

    const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches();
    const size_t N1 = vpMapPoints1.size();
    std::vector<bool> already(N1, false);
    for (size_t i = 0; i < vpMatches12.size(); ++i) {  // :661-667 (indexes vbAlreadyMatched1 by the caller's vector, as the reference)
        MapPoint* p = vpMatches12[i];
        if (p && p->GetIndexInKeyFrame(pKF2) >= 0) already[i] = true;
    }
    Queries Q;
    std::vector<int> which;
    for (size_t i1 = 0; i1 < N1; ++i1) {
        MapPoint* pMP1 = vpMapPoints1[i1];
        if (!pMP1 || already[i1] || pMP1->isBad()) continue;
        cv::Mat Xw = pMP1->GetWorldPos();
        cv::Mat Xc1 = R1w * Xw + t1w;
        cv::Mat Xc2 = sR21 * Xc1 + t21;
        if (Xc2.at<float>(2) < 0.0f) continue;
        const float invz = 1.0f / Xc2.at<float>(2);
        const float x = Xc2.at<float>(0) * invz, y = Xc2.at<float>(1) * invz;
        const float u = fx * x + cx, v = fy * y + cy;
        if (!pKF2->IsInImage(u, v)) continue;
        const float dist = cv::norm(Xc2);
        if (dist < pMP1->GetMinDistanceInvariance() || dist > pMP1->GetMaxDistanceInvariance()) continue;
        const int level = pMP1->PredictScale(dist, pKF2);
        push_descriptor(Q, pMP1->GetDescriptor());
        Q.u.push_back(u);
        Q.v.push_back(v);
        Q.radius.push_back(th * pKF2->mvScaleFactors[level]);
        Q.minLevel.push_back(level - 1);  // :705
        Q.maxLevel.push_back(level);
        which.push_back((int)i1);
    }
    int nMatches = 0;
    std::vector<int32_t> hit;
    if (!which.empty()) {
        std::vector<uint8_t> none;
        nMatches = gpu().SearchBest(view_of(*pKF2), none, Q, true, ORB_ROT_NONE, TH_HIGH, hit);
    }
    vpMatches12 = std::vector<MapPoint*>(N1, nullptr);  // :722-727
    for (size_t q = 0; q < which.size(); ++q)
        if (hit[q] >= 0) vpMatches12[which[q]] = pKF2->GetMapPoint(hit[q]);
    return nMatches;
}

// ---- :732-818 ----------------------------------------------------------------------------------------------------
int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono) {
    const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
    const cv::Mat twc = -Rcw.t() * tcw;
    const cv::Mat Rlw = LastFrame.mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tlw = LastFrame.mTcw.rowRange(0, 3).col(3);
    const cv::Mat tlc = Rlw * twc + tlw;
    const bool bForward = tlc.at<float>(2) > CurrentFrame.mb && !bMono;    // :747
    const bool bBackward = -tlc.at<float>(2) > CurrentFrame.mb && !bMono;  // :748

    Queries Q;
    std::vector<MapPoint*> who;
    for (int i = 0; i < LastFrame.N; ++i) {
        MapPoint* pMP = LastFrame.mvpMapPoints[i];
        if (!pMP || LastFrame.mvbOutlier[i]) continue;
        cv::Mat Xw = pMP->GetWorldPos();
        cv::Mat Xc = Rcw * Xw + tcw;
        const float z = Xc.at<float>(2);
        if (z <= 0) continue;
        const float u = CurrentFrame.fx * Xc.at<float>(0) / z + CurrentFrame.cx;  // :760-761 divide by z (no reciprocal here)
        const float v = CurrentFrame.fy * Xc.at<float>(1) / z + CurrentFrame.cy;
        if (u < CurrentFrame.mnMinX || u > CurrentFrame.mnMaxX || v < CurrentFrame.mnMinY || v > CurrentFrame.mnMaxY) continue;
        const int octave = LastFrame.mvKeys[i].octave;
        push_descriptor(Q, pMP->GetDescriptor());
        Q.u.push_back(u);
        Q.v.push_back(v);
        Q.radius.push_back(th * CurrentFrame.mvScaleFactors[octave]);
        // :769-774: forward motion looks at the same or coarser octaves, backward at the same or finer, else +-1
        Q.minLevel.push_back(bForward ? octave : (bBackward ? 0 : octave - 1));
        Q.maxLevel.push_back(bForward ? -1 : (bBackward ? octave : octave + 1));
        Q.angle.push_back(LastFrame.mvKeys[i].angle);
        who.push_back(pMP);
    }
    if (who.empty()) return 0;
    std::vector<uint8_t> claimed(CurrentFrame.N);
    for (int i = 0; i < CurrentFrame.N; ++i) claimed[i] = static_cast<MapPoint*>(CurrentFrame.mvpMapPoints[i]) != nullptr;  // :783
    std::vector<int32_t> hit;
    // the rotation bin of this overload is taken without the +360 wrap (:796-797)
    const int nmatches = gpu().SearchBest(view_of(CurrentFrame), claimed, Q, true, mbCheckOrientation ? ORB_ROT_NOWRAP : ORB_ROT_NONE, TH_HIGH, hit);
    for (size_t q = 0; q < who.size(); ++q)
        if (hit[q] >= 0) CurrentFrame.mvpMapPoints[hit[q]] = who[q];  // :793 (matches the rotation check dropped stay null, :811)
    return nmatches;
}

// ---- :820-894 ----------------------------------------------------------------------------------------------------
int ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist) {
    cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
    cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
    cv::Mat Ow = -Rcw.t() * tcw;
    const std::vector<MapPoint*> vpMPs = pKF->GetMapPointMatches();

    Queries Q;
    std::vector<MapPoint*> who;
    for (size_t i = 0; i < vpMPs.size(); ++i) {
        MapPoint* pMP = vpMPs[i];
        if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP)) continue;
        cv::Mat Xw = pMP->GetWorldPos();
        cv::Mat Xc = Rcw * Xw + tcw;
        const float invz = 1.0 / Xc.at<float>(2);  // :838: no sign test on the depth in this overload
        const float u = CurrentFrame.fx * Xc.at<float>(0) * invz + CurrentFrame.cx;
        const float v = CurrentFrame.fy * Xc.at<float>(1) * invz + CurrentFrame.cy;
        if (u < CurrentFrame.mnMinX || u > CurrentFrame.mnMaxX || v < CurrentFrame.mnMinY || v > CurrentFrame.mnMaxY) continue;
        const float dist3D = cv::norm(Xw - Ow);
        const int level = pMP->PredictScale(dist3D, &CurrentFrame);
        push_descriptor(Q, pMP->GetDescriptor());
        Q.u.push_back(u);
        Q.v.push_back(v);
        Q.radius.push_back(th * CurrentFrame.mvScaleFactors[level]);
        Q.minLevel.push_back(level - 1);  // :848
        Q.maxLevel.push_back(level + 1);
        Q.angle.push_back(pKF->mvKeys[i].angle);  // :872
        who.push_back(pMP);
    }
    if (who.empty()) return 0;
    std::vector<uint8_t> claimed(CurrentFrame.N);
    for (int i = 0; i < CurrentFrame.N; ++i) claimed[i] = static_cast<MapPoint*>(CurrentFrame.mvpMapPoints[i]) != nullptr;  // :857
    std::vector<int32_t> hit;
    const int nmatches = gpu().SearchBest(view_of(CurrentFrame), claimed, Q, true, mbCheckOrientation ? ORB_ROT_WRAP : ORB_ROT_NONE, ORBdist, hit);
    for (size_t q = 0; q < who.size(); ++q)
        if (hit[q] >= 0) CurrentFrame.mvpMapPoints[hit[q]] = who[q];  // :868
    return nmatches;
}

}  // namespace ORB_SLAM2
