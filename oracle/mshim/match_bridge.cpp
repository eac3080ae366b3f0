// C entry points over the reference's own ORB_SLAM2::ORBmatcher, compiled UNMODIFIED from /root/reference/src/ORBmatcher.cc
// against oracle/mshim (orbslam_objects.h: stand-in MapPoint / KeyFrame / Frame; mshim_cv.hpp: stand-in cv::Mat) --
// TEST INFRASTRUCTURE ONLY.  Every function takes the same flat arrays as the oracle's restatement of the same method
// (oracle/orb_oracle.cpp: orc_search_*), builds the object graph the method expects, runs the reference's method and
// flattens what it wrote.  Cameras are the identity (R = I, t = 0, fx = fy = 1, cx = cy = 0) and map points sit at depth 1,
// so a point meant to project to pixel (u, v) is the world point (u, v, 1) and the reference's own projection arithmetic
// returns exactly (u, v): the searches start from the same numbers on both sides.
#include <chrono>
#include <climits>
#include <thread>
#include "ORBmatcher.h"  // the reference's header; its three object headers are replaced by orbslam_objects.h

#include <climits>
#include <memory>

using namespace ORB_SLAM2;
typedef orb_oracle::KeyPoint OKP;

namespace {
cv::KeyPoint to_cv(const OKP& k) {
    cv::KeyPoint c;
    c.pt = cv::Point2f(k.x, k.y);
    c.size = k.size;
    c.angle = k.angle;
    c.response = k.response;
    c.octave = k.octave;
    c.class_id = k.class_id;
    return c;
}
cv::Mat desc_mat(const uint8_t* d, int n) {
    cv::Mat m(std::max(n, 1), 32, CV_8U);
    if (n) memcpy(m.data, d, (size_t)n * 32);
    return m;
}
cv::Mat vec3(float x, float y, float z) {
    cv::Mat m(3, 1, CV_32F);
    m.at<float>(0) = x;
    m.at<float>(1) = y;
    m.at<float>(2) = z;
    return m;
}
void fill_grid(GridHolder& g, const OrcFrame* F) {
    g.keys.assign(F->keysUn, F->keysUn + F->N);
    g.desc.assign(F->desc, F->desc + (size_t)F->N * 32);
    g.mnMinX = F->mnMinX;
    g.mnMinY = F->mnMinY;
    g.wInv = F->mfGridElementWidthInv;
    g.hInv = F->mfGridElementHeightInv;
    // image bounds from the grid pitch (64 x 48 cells, include/Frame.h:17-18)
    g.mnMaxX = F->mnMinX + 64.0f / F->mfGridElementWidthInv;
    g.mnMaxY = F->mnMinY + 48.0f / F->mfGridElementHeightInv;
}
void fill_frame(Frame& Fr, const OrcFrame* F, const float* scale, int nscale) {
    fill_grid(Fr.grid, F);
    Fr.N = F->N;
    Fr.mnMinX = Fr.grid.mnMinX;
    Fr.mnMinY = Fr.grid.mnMinY;
    Fr.mnMaxX = Fr.grid.mnMaxX;
    Fr.mnMaxY = Fr.grid.mnMaxY;
    Fr.mfGridElementWidthInv = Fr.grid.wInv;
    Fr.mfGridElementHeightInv = Fr.grid.hInv;
    for (int i = 0; i < F->N; ++i) Fr.mvKeysUn.push_back(to_cv(F->keysUn[i]));
    Fr.mvKeys = Fr.mvKeysUn;  // UndistortKeyPoints moves pt only (src/Frame.cc:384-414): angles and octaves are shared
    Fr.mDescriptors = desc_mat(F->desc, F->N);
    Fr.mvuRight.assign(F->N, -1.0f);
    if (scale) Fr.mvScaleFactors.assign(scale, scale + nscale);
    Fr.mvpMapPoints.v.assign(F->N, nullptr);
    Fr.mvbOutlier.assign(F->N, false);
}
struct Pool {
    std::vector<std::unique_ptr<MapPoint>> all;
    MapPoint* make(int id) {
        all.emplace_back(new MapPoint());
        all.back()->id = id;
        return all.back().get();
    }
};
DBoW2::FeatureVector featvec(const int* nodes, const int* off, const int* idx, int nn) {
    DBoW2::FeatureVector fv;
    for (int k = 0; k < nn; ++k)
        for (int j = off[k]; j < off[k + 1]; ++j) fv.addFeature(nodes[k], idx[j]);
    return fv;
}
const int NSCALE = 8;
}  // namespace

extern "C" {

// ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th), src/ORBmatcher.cc:19-65.
int refm_search_by_projection_map(const OrcFrame* F, const float* mvuRight, uint8_t* occupied, const float* mvScaleFactors, int nq,
                                  const uint8_t* qdesc, const float* projX, const float* projY, const float* projXR, const int* level,
                                  const float* viewCos, const uint8_t* qObserved, float th, float mfNNratio, int* featureOfQuery) {
    Frame Fr;
    fill_frame(Fr, F, mvScaleFactors, NSCALE);
    if (mvuRight) Fr.mvuRight.assign(mvuRight, mvuRight + F->N);
    Pool pool;
    for (int i = 0; i < F->N; ++i)
        if (occupied[i]) Fr.mvpMapPoints.v[i] = pool.make(-1);  // Observations() = 1
    std::vector<MapPoint*> q;
    for (int i = 0; i < nq; ++i) {
        MapPoint* p = pool.make(i);
        p->desc = desc_mat(qdesc + 32 * (size_t)i, 1);
        p->mbTrackInView = true;
        p->mTrackProjX = projX[i];
        p->mTrackProjY = projY[i];
        p->mTrackProjXR = projXR[i];
        p->mnTrackScaleLevel = level[i];
        p->mTrackViewCos = viewCos[i];
        p->nObs = (!qObserved || qObserved[i]) ? 1 : 0;
        q.push_back(p);
        featureOfQuery[i] = -1;
    }
    ORBmatcher m(mfNNratio, true);
    const int n = m.SearchByProjection(Fr, q, th);
    for (auto& e : Fr.mvpMapPoints.log)
        if (e.second && e.second->id >= 0) featureOfQuery[e.second->id] = (int)e.first;
    for (int i = 0; i < F->N; ++i) occupied[i] = Fr.mvpMapPoints.v[i] && Fr.mvpMapPoints.v[i]->nObs > 0;
    return n;
}

// ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono), src/ORBmatcher.cc:732-818.
// Query i = feature i of the last frame: its map point sits at (u, v, 1).  bForward / bBackward are produced the way the
// reference derives them: from the z translation between the two (identity-rotation) poses against CurrentFrame.mb.
int refm_search_by_projection_last(const OrcFrame* Cur, uint8_t* curHasMapPoint, const float* mvScaleFactors, int nq, const uint8_t* qdesc,
                                   const float* u, const float* v, const int* lastOctave, const float* lastAngle, float th, int bForward,
                                   int bBackward, int checkOri, int* featureOfQuery) {
    Frame C, L;
    fill_frame(C, Cur, mvScaleFactors, NSCALE);
    Pool pool;
    for (int i = 0; i < Cur->N; ++i)
        if (curHasMapPoint[i]) C.mvpMapPoints.v[i] = pool.make(-1);
    C.mb = 1.0f;
    L.N = nq;
    L.mvKeys.resize(nq);
    L.mvbOutlier.assign(nq, false);
    L.mvpMapPoints.v.assign(nq, nullptr);
    for (int i = 0; i < nq; ++i) {
        MapPoint* p = pool.make(i);
        p->desc = desc_mat(qdesc + 32 * (size_t)i, 1);
        p->pos = vec3(u[i], v[i], 1.0f);
        L.mvpMapPoints.v[i] = p;
        L.mvKeys[i].octave = lastOctave[i];
        L.mvKeys[i].angle = lastAngle[i];
        featureOfQuery[i] = -1;
    }
    L.mTcw.at<float>(2, 3) = bForward ? 2.0f : (bBackward ? -2.0f : 0.0f);  // tlc.z against mb = 1
    ORBmatcher m(0.9f, checkOri != 0);
    const int n = m.SearchByProjection(C, L, th, false);
    std::vector<int> owner(Cur->N, -1);
    for (auto& e : C.mvpMapPoints.log) {
        if (e.second) {
            featureOfQuery[e.second->id] = (int)e.first;
            owner[e.first] = e.second->id;
        } else if (owner[e.first] >= 0) {
            featureOfQuery[owner[e.first]] = -1;
            owner[e.first] = -1;
        }
    }
    for (int i = 0; i < Cur->N; ++i) curHasMapPoint[i] = C.mvpMapPoints.v[i] != nullptr;
    return n;
}

// ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12), src/ORBmatcher.cc:278-366.
int refm_search_by_bow_kf(const uint8_t* d1, const float* ang1, const uint8_t* has1, int n1, const uint8_t* d2, const float* ang2,
                          const uint8_t* has2, int n2, const int* nodes1, const int* off1, const int* idx1, int nn1, const int* nodes2,
                          const int* off2, const int* idx2, int nn2, float nnratio, int checkOri, int* matches12) {
    KeyFrame K1, K2;
    Pool pool;
    K1.N = n1;
    K2.N = n2;
    K1.mvKeysUn.resize(n1);
    K2.mvKeysUn.resize(n2);
    K1.mapPoints.assign(n1, nullptr);
    K2.mapPoints.assign(n2, nullptr);
    for (int i = 0; i < n1; ++i) {
        K1.mvKeysUn[i].angle = ang1[i];
        if (has1[i]) K1.mapPoints[i] = pool.make(i);
    }
    for (int i = 0; i < n2; ++i) {
        K2.mvKeysUn[i].angle = ang2[i];
        if (has2[i]) K2.mapPoints[i] = pool.make(i);
    }
    K1.mDescriptors = desc_mat(d1, n1);
    K2.mDescriptors = desc_mat(d2, n2);
    K1.mFeatVec = featvec(nodes1, off1, idx1, nn1);
    K2.mFeatVec = featvec(nodes2, off2, idx2, nn2);
    std::vector<MapPoint*> vp;
    ORBmatcher m(nnratio, checkOri != 0);
    const int n = m.SearchByBoW(&K1, &K2, vp);
    for (int i = 0; i < n1; ++i) matches12[i] = (i < (int)vp.size() && vp[i]) ? vp[i]->id : -1;
    return n;
}

// ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo = false), src/ORBmatcher.cc:368-467.
int refm_search_for_triangulation(const uint8_t* d1, const float* x1, const float* y1, const float* ang1, const uint8_t* has1, int n1,
                                  const uint8_t* d2, const float* x2, const float* y2, const float* ang2, const int* oct2,
                                  const uint8_t* has2, int n2, const int* nodes1, const int* off1, const int* idx1, int nn1,
                                  const int* nodes2, const int* off2, const int* idx2, int nn2, const float* F12, const float* sigma2,
                                  int checkOri, int* matches12) {
    KeyFrame K1, K2;
    Pool pool;
    K1.N = n1;
    K2.N = n2;
    K1.mvKeysUn.resize(n1);
    K2.mvKeysUn.resize(n2);
    K1.mapPoints.assign(n1, nullptr);
    K2.mapPoints.assign(n2, nullptr);
    K1.mvuRight.assign(n1, -1.0f);
    K2.mvuRight.assign(n2, -1.0f);
    for (int i = 0; i < n1; ++i) {
        K1.mvKeysUn[i].pt = cv::Point2f(x1[i], y1[i]);
        K1.mvKeysUn[i].angle = ang1[i];
        if (has1[i]) K1.mapPoints[i] = pool.make(i);
    }
    for (int i = 0; i < n2; ++i) {
        K2.mvKeysUn[i].pt = cv::Point2f(x2[i], y2[i]);
        K2.mvKeysUn[i].angle = ang2[i];
        K2.mvKeysUn[i].octave = oct2[i];
        if (has2[i]) K2.mapPoints[i] = pool.make(i);
    }
    K2.mvLevelSigma2.assign(sigma2, sigma2 + NSCALE);
    K2.tcw = vec3(0.0f, 0.0f, 1.0f);  // the (unused) epipole stays finite
    K1.mDescriptors = desc_mat(d1, n1);
    K2.mDescriptors = desc_mat(d2, n2);
    K1.mFeatVec = featvec(nodes1, off1, idx1, nn1);
    K2.mFeatVec = featvec(nodes2, off2, idx2, nn2);
    cv::Mat F(3, 3, CV_32F);
    for (int i = 0; i < 9; ++i) F.at<float>(i / 3, i % 3) = F12[i];
    std::vector<std::pair<size_t, size_t>> pairs;
    ORBmatcher m(0.6f, checkOri != 0);
    const int n = m.SearchForTriangulation(&K1, &K2, F, pairs, false);
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    for (auto& p : pairs) matches12[p.first] = (int)p.second;
    return n;
}

// ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize), src/ORBmatcher.cc:197-276.
int refm_search_for_initialization(const OKP* keys1, const uint8_t* desc1, int n1, const OrcFrame* F2, float* vbPrevMatched, int windowSize,
                                   float nnratio, int checkOri, int* matches12) {
    Frame A, B;
    fill_frame(B, F2, nullptr, 0);
    A.N = n1;
    for (int i = 0; i < n1; ++i) A.mvKeysUn.push_back(to_cv(keys1[i]));
    A.mDescriptors = desc_mat(desc1, n1);
    std::vector<cv::Point2f> prev(n1);
    for (int i = 0; i < n1; ++i) prev[i] = cv::Point2f(vbPrevMatched[2 * i], vbPrevMatched[2 * i + 1]);
    std::vector<int> m12;
    ORBmatcher m(nnratio, checkOri != 0);
    const int n = m.SearchForInitialization(A, B, prev, m12, windowSize);
    for (int i = 0; i < n1; ++i) {
        matches12[i] = m12[i];
        vbPrevMatched[2 * i] = prev[i].x;
        vbPrevMatched[2 * i + 1] = prev[i].y;
    }
    return n;
}

// ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*>&, th, ORBdist), src/ORBmatcher.cc:820-894
// (relocalisation).  Query i = map point i of the keyframe, at (u, v, 1); kfAngle = pKF->mvKeys[i].angle.
int refm_search_by_projection_reloc(const OrcFrame* Cur, uint8_t* curHasMapPoint, const float* mvScaleFactors, int nq, const uint8_t* qdesc,
                                    const float* u, const float* v, const int* predictedLevel, const float* kfAngle, float th,
                                    int ORBdist, int checkOri, int* featureOfQuery) {
    Frame C;
    fill_frame(C, Cur, mvScaleFactors, NSCALE);
    Pool pool;
    for (int i = 0; i < Cur->N; ++i)
        if (curHasMapPoint[i]) C.mvpMapPoints.v[i] = pool.make(-1);
    KeyFrame K;
    K.N = nq;
    K.mvKeys.resize(nq);
    K.mapPoints.assign(nq, nullptr);
    for (int i = 0; i < nq; ++i) {
        MapPoint* p = pool.make(i);
        p->desc = desc_mat(qdesc + 32 * (size_t)i, 1);
        p->pos = vec3(u[i], v[i], 1.0f);
        p->predictedLevel = predictedLevel[i];
        K.mapPoints[i] = p;
        K.mvKeys[i].angle = kfAngle[i];
        featureOfQuery[i] = -1;
    }
    std::set<MapPoint*> found;
    ORBmatcher m(0.9f, checkOri != 0);
    const int n = m.SearchByProjection(C, &K, found, th, ORBdist);
    std::vector<int> owner(Cur->N, -1);
    for (auto& e : C.mvpMapPoints.log) {
        if (e.second) {
            featureOfQuery[e.second->id] = (int)e.first;
            owner[e.first] = e.second->id;
        } else if (owner[e.first] >= 0) {
            featureOfQuery[owner[e.first]] = -1;
            owner[e.first] = -1;
        }
    }
    for (int i = 0; i < Cur->N; ++i) curHasMapPoint[i] = C.mvpMapPoints.v[i] != nullptr;
    return n;
}

// ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, vpPoints, vpMatched, th), src/ORBmatcher.cc:121-195 (loop closing).
// The method's radius is th * pKF->mvScaleFactors[PredictScale()]: with th = 1, PredictScale() = the query's own index and
// mvScaleFactors[i] = radius[i] it searches exactly the window the caller asks for.  claimed = vpMatched[idx] != NULL.
int refm_search_by_projection_loop(const OrcFrame* KF, uint8_t* claimed, int nq, const uint8_t* qdesc, const float* u, const float* v,
                                   const float* radius, int* featureOfQuery) {
    KeyFrame K;
    fill_grid(K.grid, KF);
    K.grid_members();
    K.N = KF->N;
    for (int i = 0; i < KF->N; ++i) K.mvKeysUn.push_back(to_cv(KF->keysUn[i]));
    K.mDescriptors = desc_mat(KF->desc, KF->N);
    K.mvScaleFactors.assign(radius, radius + nq);
    Pool pool;
    std::vector<MapPoint*> matched(KF->N, nullptr), pts;
    for (int i = 0; i < KF->N; ++i)
        if (claimed[i]) matched[i] = pool.make(-1);
    for (int i = 0; i < nq; ++i) {
        MapPoint* p = pool.make(i);
        p->desc = desc_mat(qdesc + 32 * (size_t)i, 1);
        p->pos = vec3(u[i], v[i], 1.0f);
        p->normal = vec3(0.0f, 0.0f, 1e6f);  // passes the viewing-angle gate PO.Pn >= 0.5 |PO|
        p->predictedLevel = i;
        pts.push_back(p);
        featureOfQuery[i] = -1;
    }
    ORBmatcher m(0.75f, true);
    const int n = m.SearchByProjection(&K, cv::Mat::eye(4, 4, CV_32F), pts, matched, 1);
    for (int i = 0; i < KF->N; ++i) {
        if (matched[i] && matched[i]->id >= 0) featureOfQuery[matched[i]->id] = i;
        claimed[i] = matched[i] != nullptr;
    }
    return n;
}

// ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, s12 = 1, R12 = I, t12 = 0, th), src/ORBmatcher.cc:636-730.  Query i = map point
// i of KF1 at (u, v, 1) with PredictScale() = level[i]; every feature of KF2 holds a map point, so vpMatches12[i] names the feature.
int refm_search_by_sim3(const OrcFrame* KF2, const float* mvScaleFactors, int nq, const uint8_t* qdesc, const float* u, const float* v,
                        const int* level, float th, int* featureOfQuery) {
    KeyFrame K1, K2;
    fill_grid(K2.grid, KF2);
    K2.grid_members();
    K2.N = KF2->N;
    for (int i = 0; i < KF2->N; ++i) K2.mvKeysUn.push_back(to_cv(KF2->keysUn[i]));
    K2.mDescriptors = desc_mat(KF2->desc, KF2->N);
    K2.mvScaleFactors.assign(mvScaleFactors, mvScaleFactors + NSCALE);
    Pool pool;
    K2.mapPoints.assign(KF2->N, nullptr);
    for (int i = 0; i < KF2->N; ++i) K2.mapPoints[i] = pool.make(i);
    K1.N = nq;
    K1.mapPoints.assign(nq, nullptr);
    for (int i = 0; i < nq; ++i) {
        MapPoint* p = pool.make(-1);
        p->desc = desc_mat(qdesc + 32 * (size_t)i, 1);
        p->pos = vec3(u[i], v[i], 1.0f);
        p->predictedLevel = level[i];
        K1.mapPoints[i] = p;
    }
    std::vector<MapPoint*> m12(nq, nullptr);
    ORBmatcher m(0.75f, true);
    const float s12 = 1.0f;
    const int n = m.SearchBySim3(&K1, &K2, m12, s12, cv::Mat::eye(3, 3, CV_32F), cv::Mat::zeros(3, 1, CV_32F), th);
    for (int i = 0; i < nq; ++i) featureOfQuery[i] = m12[i] ? m12[i]->id : -1;
    return n;
}

// ORBmatcher::Fuse(pKF, vpMapPoints, th) (src/ORBmatcher.cc:504-568; variant 0) and Fuse(pKF, Scw, vpPoints, th, vpReplacePoint)
// (:570-634; variant 1), search part.  Every second feature of the keyframe holds a map point with more observations than the
// queries, so a hit on it shows up as pMP->Replace(that point) / vpReplacePoint, a hit on a free feature as pMP->AddObservation(pKF, idx):
// either way the feature the reference chose is recorded.
int refm_fuse_search(const OrcFrame* KF, const float* mvScaleFactors, int nq, const uint8_t* qdesc, const float* u, const float* v,
                     const int* level, float th, int variant, int* featureOfQuery) {
    KeyFrame K;
    fill_grid(K.grid, KF);
    K.grid_members();
    K.N = KF->N;
    for (int i = 0; i < KF->N; ++i) K.mvKeysUn.push_back(to_cv(KF->keysUn[i]));
    K.mDescriptors = desc_mat(KF->desc, KF->N);
    K.mvScaleFactors.assign(mvScaleFactors, mvScaleFactors + NSCALE);
    Pool pool;
    K.mapPoints.assign(KF->N, nullptr);
    for (int i = 0; i < KF->N; i += 2) {
        K.mapPoints[i] = pool.make(i);
        K.mapPoints[i]->nObs = 5;
    }
    std::vector<MapPoint*> pts;
    for (int i = 0; i < nq; ++i) {
        MapPoint* p = pool.make(-1);
        p->desc = desc_mat(qdesc + 32 * (size_t)i, 1);
        p->pos = vec3(u[i], v[i], 1.0f);
        p->normal = vec3(0.0f, 0.0f, 1e6f);
        p->predictedLevel = level[i];
        pts.push_back(p);
        featureOfQuery[i] = -1;
    }
    ORBmatcher m(0.6f, true);
    std::vector<MapPoint*> repl(nq, nullptr);
    const int n = variant == 0 ? m.Fuse(&K, pts, th) : m.Fuse(&K, cv::Mat::eye(4, 4, CV_32F), pts, th, repl);
    for (int i = 0; i < nq; ++i) {
        if (!pts[i]->addedObservations.empty()) featureOfQuery[i] = (int)pts[i]->addedObservations[0].second;
        if (pts[i]->replacedBy) featureOfQuery[i] = pts[i]->replacedBy->id;
        if (repl[i]) featureOfQuery[i] = repl[i]->id;
    }
    return n;
}

// ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:896-908.
int refm_descriptor_distance(const uint8_t* a, const uint8_t* b) { return ORBmatcher::DescriptorDistance(desc_mat(a, 1), desc_mat(b, 1)); }

// BASELINE config 4 on the host cores: the shared scan (src/ORBmatcher.cc:49-55) over all train rows in index order, every
// distance from the reference's own ORBmatcher::DescriptorDistance; npairs independent (query set, train set) pairs of
// nq x nt rows spread over nthreads std::threads.  Returns the wall-clock seconds of the scan (buffers are the caller's).
double refm_bruteforce_many(const uint8_t* q, const uint8_t* t, int npairs, int nq, int nt, int nthreads, int32_t* best_idx,
                            int32_t* best_dist, int32_t* second_dist) {
    auto work = [&](int p0, int p1) {
        for (int p = p0; p < p1; ++p) {
            const cv::Mat Q = desc_mat(q + (size_t)p * nq * 32, nq), T = desc_mat(t + (size_t)p * nt * 32, nt);
            for (int i = 0; i < nq; ++i) {
                const cv::Mat dq = Q.row(i);
                int best = INT_MAX, second = INT_MAX, idx = -1;
                for (int j = 0; j < nt; ++j) {
                    const int d = ORBmatcher::DescriptorDistance(dq, T.row(j));
                    if (d < best) {
                        second = best;
                        best = d;
                        idx = j;
                    } else if (d < second) {
                        second = d;
                    }
                }
                best_idx[(size_t)p * nq + i] = idx;
                best_dist[(size_t)p * nq + i] = best;
                second_dist[(size_t)p * nq + i] = second;
            }
        }
    };
    const auto t0 = std::chrono::steady_clock::now();
    nthreads = std::max(1, std::min(nthreads, npairs));
    std::vector<std::thread> th;
    for (int k = 0; k < nthreads; ++k) th.emplace_back(work, (int)((long long)npairs * k / nthreads), (int)((long long)npairs * (k + 1) / nthreads));
    for (auto& x : th) x.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
}
