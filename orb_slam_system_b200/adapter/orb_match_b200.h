// Hamming-search helpers for the reference's ORBmatcher / Frame over the liborb_b200 C ABI.
//
// The reference's eleven Search*/Fuse methods (src/ORBmatcher.cc) and Frame::ComputeStereoMatches
// (src/Frame.cc:446-529) all reduce to: build a candidate list per query, scan it with
// DescriptorDistance keeping best (and second best), then apply an accept rule and -- in several
// methods -- a greedy "already matched" state.  The adapter keeps every signature and all host-side
// gating in place and replaces only the inner scans: candidates are gathered into CSR form, one
// orb_match_csr call returns (bestIdx, bestDist, secondDist) for all queries, and the method's own
// accept rule is replayed in query order on the host (INTEGRATION.md shows SearchByProjection).
#ifndef ORB_MATCH_B200_H
#define ORB_MATCH_B200_H

#include <climits>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include <opencv2/core/core.hpp>

#include "orb_b200.h"

namespace orb_b200 {

class Matcher {
public:
    explicit Matcher(int device = 0) : m_(nullptr) {
        if (orb_matcher_create(device, &m_) != ORB_OK) throw std::runtime_error(std::string("orb_matcher_create: ") + orb_last_error());
    }
    ~Matcher() { orb_matcher_destroy(m_); }
    Matcher(const Matcher&) = delete;
    Matcher& operator=(const Matcher&) = delete;

    struct Scan {
        std::vector<int32_t> bestIdx, bestDist, secondDist;
    };

    // ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:896-908) of two 1x32 CV_8U rows.
    int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
        int32_t bi = -1, bd = INT_MAX, sd = INT_MAX;
        check(orb_match_all(m_, a.ptr<uint8_t>(0), 1, b.ptr<uint8_t>(0), 1, &bi, &bd, &sd));
        return bd;
    }

    // Every query row of `q` against all rows of `t` in index order (BASELINE config 4).
    Scan MatchAll(const cv::Mat& q, const cv::Mat& t) {
        Scan s = make(q.rows);
        check(orb_match_all(m_, q.ptr<uint8_t>(0), q.rows, t.ptr<uint8_t>(0), t.rows, s.bestIdx.data(), s.bestDist.data(), s.secondDist.data()));
        return s;
    }

    // Query i scans t.row(cand[offsets[i] .. offsets[i+1])) in that order.  lastMin = the
    // SearchForTriangulation rule (src/ORBmatcher.cc:404-419) with maxDist = TH_LOW.
    Scan MatchWindows(const cv::Mat& q, const cv::Mat& t, const std::vector<int32_t>& offsets, const std::vector<int32_t>& cand,
                      bool lastMin = false, int maxDist = 50) {
        Scan s = make(q.rows);
        check(orb_match_csr(m_, q.ptr<uint8_t>(0), q.rows, t.ptr<uint8_t>(0), t.rows, offsets.data(), cand.data(),
                            lastMin ? ORB_TIE_LAST_MIN : ORB_TIE_FIRST_MIN, maxDist, s.bestIdx.data(), s.bestDist.data(), s.secondDist.data()));
        return s;
    }

    // Hamming part of Frame::ComputeStereoMatches: best right index (-1: none) and distance per left keypoint.
    void StereoMatch(const std::vector<cv::KeyPoint>& kl, const cv::Mat& dl, const std::vector<cv::KeyPoint>& kr, const cv::Mat& dr,
                     const std::vector<float>& scaleFactors, int rows, float bf, float fx, std::vector<int32_t>& bestR,
                     std::vector<int32_t>& bestDist) {
        static_assert(sizeof(cv::KeyPoint) == sizeof(orb_keypoint), "cv::KeyPoint layout");
        bestR.assign(kl.size(), -1);
        bestDist.assign(kl.size(), 100);
        check(orb_stereo_match(m_, reinterpret_cast<const orb_keypoint*>(kl.data()), dl.ptr<uint8_t>(0), (int)kl.size(),
                               reinterpret_cast<const orb_keypoint*>(kr.data()), dr.ptr<uint8_t>(0), (int)kr.size(), scaleFactors.data(),
                               (int)scaleFactors.size(), rows, bf, fx, bestR.data(), bestDist.data()));
    }

    // ---- the reference's search methods, whole (include/orb_b200.h "search methods"; INTEGRATION.md shows how
    //      ORBmatcher.cc gathers the arrays from Frame / KeyFrame / MapPoint) ----

    // What GetFeaturesInArea reads of a Frame / KeyFrame (src/Frame.cc:307-360, src/KeyFrame.cc:549-588).
    static orb_frame_view View(const std::vector<cv::KeyPoint>& keysUn, const cv::Mat& descriptors, float mnMinX, float mnMinY,
                               float gridElementWidthInv, float gridElementHeightInv) {
        static_assert(sizeof(cv::KeyPoint) == sizeof(orb_keypoint), "cv::KeyPoint layout");
        orb_frame_view v;
        v.keys_un = reinterpret_cast<const orb_keypoint*>(keysUn.data());
        v.desc = descriptors.ptr<uint8_t>(0);
        v.n = (int)keysUn.size();
        v.min_x = mnMinX;
        v.min_y = mnMinY;
        v.grid_w_inv = gridElementWidthInv;
        v.grid_h_inv = gridElementHeightInv;
        return v;
    }

    struct Queries {  // the map points that survive a method's own gating, in visiting order
        std::vector<uint8_t> desc;  // 32 bytes each (MapPoint::GetDescriptor)
        std::vector<float> u, v, uR, radius, viewCos, angle;
        std::vector<int32_t> level, minLevel, maxLevel;
        std::vector<uint8_t> observed;
        int size() const { return (int)u.size(); }
    };

    // SearchByProjection(Frame&, const vector<MapPoint*>&, th), src/ORBmatcher.cc:19-65.  occupied[i] =
    // F.mvpMapPoints[i] && Observations() > 0 (updated).  Returns nmatches; featureOfQuery[q] = index written.
    int SearchByProjection(const orb_frame_view& F, const std::vector<float>& mvuRight, std::vector<uint8_t>& occupied,
                           const std::vector<float>& mvScaleFactors, const Queries& Q, float th, float nnratio,
                           std::vector<int32_t>& featureOfQuery) {
        featureOfQuery.assign(Q.size(), -1);
        int n = 0;
        check(orb_search_by_projection_map(m_, &F, mvuRight.empty() ? nullptr : mvuRight.data(), occupied.data(), mvScaleFactors.data(),
                                           (int)mvScaleFactors.size(), Q.size(), Q.desc.data(), Q.u.data(), Q.v.data(), Q.uR.data(),
                                           Q.level.data(), Q.viewCos.data(), Q.observed.empty() ? nullptr : Q.observed.data(), th, nnratio,
                                           featureOfQuery.data(), &n));
        return n;
    }

    // The best-only projection searches (:732-818, :820-894, :121-195, :636-730, :504-634); claimed may be empty.
    int SearchBest(const orb_frame_view& F, std::vector<uint8_t>& claimed, const Queries& Q, bool levels, int rotMode, int maxDist,
                   std::vector<int32_t>& featureOfQuery) {
        featureOfQuery.assign(Q.size(), -1);
        int n = 0;
        check(orb_search_by_projection_best(m_, &F, claimed.empty() ? nullptr : claimed.data(), Q.size(), Q.desc.data(), Q.u.data(),
                                            Q.v.data(), Q.radius.data(), levels ? Q.minLevel.data() : nullptr,
                                            levels ? Q.maxLevel.data() : nullptr, rotMode == ORB_ROT_NONE ? nullptr : Q.angle.data(),
                                            rotMode, maxDist, featureOfQuery.data(), &n));
        return n;
    }

    // SearchForInitialization, :197-276 (vbPrevMatched is updated like :270-273).
    int SearchForInitialization(const std::vector<cv::KeyPoint>& keysUn1, const cv::Mat& desc1, const orb_frame_view& F2,
                                std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize, float nnratio,
                                bool checkOri) {
        static_assert(sizeof(cv::Point2f) == 2 * sizeof(float), "cv::Point2f layout");
        vnMatches12.assign(keysUn1.size(), -1);
        int n = 0;
        check(orb_search_for_initialization(m_, reinterpret_cast<const orb_keypoint*>(keysUn1.data()), desc1.ptr<uint8_t>(0),
                                            (int)keysUn1.size(), &F2, reinterpret_cast<float*>(vbPrevMatched.data()), windowSize, nnratio,
                                            checkOri ? 1 : 0, vnMatches12.data(), &n));
        return n;
    }

    // DBoW2::FeatureVector (a std::map<NodeId, std::vector<unsigned>>) flattened for the C ABI.
    struct FlatFeatureVector {
        std::vector<int32_t> nodes, off, idx;
        template <class Map>
        explicit FlatFeatureVector(const Map& fv) {
            off.push_back(0);
            for (const auto& e : fv) {
                nodes.push_back((int32_t)e.first);
                for (auto i : e.second) idx.push_back((int32_t)i);
                off.push_back((int32_t)idx.size());
            }
        }
        orb_feature_vector view() const { return orb_feature_vector{nodes.data(), off.data(), idx.data(), (int32_t)nodes.size()}; }
    };

    // SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12), :278-366: matches12[i1] = i2 or -1.
    int SearchByBoW(const cv::Mat& d1, const std::vector<float>& angle1, const std::vector<uint8_t>& hasMp1, const FlatFeatureVector& fv1,
                    const cv::Mat& d2, const std::vector<float>& angle2, const std::vector<uint8_t>& hasMp2, const FlatFeatureVector& fv2,
                    float nnratio, bool checkOri, std::vector<int32_t>& matches12) {
        matches12.assign(d1.rows, -1);
        const orb_feature_vector a = fv1.view(), b = fv2.view();
        int n = 0;
        check(orb_search_by_bow(m_, d1.ptr<uint8_t>(0), angle1.data(), hasMp1.data(), d1.rows, d2.ptr<uint8_t>(0), angle2.data(),
                                hasMp2.data(), d2.rows, &a, &b, nnratio, checkOri ? 1 : 0, matches12.data(), &n));
        return n;
    }

    // SearchForTriangulation (bOnlyStereo = false), :368-467: vMatchedPairs in i1 order.
    int SearchForTriangulation(const std::vector<cv::KeyPoint>& keysUn1, const cv::Mat& d1, const std::vector<uint8_t>& hasMp1,
                               const FlatFeatureVector& fv1, const std::vector<cv::KeyPoint>& keysUn2, const cv::Mat& d2,
                               const std::vector<uint8_t>& hasMp2, const FlatFeatureVector& fv2, const float F12[9],
                               const std::vector<float>& mvLevelSigma2, bool checkOri, std::vector<std::pair<size_t, size_t>>& vMatchedPairs) {
        std::vector<int32_t> m12(keysUn1.size(), -1);
        const orb_feature_vector a = fv1.view(), b = fv2.view();
        int n = 0;
        check(orb_search_for_triangulation(m_, reinterpret_cast<const orb_keypoint*>(keysUn1.data()), d1.ptr<uint8_t>(0), hasMp1.data(),
                                           (int)keysUn1.size(), reinterpret_cast<const orb_keypoint*>(keysUn2.data()), d2.ptr<uint8_t>(0),
                                           hasMp2.data(), (int)keysUn2.size(), &a, &b, F12, mvLevelSigma2.data(), (int)mvLevelSigma2.size(),
                                           checkOri ? 1 : 0, m12.data(), &n));
        vMatchedPairs.clear();
        vMatchedPairs.reserve(n);
        for (size_t i = 0; i < m12.size(); ++i)
            if (m12[i] >= 0) vMatchedPairs.push_back(std::make_pair(i, (size_t)m12[i]));
        return n;
    }

    // Frame::ComputeStereoMatches whole (src/Frame.cc:446-619): fills mvuRight / mvDepth from the two extractors' resident pyramids.
    void ComputeStereoMatches(orb_extractor* left, orb_extractor* right, const std::vector<cv::KeyPoint>& mvKeys, const cv::Mat& mDescriptors,
                              const std::vector<cv::KeyPoint>& mvKeysRight, const cv::Mat& mDescriptorsRight, float mbf, float fx,
                              std::vector<float>& mvuRight, std::vector<float>& mvDepth) {
        mvuRight.assign(mvKeys.size(), -1.0f);
        mvDepth.assign(mvKeys.size(), -1.0f);
        check(orb_compute_stereo_matches(m_, left, 0, right, 0, reinterpret_cast<const orb_keypoint*>(mvKeys.data()),
                                         mDescriptors.ptr<uint8_t>(0), (int)mvKeys.size(),
                                         reinterpret_cast<const orb_keypoint*>(mvKeysRight.data()), mDescriptorsRight.ptr<uint8_t>(0),
                                         (int)mvKeysRight.size(), mbf, fx, mvuRight.data(), mvDepth.data()));
    }

    // The same with Frame::mb as the reference's function reads it (see orb_compute_stereo_matches_mb).
    void ComputeStereoMatchesMb(orb_extractor* left, orb_extractor* right, const std::vector<cv::KeyPoint>& mvKeys, const cv::Mat& mDescriptors,
                                const std::vector<cv::KeyPoint>& mvKeysRight, const cv::Mat& mDescriptorsRight, float mbf, float mb,
                                std::vector<float>& mvuRight, std::vector<float>& mvDepth) {
        mvuRight.assign(mvKeys.size(), -1.0f);
        mvDepth.assign(mvKeys.size(), -1.0f);
        check(orb_compute_stereo_matches_mb(m_, left, 0, right, 0, reinterpret_cast<const orb_keypoint*>(mvKeys.data()),
                                            mDescriptors.ptr<uint8_t>(0), (int)mvKeys.size(),
                                            reinterpret_cast<const orb_keypoint*>(mvKeysRight.data()), mDescriptorsRight.ptr<uint8_t>(0),
                                            (int)mvKeysRight.size(), mbf, mb, mvuRight.data(), mvDepth.data()));
    }

private:
    static Scan make(int n) {
        Scan s;
        s.bestIdx.assign(n, -1);
        s.bestDist.assign(n, INT_MAX);
        s.secondDist.assign(n, INT_MAX);
        return s;
    }
    static void check(int rc) {
        if (rc != ORB_OK) throw std::runtime_error(std::string("orb_b200: ") + orb_last_error());
    }
    orb_matcher* m_;
};

}  // namespace orb_b200

#endif
