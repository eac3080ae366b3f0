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
