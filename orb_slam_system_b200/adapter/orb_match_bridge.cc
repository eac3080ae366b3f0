// C bridge over orb_b200::Matcher (adapter/orb_match_b200.h) so that the tests can drive the C++ adapter layer -- the
// code a maintainer pastes into ORBmatcher.cc / Frame.cc -- exactly like the reference-side bridge drives the reference.
#include <cstring>
#include <map>
#include <vector>

#include "orb_match_b200.h"

namespace {
cv::Mat rows32(const uint8_t* d, int n) {
    cv::Mat m(n > 0 ? n : 1, 32, CV_8U);
    if (n > 0) memcpy(m.data, d, (size_t)n * 32);
    m.rows = n;
    return m;
}
std::map<unsigned, std::vector<unsigned>> to_map(const int* nodes, const int* off, const int* idx, int nn) {
    std::map<unsigned, std::vector<unsigned>> fv;
    for (int k = 0; k < nn; ++k) fv[(unsigned)nodes[k]] = std::vector<unsigned>(idx + off[k], idx + off[k + 1]);
    return fv;
}
}  // namespace

extern "C" {

// SearchByBoW(KeyFrame*, KeyFrame*) through the adapter: DBoW2::FeatureVector-shaped maps in, matches12 out.
int adapter_search_by_bow(const uint8_t* d1, const float* ang1, const uint8_t* has1, int n1, const uint8_t* d2, const float* ang2,
                          const uint8_t* has2, int n2, const int* nodes1, const int* off1, const int* idx1, int nn1, const int* nodes2,
                          const int* off2, const int* idx2, int nn2, float nnratio, int checkOri, int* matches12) {
    try {
        orb_b200::Matcher m;
        orb_b200::Matcher::FlatFeatureVector f1(to_map(nodes1, off1, idx1, nn1)), f2(to_map(nodes2, off2, idx2, nn2));
        std::vector<int32_t> out;
        const int n = m.SearchByBoW(rows32(d1, n1), std::vector<float>(ang1, ang1 + n1), std::vector<uint8_t>(has1, has1 + n1), f1,
                                    rows32(d2, n2), std::vector<float>(ang2, ang2 + n2), std::vector<uint8_t>(has2, has2 + n2), f2, nnratio,
                                    checkOri != 0, out);
        memcpy(matches12, out.data(), sizeof(int) * (size_t)n1);
        return n;
    } catch (const std::exception&) {
        return -1000;
    }
}

// SearchForInitialization through the adapter.
int adapter_search_for_initialization(const orb_keypoint* k1, const uint8_t* d1, int n1, const orb_keypoint* k2, const uint8_t* d2, int n2,
                                      float minX, float minY, float wInv, float hInv, float* prevMatched, int windowSize, float nnratio,
                                      int checkOri, int* matches12) {
    try {
        orb_b200::Matcher m;
        std::vector<cv::KeyPoint> ku1(n1), ku2(n2);
        memcpy((void*)ku1.data(), k1, sizeof(orb_keypoint) * (size_t)n1);
        memcpy((void*)ku2.data(), k2, sizeof(orb_keypoint) * (size_t)n2);
        const cv::Mat D1 = rows32(d1, n1), D2 = rows32(d2, n2);
        const orb_frame_view F2 = orb_b200::Matcher::View(ku2, D2, minX, minY, wInv, hInv);
        std::vector<cv::Point2f> prev(n1);
        memcpy((void*)prev.data(), prevMatched, sizeof(float) * 2 * (size_t)n1);
        std::vector<int> out;
        const int n = m.SearchForInitialization(ku1, D1, F2, prev, out, windowSize, nnratio, checkOri != 0);
        memcpy(prevMatched, prev.data(), sizeof(float) * 2 * (size_t)n1);
        memcpy(matches12, out.data(), sizeof(int) * (size_t)n1);
        return n;
    } catch (const std::exception&) {
        return -1000;
    }
}
}
