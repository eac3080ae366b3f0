// CPU oracle for the ORB front end -- TEST INFRASTRUCTURE ONLY (see orb_oracle.h).
//
// Build: g++ -O2 -ffp-contract=off -std=c++17 -fPIC -shared (oracle/Makefile).
// -ffp-contract=off is part of the definition: the float expressions below are
// evaluated step by step in binary32, never fused (SURVEY 7 "Float reproducibility").
#include "orb_oracle.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstring>
#include <list>
#include <map>
#include <thread>

namespace orb_oracle {

// ---------------------------------------------------------------- cvRound
// OpenCV cvRound == lrint under the default rounding mode (round half to even), A.5.
int cv_round_f(float v) { return (int)lrintf(v); }
int cv_round_d(double v) { return (int)lrint(v); }

// ---------------------------------------------------------------- fastAtan2
// cv::fastAtan2 (A.4); called from IC_Angle, reference ORBextractor.cc:47.
float fast_atan2(float y, float x) {
    const float scale = (float)(180.0 / 3.141592653589793238462643383279502884);
    const float p1 = 0.9997878412794807f * scale;
    const float p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale;
    const float p7 = -0.04432655554792128f * scale;
    const float eps = (float)2.2204460492503131e-16;
    float ax = std::fabs(x), ay = std::fabs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + eps);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + eps);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ---------------------------------------------------------------- resize
// cv::resize INTER_LINEAR 8UC1 (A.2); called at reference ORBextractor.cc:511.
static void linear_axis(int ssize, int dsize, std::vector<int>& ofs, std::vector<int>& c0,
                        std::vector<int>& c1, bool clamp_coeff) {
    ofs.resize(dsize);
    c0.resize(dsize);
    c1.resize(dsize);
    double inv_scale = (double)dsize / ssize;
    double scale = 1.0 / inv_scale;
    for (int d = 0; d < dsize; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)std::floor(f);
        f -= s;
        if (clamp_coeff) {  // x axis: coefficients collapse at the image edge
            if (s < 0) { s = 0; f = 0.f; }
            if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
        }
        ofs[d] = s;
        // saturate_cast<short>(float * 2048): cvRound then clamp (never clamps here)
        c0[d] = cv_round_f((1.f - f) * 2048.f);
        c1[d] = cv_round_f(f * 2048.f);
    }
}

void resize_linear_u8(const Image& src, uint8_t* dst, int drows, int dcols, int dstep) {
    if (drows == src.rows && dcols == src.cols) {  // same-size resize is a copy
        for (int y = 0; y < drows; ++y) memcpy(dst + (size_t)y * dstep, src.row(y), dcols);
        return;
    }
    std::vector<int> xo, a0, a1, yo, b0, b1;
    linear_axis(src.cols, dcols, xo, a0, a1, true);
    // y axis: OpenCV keeps the coefficients and clips the two row indices instead.
    linear_axis(src.rows, drows, yo, b0, b1, false);
    std::vector<int> r0(dcols), r1(dcols);
    for (int y = 0; y < drows; ++y) {
        int sy0 = std::min(std::max(yo[y], 0), src.rows - 1);
        int sy1 = std::min(std::max(yo[y] + 1, 0), src.rows - 1);
        const uint8_t* S0 = src.row(sy0);
        const uint8_t* S1 = src.row(sy1);
        for (int x = 0; x < dcols; ++x) {
            int s = xo[x];
            int s1 = std::min(s + 1, src.cols - 1);
            r0[x] = S0[s] * a0[x] + S0[s1] * a1[x];
            r1[x] = S1[s] * a0[x] + S1[s1] * a1[x];
        }
        uint8_t* D = dst + (size_t)y * dstep;
        for (int x = 0; x < dcols; ++x)
            D[x] = (uint8_t)((((b0[y] * (r0[x] >> 4)) >> 16) + ((b1[y] * (r1[x] >> 4)) >> 16) + 2) >> 2);
    }
}

// ---------------------------------------------------------------- GaussianBlur
// cv::GaussianBlur 7x7 sigma 2 REFLECT_101 on 8UC1 (A.3); reference ORBextractor.cc:479.
static inline int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0) p = -p;
        else p = 2 * n - 2 - p;
    }
    return p;
}

void gaussian_blur7_u8(const Image& src, uint8_t* dst, int dstep) {
    static const int w[7] = {18, 34, 48, 56, 48, 34, 18};
    const int R = src.rows, C = src.cols;
    // horizontal pass: exact 16-bit sums (max 255*256), no intermediate rounding
    std::vector<uint16_t> h((size_t)R * C);
    std::vector<uint8_t> padded(C + 6);
    for (int y = 0; y < R; ++y) {
        const uint8_t* S = src.row(y);
        for (int x = -3; x < C + 3; ++x) padded[x + 3] = S[reflect101(x, C)];
        uint16_t* H = &h[(size_t)y * C];
        const uint8_t* P = padded.data();
        for (int x = 0; x < C; ++x)
            H[x] = (uint16_t)(w[0] * (P[x] + P[x + 6]) + w[1] * (P[x + 1] + P[x + 5]) +
                              w[2] * (P[x + 2] + P[x + 4]) + w[3] * P[x + 3]);
    }
    const bool inplace = (dst == src.data);
    std::vector<uint8_t> tmp;
    if (inplace) tmp.resize((size_t)R * C);
    for (int y = 0; y < R; ++y) {
        const uint16_t* hr[7];
        for (int j = 0; j < 7; ++j) hr[j] = &h[(size_t)reflect101(y + j - 3, R) * C];
        uint8_t* D = inplace ? &tmp[(size_t)y * C] : dst + (size_t)y * dstep;
        for (int x = 0; x < C; ++x) {
            int acc = w[0] * (hr[0][x] + hr[6][x]) + w[1] * (hr[1][x] + hr[5][x]) +
                      w[2] * (hr[2][x] + hr[4][x]) + w[3] * hr[3][x];
            D[x] = (uint8_t)((acc + 32768) >> 16);
        }
    }
    if (inplace)
        for (int y = 0; y < R; ++y) memcpy(dst + (size_t)y * dstep, &tmp[(size_t)y * C], C);
}

// ---------------------------------------------------------------- FAST
// cv::FAST(img, kps, t, true) TYPE_9_16 (A.1); reference ORBextractor.cc:295,330.
static const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

int fast_score_at(const Image& img, int x, int y, int threshold) {
    const int c = img.row(y)[x];
    int r[16];
    // necessary condition first (every 9-arc holds one pixel of each opposite pair):
    // cheap rejection, identical results
    {
        const int hi = c + threshold, lo = c - threshold;
        bool bright = true, dark = true;
        static const int order[8] = {0, 4, 2, 6, 1, 3, 5, 7};
        for (int q = 0; q < 8; ++q) {
            const int k = order[q];
            const int p0 = img.row(y + kRingDy[k])[x + kRingDx[k]];
            const int p1 = img.row(y + kRingDy[k + 8])[x + kRingDx[k + 8]];
            r[k] = p0;
            r[k + 8] = p1;
            bright = bright && (p0 > hi || p1 > hi);
            dark = dark && (p0 < lo || p1 < lo);
            if (!bright && !dark) return 0;
        }
    }
    // min and max of d over the 16 circular 9-arcs, by window doubling (2,4,8,+1)
    int lo[16], hi[16], lo2[16], hi2[16];
    for (int k = 0; k < 16; ++k) lo[k] = hi[k] = c - r[k];
    for (int k = 0; k < 16; ++k) {
        lo2[k] = std::min(lo[k], lo[(k + 1) & 15]);
        hi2[k] = std::max(hi[k], hi[(k + 1) & 15]);
    }
    int lo4[16], hi4[16];
    for (int k = 0; k < 16; ++k) {
        lo4[k] = std::min(lo2[k], lo2[(k + 2) & 15]);
        hi4[k] = std::max(hi2[k], hi2[(k + 2) & 15]);
    }
    int m = INT_MIN;
    for (int k = 0; k < 16; ++k) {
        int l9 = std::min(std::min(lo4[k], lo4[(k + 4) & 15]), lo[(k + 8) & 15]);
        int h9 = std::max(std::max(hi4[k], hi4[(k + 4) & 15]), hi[(k + 8) & 15]);
        m = std::max(m, std::max(l9, -h9));
    }
    return m > threshold ? m - 1 : 0;
}

void fast9_16_nms(const Image& img, int threshold, std::vector<FastPoint>& out) {
    const int R = img.rows, C = img.cols;
    if (R < 7 || C < 7) return;
    std::vector<int> score((size_t)R * C, 0);
    for (int y = 3; y < R - 3; ++y)
        for (int x = 3; x < C - 3; ++x) score[(size_t)y * C + x] = fast_score_at(img, x, y, threshold);
    for (int y = 3; y < R - 3; ++y)
        for (int x = 3; x < C - 3; ++x) {
            int s = score[(size_t)y * C + x];
            if (s == 0) continue;  // a corner always scores >= threshold >= 1
            const int* p = &score[(size_t)y * C + x];
            if (s > p[-1] && s > p[1] && s > p[-C - 1] && s > p[-C] && s > p[-C + 1] && s > p[C - 1] &&
                s > p[C] && s > p[C + 1])
                out.push_back({x, y, s});
        }
}

// ---------------------------------------------------------------- ctor tables
// ORBextractor::ORBextractor, reference ORBextractor.cc:116-170.
static const int8_t kPairs[728] = {
#include "brief_pairs_182.inc"
};

Tables make_tables(const Params& p) {
    Tables t;
    const int n = p.nlevels;
    const double scaleFactorD = (double)p.scaleFactor;  // the member is a double (ORBextractor.h:79)
    t.scale.assign(n, 1.0f);
    // std::partial_sum(begin, end-1, begin+1, a*scaleFactor) reading and writing the
    // same vector shifted by one (:120-124): {1, 1, s, s^2, ...}.
    if (n >= 2) {
        float acc = t.scale[0];
        t.scale[1] = acc;
        for (int i = 1; i <= n - 2; ++i) {
            acc = (float)((double)acc * scaleFactorD);
            t.scale[i + 1] = acc;
        }
    }
    t.sigma2.resize(n);
    t.inv_scale.resize(n);
    t.inv_sigma2.resize(n);
    for (int i = 0; i < n; ++i) t.sigma2[i] = t.scale[i] * t.scale[i];
    for (int i = 0; i < n; ++i) t.inv_scale[i] = 1.0f / t.scale[i];
    for (int i = 0; i < n; ++i) t.inv_sigma2[i] = 1.0f / t.sigma2[i];

    t.features_per_level.assign(n, 0);
    float factor = (float)(1.0 / scaleFactorD);  // 1.0f / double member (:141)
    // nfeatures*(1-factor) in float, pow(float,int) promotes to double (:142)
    float desired = (float)((double)((float)p.nfeatures * (1 - factor)) /
                            (1.0 - std::pow((double)factor, (double)n)));
    int sum = 0;
    for (int i = 0; i < n - 1; ++i) {
        int cur = cv_round_f(desired);
        sum += cur;
        desired *= factor;
        t.features_per_level[i] = cur;
    }
    t.features_per_level[n - 1] = std::max(p.nfeatures - sum, 0);

    t.pattern.assign(1024, 0);
    for (int i = 0; i < 728; ++i) t.pattern[i] = kPairs[i];

    // umax (:155-169)
    t.umax.assign(16, 0);
    const float half_diag = 15 * std::sqrt(2.f) / 2;
    int vmax = (int)std::floor(half_diag + 1);
    int vmin = (int)std::ceil(half_diag);
    const double hp2 = 15 * 15;
    for (int v = 0; v <= vmax; ++v) t.umax[v] = cv_round_d(std::sqrt(hp2 - v * v));
    for (int v = 15, v0 = 0; v >= vmin; --v) {
        while (t.umax[v0] == t.umax[v0 + 1]) ++v0;
        t.umax[v] = v0;
        ++v0;
    }
    return t;
}

// ---------------------------------------------------------------- octree
// ExtractorNode::DivideNode + ORBextractor::DistributeOctTree,
// reference ORBextractor.cc:178-225, :228-286.
namespace {
struct Node {
    int ulx, uly, urx, bly;  // UL=(ulx,uly) UR=(urx,uly) BL=(ulx,bly) BR=(urx,bly)
    std::vector<int> keys;   // indices into the candidate array, in insertion order
    bool no_more = false;
};
}  // namespace

bool distribute_octree(const std::vector<Cand>& keys, int minX, int maxX, int minY, int maxY, int N,
                       std::vector<int>& out_idx) {
    out_idx.clear();
    const int nIni = (maxX - minX) / (maxY - minY);  // integer division (D3)
    const float hX = static_cast<float>(maxX - minX) / nIni;
    std::list<Node> nodes;
    std::vector<Node*> roots(nIni > 0 ? nIni : 0);
    for (int i = 0; i < nIni; ++i) {
        Node n;
        n.ulx = (int)(hX * i);
        n.urx = (int)(hX * (i + 1));
        n.uly = 0;
        n.bly = maxY - minY;
        nodes.push_back(n);
        roots[i] = &nodes.back();
    }
    for (int k = 0; k < (int)keys.size(); ++k) {
        int idx = (int)(keys[k].x / hX);
        if (idx >= 0 && idx < nIni) roots[idx]->keys.push_back(k);
    }
    bool finish = false;
    int guard = 0;
    while (!finish) {
        if (++guard > 64) return false;  // the reference would loop forever (unseparable keys)
        for (auto it = nodes.begin(); it != nodes.end();) {
            if (it->keys.size() == 1) {
                it->no_more = true;
                ++it;
            } else if (it->keys.empty()) {
                it = nodes.erase(it);
            } else {
                const int halfX = (it->urx - it->ulx) / 2;  // floor (D4)
                const int halfY = (it->bly - it->uly) / 2;
                const int midx = it->ulx + halfX, midy = it->uly + halfY;
                Node c[4];
                c[0].ulx = it->ulx; c[0].urx = midx;    c[0].uly = it->uly; c[0].bly = midy;
                c[1].ulx = midx;    c[1].urx = it->urx; c[1].uly = it->uly; c[1].bly = midy;
                c[2].ulx = it->ulx; c[2].urx = midx;    c[2].uly = midy;    c[2].bly = it->bly;
                c[3].ulx = midx;    c[3].urx = it->urx; c[3].uly = midy;    c[3].bly = it->bly;
                for (int k : it->keys) {
                    const Cand& kp = keys[k];
                    if (kp.x < (float)midx) {
                        if (kp.y < (float)midy) c[0].keys.push_back(k);
                        else c[2].keys.push_back(k);
                    } else {
                        if (kp.y < (float)midy) c[1].keys.push_back(k);
                        else c[3].keys.push_back(k);
                    }
                }
                for (int q = 0; q < 4; ++q) {
                    c[q].no_more = c[q].keys.size() == 1;
                    if (!c[q].keys.empty()) nodes.push_front(c[q]);
                }
                it = nodes.erase(it);
            }
        }
        bool all_done = true;
        for (const Node& n : nodes) all_done = all_done && n.no_more;
        finish = ((long long)nodes.size() >= (long long)N) || all_done;
    }
    for (const Node& n : nodes) {
        if (n.keys.empty()) continue;
        int best = n.keys[0];  // std::max_element: first of the maxima
        for (int k : n.keys)
            if (keys[best].response < keys[k].response) best = k;
        out_idx.push_back(best);
    }
    return true;
}

// ---------------------------------------------------------------- orientation
// IC_Angle, reference ORBextractor.cc:21-48.
static float ic_angle(const Image& img, int cx, int cy, const std::vector<int>& umax) {
    int m01 = 0, m10 = 0;
    const uint8_t* center = img.row(cy) + cx;
    for (int u = -15; u <= 15; ++u) m10 += u * center[u];
    const int step = img.step;
    for (int v = 1; v <= 15; ++v) {
        int vsum = 0;
        const int d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int below = center[u + v * step], above = center[u - v * step];
            vsum += (below - above);
            m10 += u * (below + above);
        }
        m01 += v * vsum;
    }
    return fast_atan2((float)m01, (float)m10);
}

// ---------------------------------------------------------------- descriptor
// computeOrbDescriptor / getRotatedValue, reference ORBextractor.cc:53-73.
static void describe(const Image& blurred, int cx, int cy, float angle_deg, const int8_t* pattern,
                     uint8_t* desc) {
    const float factorPI = (float)(3.141592653589793238462643383279502884 / 180.f);
    const float angle = angle_deg * factorPI;
    const float a = cosf(angle), b = sinf(angle);
    const uint8_t* center = blurred.row(cy) + cx;
    const int step = blurred.step;
    auto sample = [&](int idx) -> int {
        const float px = (float)pattern[2 * idx], py = (float)pattern[2 * idx + 1];
        const int yy = cv_round_f(px * b + py * a);
        const int xx = cv_round_f(px * a - py * b);
        return center[yy * step + xx];
    };
    for (int i = 0; i < 32; ++i) {
        int val = 0;
        for (int j = 0; j < 8; ++j) {
            int t0 = sample(16 * i + 2 * j), t1 = sample(16 * i + 2 * j + 1);
            val |= (t0 < t1) << j;
        }
        desc[i] = (uint8_t)val;
    }
}

// ---------------------------------------------------------------- operator()
// ORBextractor::operator(), ComputePyramid, ComputeKeyPointsOctTree;
// reference ORBextractor.cc:442-495, :497-515, :288-357.
int extract(const Params& p, const Image& img, ExtractResult& out, bool keep_pyramid) {
    out = ExtractResult();
    if (img.rows <= 0 || img.cols <= 0 || !img.data) return 0;  // empty image: silent return (:444)
    if (p.nlevels < 1) return -1;
    const Tables t = make_tables(p);
    const int L = p.nlevels;

    // ComputePyramid (:497-515).  The 19-px reflected border the reference adds
    // around every level is never read on this path, so levels are tight.
    std::vector<std::vector<uint8_t>> pyr(L);
    std::vector<int> lr(L), lc(L);
    for (int l = 0; l < L; ++l) {
        const float s = t.inv_scale[l];
        lc[l] = cv_round_f(img.cols * s);
        lr[l] = cv_round_f(img.rows * s);
        if (lc[l] <= 0 || lr[l] <= 0) return -1;
        pyr[l].resize((size_t)lr[l] * lc[l]);
        if (l == 0) {
            for (int y = 0; y < lr[0]; ++y) memcpy(&pyr[0][(size_t)y * lc[0]], img.row(y), lc[0]);
        } else {
            Image prev{pyr[l - 1].data(), lr[l - 1], lc[l - 1], lc[l - 1]};
            resize_linear_u8(prev, pyr[l].data(), lr[l], lc[l], lc[l]);
        }
    }

    // ComputeKeyPointsOctTree (:288-357)
    std::vector<std::vector<KeyPoint>> all(L);
    out.n_candidates.assign(L, 0);
    out.n_keypoints.assign(L, 0);
    for (int l = 0; l < L; ++l) {
        const Image lev{pyr[l].data(), lr[l], lc[l], lc[l]};
        const int minBX = 19 - 3, minBY = minBX;
        const int maxBX = lc[l] - 19 + 3, maxBY = lr[l] - 19 + 3;
        const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
        const int nCols = (int)(width / 30.f), nRows = (int)(height / 30.f);
        // levels smaller than one 30-px cell: the reference's cell loops simply do not run
        // (the float division by zero only feeds the unused wCell/hCell); a non-positive
        // height is where it really faults (integer division by zero / negative vector size)
        // (:230: H == 0 -> SIGFPE; nIni = W/H < 0 -> std::vector(nIni) throws)
        if (maxBY - minBY == 0 || (maxBX - minBX) / (maxBY - minBY) < 0) return -1;
        const bool has_cells = nCols > 0 && nRows > 0;
        const int wCell = has_cells ? (int)std::ceil(width / nCols) : 0;
        const int hCell = has_cells ? (int)std::ceil(height / nRows) : 0;
        std::vector<Cand> cands;
        std::vector<FastPoint> cell;
        for (int i = 0; has_cells && i < nRows; ++i) {
            const int iniY = minBY + i * hCell;
            const int maxY = std::min(iniY + hCell + 6, maxBY);
            for (int j = 0; j < nCols; ++j) {
                const int iniX = minBX + j * wCell;
                const int maxX = std::min(iniX + wCell + 6, maxBX);
                // negative ROI extent: cv::Mat::operator()(Rect) throws in the reference
                if (maxX - iniX < 0 || maxY - iniY < 0) return -2;
                Image roi{lev.row(iniY) + iniX, maxY - iniY, maxX - iniX, lev.step};
                cell.clear();
                fast9_16_nms(roi, p.iniThFAST, cell);
                if (cell.empty()) {  // retry only if the post-NMS list is empty (:293-296)
                    fast9_16_nms(roi, p.minThFAST, cell);
                    if (roi.rows >= 7 && roi.cols >= 7) out.n_retry_cells++;
                }
                for (const FastPoint& fp : cell)
                    cands.push_back({(float)(fp.x + j * wCell), (float)(fp.y + i * hCell), (float)fp.score});
            }
        }
        out.n_candidates[l] = (int)cands.size();
        std::vector<int> keep;
        if (!distribute_octree(cands, minBX, maxBX, minBY, maxBY, t.features_per_level[l], keep))
            return -3;
        const int patch = (int)(31 * t.scale[l]);
        for (int k : keep) {
            KeyPoint kp;
            kp.x = cands[k].x + minBX;
            kp.y = cands[k].y + minBY;
            kp.size = (float)patch;
            kp.angle = -1.f;
            kp.response = cands[k].response;
            kp.octave = l;
            kp.class_id = -1;
            all[l].push_back(kp);
        }
        out.n_keypoints[l] = (int)keep.size();
    }
    for (int l = 0; l < L; ++l) {
        const Image lev{pyr[l].data(), lr[l], lc[l], lc[l]};
        for (KeyPoint& kp : all[l]) kp.angle = ic_angle(lev, cv_round_f(kp.x), cv_round_f(kp.y), t.umax);
    }

    size_t total = 0;
    for (auto& v : all) total += v.size();
    if (keep_pyramid) {
        out.pyramid = pyr;
        out.level_rows = lr;
        out.level_cols = lc;
    }
    if (total == 0) return 0;  // descriptors.release(), keypoints untouched (:460-463)
    out.descriptors.assign(total * 32, 0);
    out.keypoints.reserve(total);
    size_t off = 0;
    std::vector<uint8_t> blurred;
    for (int l = 0; l < L; ++l) {
        if (all[l].empty()) continue;
        const Image lev{pyr[l].data(), lr[l], lc[l], lc[l]};
        blurred.resize(pyr[l].size());
        gaussian_blur7_u8(lev, blurred.data(), lc[l]);
        const Image bl{blurred.data(), lr[l], lc[l], lc[l]};
        for (KeyPoint& kp : all[l]) {
            describe(bl, cv_round_f(kp.x), cv_round_f(kp.y), kp.angle, t.pattern.data(),
                     &out.descriptors[off * 32]);
            ++off;
        }
        if (l != 0) {
            const float s = t.scale[l];
            for (KeyPoint& kp : all[l]) {
                kp.x *= s;
                kp.y *= s;
            }
        }
        out.keypoints.insert(out.keypoints.end(), all[l].begin(), all[l].end());
    }
    return 0;
}

// ---------------------------------------------------------------- matcher
// ORBmatcher::DescriptorDistance, reference ORBmatcher.cc:896-908.
int descriptor_distance(const uint8_t* a, const uint8_t* b) {
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t x, y;
        memcpy(&x, a + 4 * i, 4);
        memcpy(&y, b + 4 * i, 4);
        uint32_t v = x ^ y;
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (int)((((v + (v >> 4)) & 0xF0F0F0Fu) * 0x1010101u) >> 24);
    }
    return dist;
}

// ---------------------------------------------------------------- synthetic frames (A.8)
static inline uint64_t splitmix64(uint64_t k) {
    uint64_t z = k + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void synth_frame(uint8_t* dst, int rows, int cols, int step, uint64_t seed, uint64_t frame,
                 int variant, int right) {
    static const int blocksA[4] = {5, 11, 23, 47};
    static const int blocksB[4] = {4, 8, 16, 32};
    static const double wts[4] = {0.4, 0.3, 0.2, 0.1};
    const int* bs = variant == 1 ? blocksB : blocksA;
    const uint64_t base = (seed << 40) ^ (frame << 24);
    for (int y = 0; y < rows; ++y) {
        for (int x = 0; x < cols; ++x) {
            int sx = x;
            if (right) {  // per-block disparity 0..40 px: right(x,y) = left(x + d, y)
                uint64_t dk = base ^ (7ull << 60) ^ ((uint64_t)(y / 47) << 12) ^ (uint64_t)(x / 94);
                sx = x + (int)((splitmix64(dk) & 0xFF) % 41);
            }
            double acc = 0.0;
            for (int o = 0; o < 4; ++o) {
                uint64_t k = base ^ ((uint64_t)o << 60) ^ ((uint64_t)(y / bs[o]) << 12) ^ (uint64_t)(sx / bs[o]);
                acc += wts[o] * (double)(splitmix64(k) & 0xFF);
            }
            if (variant == 1 && x < cols / 2) acc = 128.0 + (acc - 128.0) * 0.25;
            double v = std::floor(acc + 0.5);
            dst[(size_t)y * step + x] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    }
}

}  // namespace orb_oracle

// ============================================================ C ABI (ctypes)
using namespace orb_oracle;

extern "C" {

int orc_extract(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh,
                const uint8_t* img, int rows, int cols, int step, KeyPoint* kps, uint8_t* desc,
                int cap, int* count, int* per_level_candidates, int* per_level_kept,
                int* n_retry_cells) {
    Params p{nfeatures, scaleFactor, nlevels, iniTh, minTh};
    ExtractResult r;
    int rc = extract(p, Image{img, rows, cols, step}, r, false);
    if (rc != 0) return rc;
    int n = (int)r.keypoints.size();
    if (count) *count = n;
    if (n > cap) return -10;
    if (n) {
        memcpy(kps, r.keypoints.data(), (size_t)n * sizeof(KeyPoint));
        memcpy(desc, r.descriptors.data(), (size_t)n * 32);
    }
    for (int l = 0; l < nlevels; ++l) {
        if (per_level_candidates) per_level_candidates[l] = l < (int)r.n_candidates.size() ? r.n_candidates[l] : 0;
        if (per_level_kept) per_level_kept[l] = l < (int)r.n_keypoints.size() ? r.n_keypoints[l] : 0;
    }
    if (n_retry_cells) *n_retry_cells = r.n_retry_cells;
    return 0;
}

int orc_tables(int nfeatures, float scaleFactor, int nlevels, float* scale, float* inv_scale,
               float* sigma2, float* inv_sigma2, int* nfeat_per_level, int* umax16) {
    Tables t = make_tables(Params{nfeatures, scaleFactor, nlevels, 20, 7});
    for (int i = 0; i < nlevels; ++i) {
        scale[i] = t.scale[i];
        inv_scale[i] = t.inv_scale[i];
        sigma2[i] = t.sigma2[i];
        inv_sigma2[i] = t.inv_sigma2[i];
        nfeat_per_level[i] = t.features_per_level[i];
    }
    for (int i = 0; i < 16; ++i) umax16[i] = t.umax[i];
    return 0;
}

int orc_pyramid(float scaleFactor, int nlevels, const uint8_t* img, int rows, int cols, int step,
                int level, uint8_t* dst, int dst_cap, int* lrows, int* lcols) {
    Tables t = make_tables(Params{1000, scaleFactor, nlevels, 20, 7});
    std::vector<uint8_t> prev((size_t)rows * cols), cur;
    for (int y = 0; y < rows; ++y) memcpy(&prev[(size_t)y * cols], img + (size_t)y * step, cols);
    int pr = rows, pc = cols;
    for (int l = 1; l <= level; ++l) {
        int c = cv_round_f(cols * t.inv_scale[l]), r = cv_round_f(rows * t.inv_scale[l]);
        cur.assign((size_t)r * c, 0);
        resize_linear_u8(Image{prev.data(), pr, pc, pc}, cur.data(), r, c, c);
        prev.swap(cur);
        pr = r;
        pc = c;
    }
    *lrows = pr;
    *lcols = pc;
    if ((size_t)pr * pc > (size_t)dst_cap) return -10;
    memcpy(dst, prev.data(), (size_t)pr * pc);
    return 0;
}

void orc_resize(const uint8_t* src, int srows, int scols, int sstep, uint8_t* dst, int drows,
                int dcols, int dstep) {
    resize_linear_u8(Image{src, srows, scols, sstep}, dst, drows, dcols, dstep);
}

void orc_blur7(const uint8_t* src, int rows, int cols, int step, uint8_t* dst, int dstep) {
    gaussian_blur7_u8(Image{src, rows, cols, step}, dst, dstep);
}

int orc_fast(const uint8_t* img, int rows, int cols, int step, int threshold, int* xys, int cap) {
    std::vector<FastPoint> v;
    fast9_16_nms(Image{img, rows, cols, step}, threshold, v);
    int n = (int)v.size();
    for (int i = 0; i < n && i < cap; ++i) {
        xys[3 * i] = v[i].x;
        xys[3 * i + 1] = v[i].y;
        xys[3 * i + 2] = v[i].score;
    }
    return n;
}

float orc_fast_atan2(float y, float x) { return fast_atan2(y, x); }

int orc_octree(const float* xyr, int n, int minX, int maxX, int minY, int maxY, int N, int* out_idx,
               int cap) {
    std::vector<Cand> keys(n);
    for (int i = 0; i < n; ++i) keys[i] = Cand{xyr[3 * i], xyr[3 * i + 1], xyr[3 * i + 2]};
    std::vector<int> out;
    if (!distribute_octree(keys, minX, maxX, minY, maxY, N, out)) return -3;
    int k = (int)out.size();
    for (int i = 0; i < k && i < cap; ++i) out_idx[i] = out[i];
    return k;
}

float orc_ic_angle(const uint8_t* img, int rows, int cols, int step, int x, int y) {
    Tables t = make_tables(Params{1000, 1.2f, 8, 20, 7});
    return ic_angle(Image{img, rows, cols, step}, x, y, t.umax);
}

void orc_describe(const uint8_t* blurred, int rows, int cols, int step, int x, int y, float angle,
                  uint8_t* desc32) {
    Tables t = make_tables(Params{1000, 1.2f, 8, 20, 7});
    describe(Image{blurred, rows, cols, step}, x, y, angle, t.pattern.data(), desc32);
}

int orc_distance(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }

// The scan every ORBmatcher search shares (reference ORBmatcher.cc:49-55,
// :225-231, :321-327): candidates in list order, strict '<' updates.
void orc_match_all(const uint8_t* q, int nq, const uint8_t* t, int nt, int* best_idx,
                   int* best_dist, int* second_dist) {
    for (int i = 0; i < nq; ++i) {
        int best = INT_MAX, second = INT_MAX, idx = -1;
        for (int j = 0; j < nt; ++j) {
            int d = descriptor_distance(q + 32 * (size_t)i, t + 32 * (size_t)j);
            if (d < best) {
                second = best;
                best = d;
                idx = j;
            } else if (d < second) {
                second = d;
            }
        }
        best_idx[i] = idx;
        best_dist[i] = best;
        second_dist[i] = second;
    }
}

// Windowed variant: query i scans cand[offsets[i]..offsets[i+1]).  tie_last=0:
// the scan above.  tie_last=1: SearchForTriangulation's rule (reference
// ORBmatcher.cc:404-419): start at max_dist, accept d <= max_dist && d <= best
// (ties -> last candidate); second is not tracked (INT_MAX).  With tie_last=0
// max_dist is ignored (thresholds are applied by the caller).
void orc_match_csr(const uint8_t* q, int nq, const uint8_t* t, int nt, const int* offsets,
                   const int* cand, int tie_last, int max_dist, int* best_idx, int* best_dist,
                   int* second_dist) {
    (void)nt;
    for (int i = 0; i < nq; ++i) {
        int best = tie_last ? max_dist : INT_MAX, second = INT_MAX, idx = -1;
        for (int c = offsets[i]; c < offsets[i + 1]; ++c) {
            int j = cand[c];
            int d = descriptor_distance(q + 32 * (size_t)i, t + 32 * (size_t)j);
            if (tie_last) {
                if (d > max_dist || d > best) continue;
                best = d;
                idx = j;
            } else if (d < best) {
                second = best;
                best = d;
                idx = j;
            } else if (d < second) {
                second = d;
            }
        }
        best_idx[i] = idx;
        best_dist[i] = best;
        second_dist[i] = second;
    }
}

// Hamming part of Frame::ComputeStereoMatches, reference Frame.cc:446-529.
// best_r[i] = right index of the best candidate (or -1 when the scan never
// improved on TH_HIGH=100), best_dist[i] = its distance (100 when none).
// Returns 0, or -1 if a right keypoint's row band leaves the image (the
// reference indexes vRowIndices out of bounds there).
int orc_stereo_match(const KeyPoint* kl, const uint8_t* dl, int nl, const KeyPoint* kr,
                     const uint8_t* dr, int nr, const float* scale, int nlevels, int rows, float bf,
                     float fx, int* best_r, int* best_dist) {
    (void)nlevels;
    std::vector<std::vector<int>> rowidx(rows);
    for (int iR = 0; iR < nr; ++iR) {
        const float y = kr[iR].y;
        const float r = 2.0f * scale[kr[iR].octave];
        const int maxr = (int)std::ceil(y + r), minr = (int)std::floor(y - r);
        if (minr < 0 || maxr >= rows) return -1;
        for (int yi = minr; yi <= maxr; ++yi) rowidx[yi].push_back(iR);
    }
    const float mb = bf / fx;  // Frame.cc:215
    const float minD = 0, maxD = bf / mb;
    for (int iL = 0; iL < nl; ++iL) {
        best_r[iL] = -1;
        best_dist[iL] = 100;
        const int levelL = kl[iL].octave;
        const float vL = kl[iL].y, uL = kl[iL].x;
        const std::vector<int>& c = rowidx[(size_t)vL];
        if (c.empty()) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int best = 100, bi = -1;
        for (int iR : c) {
            if (kr[iR].octave < levelL - 1 || kr[iR].octave > levelL + 1) continue;
            const float uR = kr[iR].x;
            if (uR >= minU && uR <= maxU) {
                int d = descriptor_distance(dl + 32 * (size_t)iL, dr + 32 * (size_t)iR);
                if (d < best) {
                    best = d;
                    bi = iR;
                }
            }
        }
        best_r[iL] = bi;
        best_dist[iL] = best;
    }
    return 0;
}

// Frame::ComputeStereoMatches whole, reference Frame.cc:446-619: the Hamming search above, then the
// 11x11 SAD sliding window on the pyramid level of the left keypoint (:531-575), the parabola fit
// (:577-586), the disparity gate (:588-603) and the median*2.1 outlier cut (:606-619).
// The pyramids are ComputePyramid's (ORBextractor.cc:497-515) of the two level-0 images.
// mvuRight / mvDepth get nl floats.  Returns 0; -1 if a right keypoint's row band leaves the image
// (out-of-bounds vRowIndices); -2 if a rowRange / colRange leaves its level image (cv::Exception in
// the reference).  An empty vDistIdx (the reference then reads vDistIdx[0] of an empty vector) is
// treated as "nothing to cut".
int orc_compute_stereo_matches(float scaleFactor, int nlevels, const uint8_t* imgL, const uint8_t* imgR, int rows, int cols, int step,
                               const KeyPoint* kl, const uint8_t* dl, int nl, const KeyPoint* kr, const uint8_t* dr, int nr, float mbf,
                               float fx, float* mvuRight, float* mvDepth) {
    const int TH_HIGH = 100, TH_LOW = 50;
    Tables t = make_tables(Params{1000, scaleFactor, nlevels, 20, 7});
    const std::vector<float>& mvScaleFactors = t.scale;
    const std::vector<float>& mvInvScaleFactors = t.inv_scale;
    // both pyramids
    std::vector<std::vector<uint8_t>> pyr[2];
    std::vector<int> lr(nlevels), lc(nlevels);
    for (int side = 0; side < 2; ++side) {
        const uint8_t* img = side ? imgR : imgL;
        pyr[side].resize(nlevels);
        for (int l = 0; l < nlevels; ++l) {
            lc[l] = l ? cv_round_f(cols * t.inv_scale[l]) : cols;
            lr[l] = l ? cv_round_f(rows * t.inv_scale[l]) : rows;
            pyr[side][l].assign((size_t)lr[l] * lc[l], 0);
            if (l == 0)
                for (int y = 0; y < rows; ++y) memcpy(&pyr[side][0][(size_t)y * cols], img + (size_t)y * step, cols);
            else
                resize_linear_u8(Image{pyr[side][l - 1].data(), lr[l - 1], lc[l - 1], lc[l - 1]}, pyr[side][l].data(), lr[l], lc[l], lc[l]);
        }
    }
    const int N = nl;
    for (int i = 0; i < N; ++i) {
        mvuRight[i] = -1.0f;
        mvDepth[i] = -1.0f;
    }
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
    const int nRows = rows;
    std::vector<std::vector<size_t>> vRowIndices(nRows);
    for (int iR = 0; iR < nr; iR++) {
        const float kpY = kr[iR].y;
        const float r = 2.0f * mvScaleFactors[kr[iR].octave];
        const int maxr = ceil(kpY + r);
        const int minr = floor(kpY - r);
        if (minr < 0 || maxr >= nRows) return -1;
        for (int yi = minr; yi <= maxr; yi++) vRowIndices[yi].push_back(iR);
    }
    const float mb = mbf / fx;
    const float minZ = mb;
    const float minD = 0;
    const float maxD = mbf / minZ;
    std::vector<std::pair<int, int>> vDistIdx;
    for (int iL = 0; iL < N; iL++) {
        const KeyPoint& kpL = kl[iL];
        const int levelL = kpL.octave;
        const float vL = kpL.y;
        const float uL = kpL.x;
        const std::vector<size_t>& vCandidates = vRowIndices[(size_t)vL];
        if (vCandidates.empty()) continue;
        const float minU = uL - maxD;
        const float maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = TH_HIGH;
        size_t bestIdxR = 0;
        for (size_t iC = 0; iC < vCandidates.size(); iC++) {
            const size_t iR = vCandidates[iC];
            const KeyPoint& kpR = kr[iR];
            if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
            const float uR = kpR.x;
            if (uR >= minU && uR <= maxU) {
                const int dist = descriptor_distance(dl + 32 * (size_t)iL, dr + 32 * iR);
                if (dist < bestDist) {
                    bestDist = dist;
                    bestIdxR = iR;
                }
            }
        }
        if (bestDist < thOrbDist) {
            const float uR0 = kr[bestIdxR].x;
            const float scaleFactorL = mvInvScaleFactors[kpL.octave];
            const float scaleduL = round(kpL.x * scaleFactorL);
            const float scaledvL = round(kpL.y * scaleFactorL);
            const float scaleduR0 = round(uR0 * scaleFactorL);
            const int w = 5;
            const int R = lr[kpL.octave], C = lc[kpL.octave];
            const uint8_t* PL = pyr[0][kpL.octave].data();
            const uint8_t* PR = pyr[1][kpL.octave].data();
            // IL = left.rowRange(scaledvL-w, scaledvL+w+1).colRange(scaleduL-w, scaleduL+w+1)
            const int r0 = (int)(scaledvL - w), r1 = (int)(scaledvL + w + 1);
            const int c0 = (int)(scaleduL - w), c1 = (int)(scaleduL + w + 1);
            if (r0 < 0 || r1 > R || c0 < 0 || c1 > C) return -2;
            float IL[11][11];
            for (int a = 0; a < 11; ++a)
                for (int b = 0; b < 11; ++b) IL[a][b] = (float)PL[(size_t)(r0 + a) * C + c0 + b];
            const float cL = IL[w][w];
            for (int a = 0; a < 11; ++a)
                for (int b = 0; b < 11; ++b) IL[a][b] = IL[a][b] - cL * 1.0f;
            int bestDistS = INT_MAX;
            int bestincR = 0;
            const int L = 5;
            std::vector<float> vDists(2 * L + 1);
            const float iniu = scaleduR0 + L - w;
            const float endu = scaleduR0 + L + w + 1;
            if (iniu < 0 || endu >= C) continue;
            for (int incR = -L; incR <= +L; incR++) {
                const int q0 = (int)(scaleduR0 + incR - w), q1 = (int)(scaleduR0 + incR + w + 1);
                if (q0 < 0 || q1 > C) return -2;
                float IR[11][11];
                for (int a = 0; a < 11; ++a)
                    for (int b = 0; b < 11; ++b) IR[a][b] = (float)PR[(size_t)(r0 + a) * C + q0 + b];
                const float cR = IR[w][w];
                double acc = 0;  // cv::norm(IL, IR, NORM_L1): sum of |a - b| in double
                for (int a = 0; a < 11; ++a)
                    for (int b = 0; b < 11; ++b) acc += std::fabs((double)(IL[a][b] - (IR[a][b] - cR * 1.0f)));
                float dist = (float)acc;
                if (dist < bestDistS) {
                    bestDistS = dist;
                    bestincR = incR;
                }
                vDists[L + incR] = dist;
            }
            if (bestincR == -L || bestincR == L) continue;
            const float dist1 = vDists[L + bestincR - 1];
            const float dist2 = vDists[L + bestincR];
            const float dist3 = vDists[L + bestincR + 1];
            const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = mvScaleFactors[kpL.octave] * ((float)scaleduR0 + (float)bestincR + deltaR);
            float disparity = (uL - bestuR);
            if (disparity >= minD && disparity < maxD) {
                if (disparity <= 0) {
                    disparity = 0.01;
                    bestuR = uL - 0.01;
                }
                mvDepth[iL] = mbf / disparity;
                mvuRight[iL] = bestuR;
                vDistIdx.push_back(std::pair<int, int>(bestDistS, iL));
            }
        }
    }
    if (vDistIdx.empty()) return 0;
    sort(vDistIdx.begin(), vDistIdx.end());
    const float median = vDistIdx[vDistIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    for (int i = (int)vDistIdx.size() - 1; i >= 0; i--) {
        if (vDistIdx[i].first < thDist)
            break;
        else {
            mvuRight[vDistIdx[i].second] = -1;
            mvDepth[vDistIdx[i].second] = -1;
        }
    }
    return 0;
}

// ---- ORBmatcher::ComputeThreeMaxima, reference ORBmatcher.cc:469-502 ----
static void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
    int topIdx[3] = {-1, -1, -1}, topVal[3] = {0, 0, 0};
    for (int i = 0; i < L; ++i) {
        const int value = (int)histo[i].size();
        for (int j = 0; j < 3; ++j) {
            if (value > topVal[j]) {
                for (int k = 2; k > j; --k) {
                    topVal[k] = topVal[k - 1];
                    topIdx[k] = topIdx[k - 1];
                }
                topVal[j] = value;
                topIdx[j] = i;
                break;
            }
        }
    }
    ind1 = topIdx[0];
    ind2 = topIdx[1];
    ind3 = topIdx[2];
    if (topVal[1] < 0.1f * topVal[0]) {
        ind2 = -1;
        ind3 = -1;
    } else if (topVal[2] < 0.1f * topVal[0]) {
        ind3 = -1;
    }
}

// Feature vectors arrive as sorted node ids with CSR index lists (DBoW2::FeatureVector is a
// std::map<NodeId, std::vector<unsigned>>, so iteration is by ascending node id and the
// lower_bound jumps of the reference are a plain merge of the two id lists).

// ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, ...), reference ORBmatcher.cc:278-366.
// has1/has2: the keypoint holds a good map point.  matches12[i1] = i2 or -1.  Returns nmatches.
int orc_search_by_bow_kf(const uint8_t* d1, const float* ang1, const uint8_t* has1, int n1, const uint8_t* d2, const float* ang2,
                         const uint8_t* has2, int n2, const int* nodes1, const int* off1, const int* idx1, int nn1, const int* nodes2,
                         const int* off2, const int* idx2, int nn2, float nnratio, int checkOri, int* matches12) {
    const int TH_LOW = 50, HISTO = 30;
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    std::vector<bool> matched2(n2, false);
    std::vector<int> rotHist[30];
    const float factor = 1.0f / HISTO;
    int nmatches = 0;
    int a = 0, b = 0;
    while (a < nn1 && b < nn2) {
        if (nodes1[a] == nodes2[b]) {
            for (int ia = off1[a]; ia < off1[a + 1]; ++ia) {
                const int i1 = idx1[ia];
                if (!has1[i1]) continue;
                int best1 = INT_MAX, bestIdx2 = -1, best2 = INT_MAX;
                for (int ib = off2[b]; ib < off2[b + 1]; ++ib) {
                    const int i2 = idx2[ib];
                    if (matched2[i2] || !has2[i2]) continue;
                    const int dist = descriptor_distance(d1 + 32 * (size_t)i1, d2 + 32 * (size_t)i2);
                    if (dist < best1) {
                        best2 = best1;
                        best1 = dist;
                        bestIdx2 = i2;
                    } else if (dist < best2) {
                        best2 = dist;
                    }
                }
                if (best1 < TH_LOW && static_cast<float>(best1) < nnratio * static_cast<float>(best2)) {
                    matches12[i1] = bestIdx2;
                    matched2[bestIdx2] = true;
                    nmatches++;
                    if (checkOri) {
                        float rot = ang1[i1] - ang2[bestIdx2];
                        if (rot < 0.0f) rot += 360.0f;
                        int bin = (int)std::round(rot * factor);
                        if (bin == HISTO) bin = 0;
                        rotHist[bin].push_back(i1);
                    }
                }
            }
            ++a;
            ++b;
        } else if (nodes1[a] < nodes2[b]) {
            ++a;
        } else {
            ++b;
        }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO, ind1, ind2, ind3);
        for (int i = 0; i < HISTO; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int i1 : rotHist[i]) {
                matches12[i1] = -1;
                nmatches--;
            }
        }
    }
    return nmatches;
}

// ORBmatcher::SearchForTriangulation + CheckDistEpipolarLine, reference ORBmatcher.cc:368-467, :71-85
// (bOnlyStereo = false, the only way it is called: LocalMapping.cc:187).  x/y/octave: undistorted
// keypoints; has1/has2: the keypoint already holds a map point; F12 row-major 3x3; sigma2 =
// KF2's mvLevelSigma2.  matches12[i1] = i2 or -1.  Returns nmatches.
int orc_search_for_triangulation(const uint8_t* d1, const float* x1, const float* y1, const float* ang1, const uint8_t* has1, int n1,
                                 const uint8_t* d2, const float* x2, const float* y2, const float* ang2, const int* oct2,
                                 const uint8_t* has2, int n2, const int* nodes1, const int* off1, const int* idx1, int nn1,
                                 const int* nodes2, const int* off2, const int* idx2, int nn2, const float* F12, const float* sigma2,
                                 int checkOri, int* matches12) {
    const int TH_LOW = 50, HISTO = 30;
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    std::vector<int> rotHist[30];
    const float factor = 1.0f / HISTO;
    int nmatches = 0;
    int a = 0, b = 0;
    while (a < nn1 && b < nn2) {
        if (nodes1[a] == nodes2[b]) {
            for (int ia = off1[a]; ia < off1[a + 1]; ++ia) {
                const int i1 = idx1[ia];
                if (has1[i1]) continue;
                int bestDist = TH_LOW, bestIdx2 = -1;
                for (int ib = off2[b]; ib < off2[b + 1]; ++ib) {
                    const int i2 = idx2[ib];
                    if (has2[i2]) continue;  // vbMatched2 is never set in this fork
                    const int dist = descriptor_distance(d1 + 32 * (size_t)i1, d2 + 32 * (size_t)i2);
                    if (dist > TH_LOW || dist > bestDist) continue;
                    // CheckDistEpipolarLine
                    const float ea = x1[i1] * F12[0] + y1[i1] * F12[3] + F12[6];
                    const float eb = x1[i1] * F12[1] + y1[i1] * F12[4] + F12[7];
                    const float ec = x1[i1] * F12[2] + y1[i1] * F12[5] + F12[8];
                    const float num = ea * x2[i2] + eb * y2[i2] + ec;
                    const float den = ea * ea + eb * eb;
                    bool ok = false;
                    if (den != 0) {
                        const float dsqr = num * num / den;
                        ok = dsqr < 3.84 * sigma2[oct2[i2]];
                    }
                    if (ok) {
                        bestIdx2 = i2;
                        bestDist = dist;
                        if (checkOri) {
                            float rot = ang1[i1] - ang2[i2];
                            if (rot < 0.0) rot += 360.0f;
                            int bin = static_cast<int>(std::round(rot * factor)) % HISTO;
                            rotHist[bin].push_back(i1);
                        }
                    }
                }
                if (bestIdx2 >= 0) {
                    matches12[i1] = bestIdx2;
                    nmatches++;
                }
            }
            ++a;
            ++b;
        } else if (nodes1[a] < nodes2[b]) {
            ++a;
        } else {
            ++b;
        }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO, ind1, ind2, ind3);
        for (int i = 0; i < HISTO; ++i) {
            if (i != ind1 && i != ind2 && i != ind3) {
                for (int i1 : rotHist[i]) {
                    if (matches12[i1] >= 0) {
                        matches12[i1] = -1;
                        nmatches--;
                    }
                }
            }
        }
    }
    return nmatches;
}

// ---- Frame grid and windowed searches ------------------------------------------------------
// The Frame / KeyFrame members the searches read, as plain arrays (KeyPoint = mvKeysUn).
struct OrcFrame {
    const KeyPoint* keysUn;
    const uint8_t* desc;
    int N;
    float mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv;
};

namespace {
const int FRAME_GRID_ROWS = 48, FRAME_GRID_COLS = 64;  // reference include/Frame.h:17-18

struct Grid {
    std::vector<size_t> cell[64][48];
};

// Frame::PosInGrid + Frame::AssignFeaturesToGrid, reference src/Frame.cc:362-372, :210-225.
void assign_features_to_grid(const OrcFrame& F, Grid& g) {
    for (int i = 0; i < F.N; i++) {
        const KeyPoint& kp = F.keysUn[i];
        int posX = round((kp.x - F.mnMinX) * F.mfGridElementWidthInv);
        int posY = round((kp.y - F.mnMinY) * F.mfGridElementHeightInv);
        if (posX < 0 || posX >= FRAME_GRID_COLS || posY < 0 || posY >= FRAME_GRID_ROWS) continue;
        g.cell[posX][posY].push_back(i);
    }
}

// Frame::GetFeaturesInArea, reference src/Frame.cc:307-360 (KeyFrame::GetFeaturesInArea,
// src/KeyFrame.cc:549-588, is the same walk without the level test: minLevel = maxLevel = -1).
std::vector<size_t> features_in_area(const OrcFrame& F, const Grid& g, float x, float y, float r, int minLevel, int maxLevel) {
    std::vector<size_t> vIndices;
    const int nMinCellX = std::max(0, (int)floor((x - F.mnMinX - r) * F.mfGridElementWidthInv));
    if (nMinCellX >= FRAME_GRID_COLS) return vIndices;
    const int nMaxCellX = std::min((int)FRAME_GRID_COLS - 1, (int)ceil((x - F.mnMinX + r) * F.mfGridElementWidthInv));
    if (nMaxCellX < 0) return vIndices;
    const int nMinCellY = std::max(0, (int)floor((y - F.mnMinY - r) * F.mfGridElementHeightInv));
    if (nMinCellY >= FRAME_GRID_ROWS) return vIndices;
    const int nMaxCellY = std::min((int)FRAME_GRID_ROWS - 1, (int)ceil((y - F.mnMinY + r) * F.mfGridElementHeightInv));
    if (nMaxCellY < 0) return vIndices;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++) {
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
            const std::vector<size_t>& vCell = g.cell[ix][iy];
            for (size_t j = 0, jend = vCell.size(); j < jend; j++) {
                const KeyPoint& kpUn = F.keysUn[vCell[j]];
                if (bCheckLevels) {
                    if (kpUn.octave < minLevel) continue;
                    if (maxLevel >= 0)
                        if (kpUn.octave > maxLevel) continue;
                }
                const float distx = kpUn.x - x;
                const float disty = kpUn.y - y;
                if (fabs(distx) < r && fabs(disty) < r) vIndices.push_back(vCell[j]);
            }
        }
    }
    return vIndices;
}
}  // namespace

// Candidate lists alone (for testing the window generation): CSR offsets + indices.  Returns the total.
int orc_features_in_area(const OrcFrame* F, int nq, const float* x, const float* y, const float* r, const int* minLevel,
                         const int* maxLevel, int* offsets, int* cand, int cap) {
    Grid* g = new Grid();
    assign_features_to_grid(*F, *g);
    int total = 0;
    offsets[0] = 0;
    for (int i = 0; i < nq; ++i) {
        auto v = features_in_area(*F, *g, x[i], y[i], r[i], minLevel ? minLevel[i] : -1, maxLevel ? maxLevel[i] : -1);
        for (size_t idx : v) {
            if (total < cap) cand[total] = (int)idx;
            ++total;
        }
        offsets[i + 1] = total;
    }
    delete g;
    return total;
}

// ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th), reference ORBmatcher.cc:19-65.
// Queries = the map points that survive :24.  occupied[idx] stands for
// `F.mvpMapPoints[idx] && F.mvpMapPoints[idx]->Observations() > 0`; qObserved for the query's Observations() > 0.
int orc_search_by_projection_map(const OrcFrame* F, const float* mvuRight, uint8_t* occupied, const float* mvScaleFactors, int nq,
                                 const uint8_t* qdesc, const float* projX, const float* projY, const float* projXR, const int* level,
                                 const float* viewCos, const uint8_t* qObserved, float th, float mfNNratio, int* featureOfQuery) {
    const int TH_HIGH = 100;
    Grid* g = new Grid();
    assign_features_to_grid(*F, *g);
    int nmatches = 0;
    const bool bFactor = th != 1.0;
    for (int q = 0; q < nq; ++q) {
        featureOfQuery[q] = -1;
        const int nPredictedLevel = level[q];
        const float r = ((viewCos[q] > 0.998) ? 2.5f : 4.0f) * (bFactor ? th : 1);
        auto vIndices = features_in_area(*F, *g, projX[q], projY[q], r * mvScaleFactors[nPredictedLevel], nPredictedLevel - 1, nPredictedLevel);
        if (vIndices.empty()) continue;
        int bestDist = INT_MAX, bestIdx = -1, secondBestDist = INT_MAX;
        for (auto idx : vIndices) {
            if (occupied[idx]) continue;
            if (mvuRight && mvuRight[idx] > 0) {
                float er = fabs(projXR[q] - mvuRight[idx]);
                if (er > r * mvScaleFactors[nPredictedLevel]) continue;
            }
            int dist = descriptor_distance(qdesc + 32 * (size_t)q, F->desc + 32 * idx);
            if (dist < bestDist) {
                secondBestDist = bestDist;
                bestDist = dist;
                bestIdx = (int)idx;
            } else if (dist < secondBestDist) {
                secondBestDist = dist;
            }
        }
        if (bestDist <= TH_HIGH && (bestDist <= mfNNratio * secondBestDist)) {
            featureOfQuery[q] = bestIdx;
            if (!qObserved || qObserved[q]) occupied[bestIdx] = 1;
            nmatches++;
        }
    }
    delete g;
    return nmatches;
}

// ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono), reference
// ORBmatcher.cc:732-818, from the projected (u, v) on (:755-761 are the caller's).  Queries = last-frame
// features that reach :763.  Returns nmatches, or INT_MIN when a negative bin would index rotHist out of
// bounds (D9: the reference has no `rot += 360`).
int orc_search_by_projection_last(const OrcFrame* Cur, uint8_t* curHasMapPoint, const float* mvScaleFactors, int nq, const uint8_t* qdesc,
                                  const float* u, const float* v, const int* lastOctave, const float* lastAngle, float th, int bForward,
                                  int bBackward, int checkOri, int* featureOfQuery) {
    const int TH_HIGH = 100, HISTO = 30;
    Grid* g = new Grid();
    assign_features_to_grid(*Cur, *g);
    std::vector<int> rotHist[30];
    std::vector<int> owner(Cur->N, -1);
    const float factor = 1.0f / HISTO;
    int nmatches = 0;
    bool ub = false;
    for (int i = 0; i < nq && !ub; ++i) {
        featureOfQuery[i] = -1;
        int nLastOctave = lastOctave[i];
        float radius = th * mvScaleFactors[nLastOctave];
        std::vector<size_t> vIndices2;
        if (bForward)
            vIndices2 = features_in_area(*Cur, *g, u[i], v[i], radius, nLastOctave, -1);
        else if (bBackward)
            vIndices2 = features_in_area(*Cur, *g, u[i], v[i], radius, 0, nLastOctave);
        else
            vIndices2 = features_in_area(*Cur, *g, u[i], v[i], radius, nLastOctave - 1, nLastOctave + 1);
        if (vIndices2.empty()) continue;
        int bestDist = INT_MAX, bestIdx2 = -1;
        for (auto idx : vIndices2) {
            if (curHasMapPoint[idx]) continue;
            int dist = descriptor_distance(qdesc + 32 * (size_t)i, Cur->desc + 32 * idx);
            if (dist < bestDist) {
                bestDist = dist;
                bestIdx2 = (int)idx;
            }
        }
        if (bestDist <= TH_HIGH) {
            curHasMapPoint[bestIdx2] = 1;
            featureOfQuery[i] = bestIdx2;
            owner[bestIdx2] = i;
            nmatches++;
            if (checkOri) {
                float rot = lastAngle[i] - Cur->keysUn[bestIdx2].angle;
                int bin = static_cast<int>(round(rot * factor)) % HISTO;
                if (bin < 0) {
                    ub = true;
                    break;
                }
                rotHist[bin].push_back(bestIdx2);
            }
        }
    }
    if (checkOri && !ub) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO, ind1, ind2, ind3);
        for (int i = 0; i < HISTO; i++) {
            if (i != ind1 && i != ind2 && i != ind3) {
                for (auto idx : rotHist[i]) {
                    curHasMapPoint[idx] = 0;
                    featureOfQuery[owner[idx]] = -1;
                    nmatches--;
                }
            }
        }
    }
    delete g;
    return ub ? INT_MIN : nmatches;
}

// ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*>&, th, ORBdist),
// reference ORBmatcher.cc:820-894, from the projected (u, v) and predicted level on.
int orc_search_by_projection_reloc(const OrcFrame* Cur, uint8_t* curHasMapPoint, const float* mvScaleFactors, int nq, const uint8_t* qdesc,
                                   const float* u, const float* v, const int* predictedLevel, const float* kfAngle, float th,
                                   int ORBdist, int checkOri, int* featureOfQuery) {
    const int HISTO = 30;
    Grid* g = new Grid();
    assign_features_to_grid(*Cur, *g);
    std::vector<int> rotationHistogram[30];
    std::vector<int> owner(Cur->N, -1);
    float histogramFactor = 1.0f / HISTO;
    int matchesCount = 0;
    for (int i = 0; i < nq; ++i) {
        featureOfQuery[i] = -1;
        float radius = th * mvScaleFactors[predictedLevel[i]];
        auto candidates = features_in_area(*Cur, *g, u[i], v[i], radius, predictedLevel[i] - 1, predictedLevel[i] + 1);
        if (candidates.empty()) continue;
        int bestDist = INT_MAX, bestIdx = -1;
        for (size_t idx : candidates) {
            if (curHasMapPoint[idx]) continue;
            int dist = descriptor_distance(qdesc + 32 * (size_t)i, Cur->desc + 32 * idx);
            if (dist < bestDist) {
                bestDist = dist;
                bestIdx = (int)idx;
            }
        }
        if (bestDist <= ORBdist) {
            curHasMapPoint[bestIdx] = 1;
            featureOfQuery[i] = bestIdx;
            owner[bestIdx] = i;
            matchesCount++;
            if (checkOri) {
                float rotationDiff = kfAngle[i] - Cur->keysUn[bestIdx].angle;
                if (rotationDiff < 0) rotationDiff += 360.0f;
                int bin = static_cast<int>(round(rotationDiff * histogramFactor)) % HISTO;
                rotationHistogram[bin].push_back(bestIdx);
            }
        }
    }
    if (checkOri) {
        int topBins[3];
        three_maxima(rotationHistogram, HISTO, topBins[0], topBins[1], topBins[2]);
        for (int i = 0; i < HISTO; ++i) {
            if (i != topBins[0] && i != topBins[1] && i != topBins[2]) {
                for (size_t idx : rotationHistogram[i]) {
                    curHasMapPoint[idx] = 0;
                    featureOfQuery[owner[idx]] = -1;
                    matchesCount--;
                }
            }
        }
    }
    delete g;
    return matchesCount;
}

// The KeyFrame-window searches from the projected (u, v, radius) on: SearchByProjection(KeyFrame*, Scw, ...)
// reference ORBmatcher.cc:166-191 (vpMatched = claimed, no octave gate: level == NULL, accept <= TH_LOW);
// SearchBySim3 :694-714 and Fuse :533-547 / :597-611 (no claims: claimed == NULL, octave gate
// `kp.octave < level-1 || kp.octave > level`, accept <= maxDist).
int orc_search_kf_window(const OrcFrame* KF, uint8_t* claimed, int nq, const uint8_t* qdesc, const float* u, const float* v,
                         const float* radius, const int* level, int maxDist, int* featureOfQuery) {
    Grid* g = new Grid();
    assign_features_to_grid(*KF, *g);
    int nmatches = 0;
    for (int i = 0; i < nq; ++i) {
        featureOfQuery[i] = -1;
        const auto vIndices = features_in_area(*KF, *g, u[i], v[i], radius[i], -1, -1);
        if (vIndices.empty()) continue;
        int bestDist = INT_MAX, bestIdx = -1;
        for (auto idx : vIndices) {
            if (claimed && claimed[idx]) continue;
            if (level) {
                const KeyPoint& kp = KF->keysUn[idx];
                if (kp.octave < level[i] - 1 || kp.octave > level[i]) continue;
            }
            int dist = descriptor_distance(qdesc + 32 * (size_t)i, KF->desc + 32 * idx);
            if (dist < bestDist) {
                bestDist = dist;
                bestIdx = (int)idx;
            }
        }
        if (bestDist <= maxDist) {
            if (claimed) claimed[bestIdx] = 1;
            featureOfQuery[i] = bestIdx;
            nmatches++;
        }
    }
    delete g;
    return nmatches;
}

// ORBmatcher::SearchForInitialization, reference ORBmatcher.cc:197-276.  vbPrevMatched as (x, y) pairs.
int orc_search_for_initialization(const KeyPoint* keys1, const uint8_t* desc1, int n1, const OrcFrame* F2, float* vbPrevMatched,
                                  int windowSize, float mfNNratio, int checkOri, int* vnMatches12) {
    const int TH_LOW = 50, HISTO = 30;
    Grid* g = new Grid();
    assign_features_to_grid(*F2, *g);
    int nmatches = 0;
    for (int i = 0; i < n1; ++i) vnMatches12[i] = -1;
    std::vector<int> rotHist[30];
    const float factor = 1.0f / HISTO;
    std::vector<int> vMatchedDistance(F2->N, INT_MAX);
    std::vector<int> vnMatches21(F2->N, -1);
    for (int i1 = 0; i1 < n1; ++i1) {
        const KeyPoint& kp1 = keys1[i1];
        if (kp1.octave > 0) continue;
        auto vIndices2 = features_in_area(*F2, *g, vbPrevMatched[2 * i1], vbPrevMatched[2 * i1 + 1], windowSize, kp1.octave, kp1.octave);
        if (vIndices2.empty()) continue;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (auto i2 : vIndices2) {
            int dist = descriptor_distance(desc1 + 32 * (size_t)i1, F2->desc + 32 * i2);
            if (dist < vMatchedDistance[i2]) {
                if (dist < bestDist) {
                    bestDist2 = bestDist;
                    bestDist = dist;
                    bestIdx2 = (int)i2;
                } else if (dist < bestDist2) {
                    bestDist2 = dist;
                }
            }
        }
        if (bestDist <= TH_LOW && bestDist < static_cast<float>(bestDist2) * mfNNratio) {
            if (vnMatches21[bestIdx2] >= 0) {
                vnMatches12[vnMatches21[bestIdx2]] = -1;
                nmatches--;
            }
            vnMatches12[i1] = bestIdx2;
            vnMatches21[bestIdx2] = i1;
            vMatchedDistance[bestIdx2] = bestDist;
            nmatches++;
            if (checkOri) {
                float rot = kp1.angle - F2->keysUn[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = round(rot * factor);
                if (bin == HISTO) bin = 0;
                rotHist[bin].push_back(i1);
            }
        }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO, ind1, ind2, ind3);
        for (int i = 0; i < HISTO; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx1 : rotHist[i]) {
                if (vnMatches12[idx1] >= 0) {
                    vnMatches12[idx1] = -1;
                    nmatches--;
                }
            }
        }
    }
    for (int i1 = 0; i1 < n1; ++i1)
        if (vnMatches12[i1] >= 0) {
            vbPrevMatched[2 * i1] = F2->keysUn[vnMatches12[i1]].x;
            vbPrevMatched[2 * i1 + 1] = F2->keysUn[vnMatches12[i1]].y;
        }
    delete g;
    return nmatches;
}

// ---- DBoW2 vocabulary transform ------------------------------------------------------------------
// TemplatedVocabulary::transform(features, BowVector&, FeatureVector&, levelsup), reference
// Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1126-1187, with the per-feature descent :1218-1260,
// FORB::distance (FORB.cpp:81-101), BowVector::addWeight / addIfNotExist / normalize (BowVector.cpp:34-84)
// and FeatureVector::addFeature (FeatureVector.cpp:32-46).  The tree arrives flattened (node 0 = root,
// children of node i = children[child_off[i] .. child_off[i+1]) in m_nodes[i].children order).
// Pinned against the reference's own DBoW2 compiled from source (oracle/_ref/libref_dbow.so) in
// tests/test_vocabulary_oracle.py.  Returns 0; -1 capacity; -3 when a feature with weight > 0 ends above
// level L - levelsup (the reference then stores an uninitialised NodeId).
int orc_voc_transform(int n_nodes, const int* child_off, const int* children, const uint8_t* node_desc, const double* node_weight,
                      const int* node_word, int m_L, int weighting, int scoring, const uint8_t* feat, int n, int levelsup,
                      int* word_of_feature, int* node_of_feature, int* bow_ids, double* bow_vals, int bow_cap, int* bow_n,
                      int* fv_nodes, int* fv_off, int* fv_idx, int fv_cap, int* fv_n) {
    *bow_n = 0;
    *fv_n = 0;
    fv_off[0] = 0;
    if (n_nodes < 2) return 0;  // empty()
    std::map<unsigned, double> v;
    std::map<unsigned, std::vector<unsigned>> fv;
    const bool must = scoring != 5;                 // DotProductScoring is the only MUSTNORMALIZE = false
    const bool normL2 = scoring == 1;
    const bool tf = weighting == 0 || weighting == 1;  // TF_IDF, TF
    for (int i_feature = 0; i_feature < n; ++i_feature) {
        const uint8_t* feature = feat + 32 * (size_t)i_feature;
        // transform(feature, id, w, &nid, levelsup)
        const int nid_level = m_L - levelsup;
        long nid = -1;
        if (nid_level <= 0) nid = 0;
        int final_id = 0, current_level = 0;
        do {
            ++current_level;
            const int b = child_off[final_id], e = child_off[final_id + 1];
            final_id = children[b];
            double best_d = descriptor_distance(feature, node_desc + 32 * (size_t)final_id);
            for (int c = b + 1; c < e; ++c) {
                const int id = children[c];
                double d = descriptor_distance(feature, node_desc + 32 * (size_t)id);
                if (d < best_d) {
                    best_d = d;
                    final_id = id;
                }
            }
            if (current_level == nid_level) nid = final_id;
        } while (child_off[final_id] != child_off[final_id + 1]);
        const unsigned id = (unsigned)node_word[final_id];
        const double w = node_weight[final_id];
        if (word_of_feature) word_of_feature[i_feature] = (int)id;
        if (node_of_feature) node_of_feature[i_feature] = (int)nid;
        if (w > 0) {
            if (nid < 0) return -3;
            auto vit = v.lower_bound(id);
            if (vit != v.end() && !(v.key_comp()(id, vit->first))) {
                if (tf) vit->second += w;
            } else {
                v.insert(vit, std::make_pair(id, w));
            }
            fv[(unsigned)nid].push_back((unsigned)i_feature);
        }
    }
    if (tf && !v.empty() && !must) {
        const double nd = v.size();
        for (auto& e : v) e.second /= nd;
    }
    if (must) {
        double norm = 0.0;
        if (!normL2) {
            for (auto& e : v) norm += fabs(e.second);
        } else {
            for (auto& e : v) norm += e.second * e.second;
            norm = sqrt(norm);
        }
        if (norm > 0.0)
            for (auto& e : v) e.second /= norm;
    }
    *bow_n = (int)v.size();
    *fv_n = (int)fv.size();
    if ((int)v.size() > bow_cap || (int)fv.size() > fv_cap) return -1;
    int k = 0;
    for (auto& e : v) {
        bow_ids[k] = (int)e.first;
        bow_vals[k] = e.second;
        ++k;
    }
    k = 0;
    int c = 0;
    for (auto& e : fv) {
        fv_nodes[k] = (int)e.first;
        for (unsigned idx : e.second) fv_idx[c++] = (int)idx;
        fv_off[++k] = c;
    }
    return 0;
}

// ---- image ingest -----------------------------------------------------------------------------------
// cv::cvtColor(src, dst, CV_RGB2GRAY / CV_BGR2GRAY) on 8-bit images, as called by Tracking::GrabImage*
// (reference src/Tracking.cc:118-126, :136-141, :155-160).  OpenCV is an external dependency of the reference
// (3.4.15); the arithmetic restated here is its published fixed-point form:
//   variant 4 (OpenCV 4.x, pinned against cv2 4.13.0): (R*9798 + G*19235 + B*3735 + (1 << 14)) >> 15
//   variant 3 (OpenCV 2.4 - 3.4 color.cpp: yuv_shift = 14, R2Y = 4899, G2Y = 9617, B2Y = 1868; not checkable here):
//             (R*4899 + G*9617 + B*1868 + (1 << 13)) >> 14
// channels = 3 or 4 (alpha ignored); bgr != 0: the first channel is blue.
void orc_cvt_gray(const uint8_t* src, int rows, int cols, int step, int channels, int bgr, int variant, uint8_t* dst, int dstep) {
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            const uint8_t* p = src + (size_t)y * step + (size_t)x * channels;
            const int r = bgr ? p[2] : p[0], g = p[1], b = bgr ? p[0] : p[2];
            dst[(size_t)y * dstep + x] = variant == 3 ? (uint8_t)((r * 4899 + g * 9617 + b * 1868 + (1 << 13)) >> 14)
                                                      : (uint8_t)((r * 9798 + g * 19235 + b * 3735 + (1 << 14)) >> 15);
        }
}

// cv::remap(src, dst, map1, map2, cv::INTER_LINEAR) with CV_32FC1 maps and the default BORDER_CONSTANT 0, as called by
// Examples/Stereo/stereo_euroc.cc:136-137 (maps from cv::initUndistortRectifyMap, :97-98).  OpenCV imgwarp.cpp:
// coordinates are rounded to 1/32 px (sx = cvRound(map1 * 32), ix = sx >> 5 saturated to short, fx = sx & 31), the four
// taps are blended with the 15-bit table BilinearTab_i[fy][fx] = cvRound of the float products, saturated to short
// (so 1.0 becomes 32767 and the missing unit goes to the last tap), result (sum + (1 << 14)) >> 15; taps outside the
// source read 0.  Pinned against cv2 4.13.0 (0 mismatches, tests/test_ingest_oracle.py).
void orc_remap_linear(const uint8_t* src, int srows, int scols, int sstep, int channels, const float* mapx, const float* mapy,
                      int mstep, int drows, int dcols, uint8_t* dst, int dstep) {
    for (int y = 0; y < drows; ++y)
        for (int x = 0; x < dcols; ++x) {
            const int sx = cv_round_f(mapx[(size_t)y * mstep + x] * 32.0f), sy = cv_round_f(mapy[(size_t)y * mstep + x] * 32.0f);
            const int fx = sx & 31, fy = sy & 31;
            int ix = sx >> 5, iy = sy >> 5;
            ix = ix < -32768 ? -32768 : (ix > 32767 ? 32767 : ix);
            iy = iy < -32768 ? -32768 : (iy > 32767 ? 32767 : iy);
            const float a = 1.0f - fy / 32.0f, b = fy / 32.0f, c = 1.0f - fx / 32.0f, d = fx / 32.0f;
            const float t[4] = {a * c, a * d, b * c, b * d};
            int w[4], isum = 0;
            for (int k = 0; k < 4; ++k) {
                int v = cv_round_f(t[k] * 32768.0f);
                w[k] = v > 32767 ? 32767 : (v < -32768 ? -32768 : v);
                isum += w[k];
            }
            if (isum != 32768) w[3] -= isum - 32768;  // only (fy, fx) = (0, 0): {32767, 0, 0, 1}
            for (int ch = 0; ch < channels; ++ch) {
                auto tap = [&](int yy, int xx) -> int {
                    return (xx >= 0 && yy >= 0 && xx < scols && yy < srows) ? src[(size_t)yy * sstep + (size_t)xx * channels + ch] : 0;
                };
                const int v = tap(iy, ix) * w[0] + tap(iy, ix + 1) * w[1] + tap(iy + 1, ix) * w[2] + tap(iy + 1, ix + 1) * w[3];
                int o = (v + (1 << 14)) >> 15;
                o = o < 0 ? 0 : (o > 255 ? 255 : o);
                dst[(size_t)y * dstep + (size_t)x * channels + ch] = (uint8_t)o;
            }
        }
}

void orc_synth_frame(uint8_t* dst, int rows, int cols, int step, uint64_t seed, uint64_t frame,
                     int variant, int right) {
    synth_frame(dst, rows, cols, step, seed, frame, variant, right);
}

double orc_extract_many(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh,
                        int rows, int cols, uint64_t seed, int first_frame, int nframes,
                        int nthreads, long long* total_keypoints) {
    Params p{nfeatures, scaleFactor, nlevels, iniTh, minTh};
    std::vector<std::vector<uint8_t>> frames(nframes);
    {
        std::atomic<int> next{0};
        std::vector<std::thread> th;
        for (int w = 0; w < nthreads; ++w)
            th.emplace_back([&] {
                for (int f; (f = next.fetch_add(1)) < nframes;) {
                    frames[f].resize((size_t)rows * cols);
                    synth_frame(frames[f].data(), rows, cols, cols, seed, (uint64_t)(first_frame + f), 0, 0);
                }
            });
        for (auto& t : th) t.join();
    }
    std::atomic<int> next{0};
    std::atomic<long long> total{0};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int w = 0; w < nthreads; ++w)
        th.emplace_back([&] {
            for (int f; (f = next.fetch_add(1)) < nframes;) {
                ExtractResult r;
                extract(p, Image{frames[f].data(), rows, cols, cols}, r, false);
                total += (long long)r.keypoints.size();
            }
        });
    for (auto& t : th) t.join();
    auto t1 = std::chrono::steady_clock::now();
    if (total_keypoints) *total_keypoints = total.load();
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
