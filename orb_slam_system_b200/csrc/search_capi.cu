// The reference's search methods as whole calls of the C ABI (include/orb_b200.h, "search methods"
// section): candidate windows and every candidate distance come from the GPU (k_grid_assign, k_window,
// k_dist_csr); the accept rule and the greedy "already matched" state of each method are replayed in
// query order on the host, worded like the reference so the float comparisons round the same way
// (this file is compiled with -ffp-contract=off).  Citations: reference src/ORBmatcher.cc.
#include <limits.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <vector>

#include "capi_internal.h"
#include "match_kernels.h"

namespace {

const int TH_HIGH = 100;    // :13
const int TH_LOW = 50;      // :14
const int HISTO_LENGTH = 30;  // :15

struct Windows {
    std::vector<int32_t> off, cand, dist;
};

// Grid + GetFeaturesInArea + distances for nq queries against one frame view; results on the host.
int run_windows(orb_matcher* m, const orb_frame_view* F, int nq, const uint8_t* qdesc, const float* x, const float* y, const float* r,
                const int32_t* minl, const int32_t* maxl, Windows& W) {
    W.off.assign((size_t)nq + 1, 0);
    W.cand.clear();
    W.dist.clear();
    if (nq == 0 || F->n == 0) return ORB_OK;
    CUDA_TRY(cudaSetDevice(m->device));
    cudaStream_t st = m->stream;
    void *dk, *dd, *dgrid, *dq, *dqf, *dql, *dcnt;
    int rc;
    const size_t n = (size_t)F->n;
    if ((rc = orb_matcher_scratch(m, 8, n * 28 + 32, &dk)) || (rc = orb_matcher_scratch(m, 9, n * 32 + 32, &dd)) ||
        (rc = orb_matcher_scratch(m, 10, sizeof(int) * (64 * 48 + 1 + n), &dgrid)) ||
        (rc = orb_matcher_scratch(m, 11, (size_t)nq * 32 + 32, &dq)) || (rc = orb_matcher_scratch(m, 12, sizeof(float) * 3 * (size_t)nq, &dqf)) ||
        (rc = orb_matcher_scratch(m, 13, sizeof(int) * 2 * (size_t)nq, &dql)) ||
        (rc = orb_matcher_scratch(m, 14, sizeof(int) * (2 * (size_t)nq + 1), &dcnt)))
        return rc;
    CUDA_TRY(cudaMemcpyAsync(dk, F->keys_un, n * 28, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(dd, F->desc, n * 32, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(dq, qdesc, (size_t)nq * 32, cudaMemcpyHostToDevice, st));
    float* dx = (float*)dqf;
    float* dy = dx + nq;
    float* dr = dy + nq;
    CUDA_TRY(cudaMemcpyAsync(dx, x, sizeof(float) * nq, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(dy, y, sizeof(float) * nq, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(dr, r, sizeof(float) * nq, cudaMemcpyHostToDevice, st));
    int* dmin = nullptr;
    int* dmax = nullptr;
    if (minl) {
        dmin = (int*)dql;
        CUDA_TRY(cudaMemcpyAsync(dmin, minl, sizeof(int) * nq, cudaMemcpyHostToDevice, st));
    }
    if (maxl) {
        dmax = (int*)dql + nq;
        CUDA_TRY(cudaMemcpyAsync(dmax, maxl, sizeof(int) * nq, cudaMemcpyHostToDevice, st));
    }
    int* cell_start = (int*)dgrid;
    int* members = cell_start + 64 * 48 + 1;
    int* counts = (int*)dcnt;
    int* offsets = counts + nq;
    CUDA_TRY(orbk_grid_assign((const orb_kp28*)dk, F->n, F->min_x, F->min_y, F->grid_w_inv, F->grid_h_inv, cell_start, members, st));
    CUDA_TRY(orbk_window_count((const orb_kp28*)dk, cell_start, members, F->min_x, F->min_y, F->grid_w_inv, F->grid_h_inv, nq, dx, dy, dr,
                               dmin, dmax, counts, offsets, st));
    CUDA_TRY(cudaMemcpyAsync(W.off.data(), offsets, sizeof(int) * ((size_t)nq + 1), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const int total = W.off[nq];
    if (total == 0) return ORB_OK;
    void* dout;
    if ((rc = orb_matcher_scratch(m, 15, sizeof(int) * 2 * (size_t)total, &dout))) return rc;
    int* dcand = (int*)dout;
    int* ddist = dcand + total;
    CUDA_TRY(orbk_window_fill((const orb_kp28*)dk, (const uint8_t*)dd, cell_start, members, F->min_x, F->min_y, F->grid_w_inv,
                              F->grid_h_inv, nq, (const uint8_t*)dq, dx, dy, dr, dmin, dmax, offsets, dcand, ddist, st));
    W.cand.resize(total);
    W.dist.resize(total);
    CUDA_TRY(cudaMemcpyAsync(W.cand.data(), dcand, sizeof(int) * (size_t)total, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(W.dist.data(), ddist, sizeof(int) * (size_t)total, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return ORB_OK;
}

int check_view(const orb_frame_view* F) {
    if (!F) return orb_fail(ORB_ERR_INVALID, "null frame view");
    if (F->n < 0) return orb_fail(ORB_ERR_INVALID, "negative keypoint count");
    if (F->n > 0 && (!F->keys_un || !F->desc)) return orb_fail(ORB_ERR_INVALID, "frame view without keypoints/descriptors");
    return ORB_OK;
}

// ORBmatcher::ComputeThreeMaxima (:469-502) on bin sizes.
void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
    int idx[3] = {-1, -1, -1};
    int val[3] = {0, 0, 0};
    for (int i = 0; i < L; ++i) {
        const int s = (int)histo[i].size();
        int j = 0;
        while (j < 3 && !(s > val[j])) ++j;
        if (j == 3) continue;
        for (int k = 2; k > j; --k) {
            val[k] = val[k - 1];
            idx[k] = idx[k - 1];
        }
        val[j] = s;
        idx[j] = i;
    }
    ind1 = idx[0];
    ind2 = idx[1];
    ind3 = idx[2];
    if (val[1] < 0.1f * val[0]) {
        ind2 = -1;
        ind3 = -1;
    } else if (val[2] < 0.1f * val[0]) {
        ind3 = -1;
    }
}

// Shared-node walk of two DBoW2 feature vectors (:303-352): std::map iteration with lower_bound jumps
// = merge of two ascending id lists.  Calls f(a, b) for every pair of positions with equal node id.
template <class Fn>
void for_shared_nodes(const orb_feature_vector* A, const orb_feature_vector* B, Fn f) {
    int a = 0, b = 0;
    while (a < A->n_nodes && b < B->n_nodes) {
        if (A->nodes[a] == B->nodes[b]) {
            f(a, b);
            ++a;
            ++b;
        } else if (A->nodes[a] < B->nodes[b]) {
            ++a;
        } else {
            ++b;
        }
    }
}

int check_fv(const orb_feature_vector* fv, int n, const char* name) {
    if (!fv || fv->n_nodes < 0) return orb_fail(ORB_ERR_INVALID, "%s: null feature vector", name);
    if (fv->n_nodes == 0) return ORB_OK;
    if (!fv->nodes || !fv->off || !fv->idx) return orb_fail(ORB_ERR_INVALID, "%s: null arrays", name);
    if (fv->off[0] != 0) return orb_fail(ORB_ERR_INVALID, "%s: off[0] != 0", name);
    for (int k = 0; k < fv->n_nodes; ++k) {
        if (fv->off[k + 1] < fv->off[k]) return orb_fail(ORB_ERR_INVALID, "%s: offsets decrease", name);
        if (k && fv->nodes[k] <= fv->nodes[k - 1]) return orb_fail(ORB_ERR_INVALID, "%s: node ids must ascend", name);
    }
    for (int c = 0; c < fv->off[fv->n_nodes]; ++c)
        if (fv->idx[c] < 0 || fv->idx[c] >= n) return orb_fail(ORB_ERR_INVALID, "%s: feature index out of range", name);
    return ORB_OK;
}

// Distances of every (feature of KF1 in a shared node) x (eligible feature of KF2 in that node) pair.
// Rows of the CSR = the KF1 features in visiting order; q1[row] = its index.
struct BowPairs {
    std::vector<int32_t> q1, off, cand, dist;
};

int bow_distances(orb_matcher* m, const uint8_t* desc1, int n1, const uint8_t* desc2, int n2, const orb_feature_vector* fv1,
                  const orb_feature_vector* fv2, const uint8_t* skip1, bool skip1_if_set, const uint8_t* skip2, bool skip2_if_set,
                  BowPairs& P) {
    P.off.assign(1, 0);
    for_shared_nodes(fv1, fv2, [&](int a, int b) {
        for (int ia = fv1->off[a]; ia < fv1->off[a + 1]; ++ia) {
            const int i1 = fv1->idx[ia];
            if ((skip1[i1] != 0) == skip1_if_set) continue;
            P.q1.push_back(i1);
            for (int ib = fv2->off[b]; ib < fv2->off[b + 1]; ++ib) {
                const int i2 = fv2->idx[ib];
                if ((skip2[i2] != 0) == skip2_if_set) continue;
                P.cand.push_back(i2);
            }
            P.off.push_back((int32_t)P.cand.size());
        }
    });
    P.dist.assign(P.cand.size(), 0);
    if (P.cand.empty()) return ORB_OK;
    // gather the query rows so that row r of the CSR is descriptor q1[r]
    std::vector<uint8_t> q((size_t)P.q1.size() * 32);
    for (size_t r = 0; r < P.q1.size(); ++r) memcpy(&q[r * 32], desc1 + (size_t)P.q1[r] * 32, 32);
    (void)n1;
    return orb_distances_csr(m, q.data(), (int)P.q1.size(), desc2, n2, P.off.data(), P.cand.data(), P.dist.data());
}

}  // namespace

extern "C" int orb_window_search(orb_matcher* m, const orb_frame_view* F, int nq, const uint8_t* qdesc, const float* x, const float* y,
                                 const float* r, const int32_t* min_level, const int32_t* max_level, int32_t* offsets, int32_t* cand,
                                 int32_t* dist, int cap, int* total) {
    if (!m || !offsets || !total) return orb_fail(ORB_ERR_INVALID, "null argument");
    int rc = check_view(F);
    if (rc) return rc;
    if (nq < 0 || cap < 0) return orb_fail(ORB_ERR_INVALID, "negative count");
    if (nq > 0 && (!qdesc || !x || !y || !r)) return orb_fail(ORB_ERR_INVALID, "null query arrays");
    Windows W;
    if ((rc = run_windows(m, F, nq, qdesc, x, y, r, min_level, max_level, W))) return rc;
    memcpy(offsets, W.off.data(), sizeof(int32_t) * ((size_t)nq + 1));
    *total = W.off[nq];
    if (*total > cap) return orb_fail(ORB_ERR_CAPACITY, "%d candidates, capacity %d", *total, cap);
    if (*total) {
        if (!cand || !dist) return orb_fail(ORB_ERR_INVALID, "null output arrays");
        memcpy(cand, W.cand.data(), sizeof(int32_t) * (size_t)*total);
        memcpy(dist, W.dist.data(), sizeof(int32_t) * (size_t)*total);
    }
    return ORB_OK;
}

// ---- SearchByProjection(Frame&, const vector<MapPoint*>&, th), :19-65 -----------------------------
extern "C" int orb_search_by_projection_map(orb_matcher* m, const orb_frame_view* F, const float* u_right, uint8_t* occupied,
                                            const float* scale_factors, int nlevels, int nq, const uint8_t* qdesc, const float* proj_x,
                                            const float* proj_y, const float* proj_xr, const int32_t* level, const float* view_cos,
                                            const uint8_t* q_observed, float th, float nnratio, int32_t* feature_of_query,
                                            int* nmatches) {
    if (!m || !nmatches) return orb_fail(ORB_ERR_INVALID, "null argument");
    int rc = check_view(F);
    if (rc) return rc;
    if (nq < 0 || nlevels < 1) return orb_fail(ORB_ERR_INVALID, "bad count");
    *nmatches = 0;
    if (nq == 0) return ORB_OK;
    if (!qdesc || !proj_x || !proj_y || !level || !view_cos || !scale_factors || !feature_of_query || (F->n && !occupied))
        return orb_fail(ORB_ERR_INVALID, "null argument");
    if (u_right && !proj_xr) return orb_fail(ORB_ERR_INVALID, "u_right without proj_xr");
    const bool bFactor = th != 1.0;  // :21
    std::vector<float> rad(nq), win(nq);
    std::vector<int32_t> lo(nq), hi(nq);
    for (int i = 0; i < nq; ++i) {
        if (level[i] < 0 || level[i] >= nlevels) return orb_fail(ORB_ERR_INVALID, "query %d: level out of range", i);
        const float r = (view_cos[i] > 0.998 ? 2.5f : 4.0f) * (bFactor ? th : 1);  // RadiusByViewingCos :67-69, :27
        rad[i] = r;
        win[i] = r * scale_factors[level[i]];  // :28
        lo[i] = level[i] - 1;
        hi[i] = level[i];
        feature_of_query[i] = -1;
    }
    Windows W;
    if ((rc = run_windows(m, F, nq, qdesc, proj_x, proj_y, win.data(), lo.data(), hi.data(), W))) return rc;
    int count = 0;
    for (int i = 0; i < nq; ++i) {
        const int b = W.off[i], e = W.off[i + 1];
        if (b == e) continue;  // :30
        int bestDist = INT_MAX, bestIdx = -1, secondBestDist = INT_MAX;
        for (int c = b; c < e; ++c) {
            const int idx = W.cand[c];
            if (occupied[idx]) continue;  // :38
            if (u_right && u_right[idx] > 0) {  // :41-45
                const float er = fabsf(proj_xr[i] - u_right[idx]);
                if (er > rad[i] * scale_factors[level[i]]) continue;
            }
            const int dist = W.dist[c];
            if (dist < bestDist) {
                secondBestDist = bestDist;
                bestDist = dist;
                bestIdx = idx;
            } else if (dist < secondBestDist) {
                secondBestDist = dist;
            }
        }
        if (bestDist <= TH_HIGH && (bestDist <= nnratio * secondBestDist)) {  // :58
            feature_of_query[i] = bestIdx;
            if (!q_observed || q_observed[i]) occupied[bestIdx] = 1;  // F.mvpMapPoints[bestIdx] = pMP, seen through :38
            ++count;
        }
    }
    *nmatches = count;
    return ORB_OK;
}

// ---- the best-only projection searches, :732-818, :820-894, :121-195, :636-730, :504-634 -------------
extern "C" int orb_search_by_projection_best(orb_matcher* m, const orb_frame_view* F, uint8_t* claimed, int nq, const uint8_t* qdesc,
                                             const float* u, const float* v, const float* radius, const int32_t* min_level,
                                             const int32_t* max_level, const float* q_angle, int rot_mode, int max_dist,
                                             int32_t* feature_of_query, int* nmatches) {
    if (!m || !nmatches) return orb_fail(ORB_ERR_INVALID, "null argument");
    int rc = check_view(F);
    if (rc) return rc;
    if (nq < 0) return orb_fail(ORB_ERR_INVALID, "bad count");
    if (rot_mode != ORB_ROT_NONE && rot_mode != ORB_ROT_WRAP && rot_mode != ORB_ROT_NOWRAP) return orb_fail(ORB_ERR_INVALID, "bad rot_mode");
    *nmatches = 0;
    if (nq == 0) return ORB_OK;
    if (!qdesc || !u || !v || !radius || !feature_of_query) return orb_fail(ORB_ERR_INVALID, "null argument");
    if (rot_mode != ORB_ROT_NONE && !q_angle) return orb_fail(ORB_ERR_INVALID, "rotation check without query angles");
    Windows W;
    if ((rc = run_windows(m, F, nq, qdesc, u, v, radius, min_level, max_level, W))) return rc;
    std::vector<int> rotHist[HISTO_LENGTH];
    std::vector<int> owner;  // query that claimed a feature (for the histogram reset)
    if (rot_mode != ORB_ROT_NONE) owner.assign((size_t)F->n, -1);
    const float factor = 1.0f / HISTO_LENGTH;
    int count = 0;
    for (int i = 0; i < nq; ++i) {
        feature_of_query[i] = -1;
        int bestDist = INT_MAX, bestIdx2 = -1;
        for (int c = W.off[i]; c < W.off[i + 1]; ++c) {
            const int idx = W.cand[c];
            if (claimed && claimed[idx]) continue;  // :783, :855, :177
            const int dist = W.dist[c];
            if (dist < bestDist) {
                bestDist = dist;
                bestIdx2 = idx;
            }
        }
        if (bestDist <= max_dist) {  // :792, :867, :188, :711, :547
            feature_of_query[i] = bestIdx2;
            if (claimed) claimed[bestIdx2] = 1;
            ++count;
            if (rot_mode != ORB_ROT_NONE) {
                float rot = q_angle[i] - F->keys_un[bestIdx2].angle;
                if (rot_mode == ORB_ROT_WRAP && rot < 0) rot += 360.0f;  // :874
                const int bin = static_cast<int>(round(rot * factor)) % HISTO_LENGTH;  // :797, :875
                if (bin < 0) return orb_fail(ORB_ERR_SHAPE, "query %d: negative rotation bin %d (the reference indexes rotHist out of bounds)", i, bin);
                rotHist[bin].push_back(bestIdx2);
                owner[bestIdx2] = i;
            }
        }
    }
    if (rot_mode != ORB_ROT_NONE) {  // :803-815, :881-891
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int b = 0; b < HISTO_LENGTH; ++b) {
            if (b == ind1 || b == ind2 || b == ind3) continue;
            for (int idx : rotHist[b]) {
                if (claimed) claimed[idx] = 0;  // CurrentFrame.mvpMapPoints[idx] = nullptr
                if (owner[idx] >= 0) feature_of_query[owner[idx]] = -1;
                --count;
            }
        }
    }
    *nmatches = count;
    return ORB_OK;
}

// ---- SearchForInitialization, :197-276 ----------------------------------------------------------------
extern "C" int orb_search_for_initialization(orb_matcher* m, const orb_keypoint* keys1, const uint8_t* desc1, int n1,
                                             const orb_frame_view* F2, float* prev_matched, int window_size, float nnratio,
                                             int check_ori, int32_t* matches12, int* nmatches) {
    if (!m || !nmatches) return orb_fail(ORB_ERR_INVALID, "null argument");
    int rc = check_view(F2);
    if (rc) return rc;
    if (n1 < 0) return orb_fail(ORB_ERR_INVALID, "bad count");
    *nmatches = 0;
    if (n1 == 0) return ORB_OK;
    if (!keys1 || !desc1 || !prev_matched || !matches12) return orb_fail(ORB_ERR_INVALID, "null argument");
    // queries = the level-0 keypoints of F1 (:212), in index order
    std::vector<int32_t> qi, lo, hi;
    std::vector<float> x, y, r;
    std::vector<uint8_t> q;
    for (int i1 = 0; i1 < n1; ++i1) {
        matches12[i1] = -1;
        if (keys1[i1].octave > 0) continue;
        qi.push_back(i1);
        x.push_back(prev_matched[2 * i1]);
        y.push_back(prev_matched[2 * i1 + 1]);
        r.push_back((float)window_size);
        lo.push_back(keys1[i1].octave);
        hi.push_back(keys1[i1].octave);
        q.insert(q.end(), desc1 + (size_t)i1 * 32, desc1 + (size_t)i1 * 32 + 32);
    }
    Windows W;
    if ((rc = run_windows(m, F2, (int)qi.size(), q.data(), x.data(), y.data(), r.data(), lo.data(), hi.data(), W))) return rc;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    std::vector<int> vMatchedDistance((size_t)F2->n, INT_MAX), vnMatches21((size_t)F2->n, -1);
    int count = 0;
    for (size_t k = 0; k < qi.size(); ++k) {
        const int i1 = qi[k];
        if (W.off[k] == W.off[k + 1]) continue;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int c = W.off[k]; c < W.off[k + 1]; ++c) {
            const int i2 = W.cand[c], dist = W.dist[c];
            if (dist < vMatchedDistance[i2]) {  // :224
                if (dist < bestDist) {
                    bestDist2 = bestDist;
                    bestDist = dist;
                    bestIdx2 = i2;
                } else if (dist < bestDist2) {
                    bestDist2 = dist;
                }
            }
        }
        if (bestDist <= TH_LOW && bestDist < static_cast<float>(bestDist2) * nnratio) {  // :235
            if (vnMatches21[bestIdx2] >= 0) {
                matches12[vnMatches21[bestIdx2]] = -1;
                --count;
            }
            matches12[i1] = bestIdx2;
            vnMatches21[bestIdx2] = i1;
            vMatchedDistance[bestIdx2] = bestDist;
            ++count;
            if (check_ori) {
                float rot = keys1[i1].angle - F2->keys_un[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                if (bin < 0 || bin > HISTO_LENGTH) return orb_fail(ORB_ERR_SHAPE, "keypoint %d: rotation bin %d out of range (angles must lie in [0, 360))", i1, bin);
                rotHist[bin].push_back(i1);
            }
        }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int b = 0; b < HISTO_LENGTH; ++b) {
            if (b == ind1 || b == ind2 || b == ind3) continue;
            for (int idx1 : rotHist[b]) {
                if (matches12[idx1] >= 0) {
                    matches12[idx1] = -1;
                    --count;
                }
            }
        }
    }
    for (int i1 = 0; i1 < n1; ++i1)  // :270-273
        if (matches12[i1] >= 0) {
            prev_matched[2 * i1] = F2->keys_un[matches12[i1]].x;
            prev_matched[2 * i1 + 1] = F2->keys_un[matches12[i1]].y;
        }
    *nmatches = count;
    return ORB_OK;
}

// ---- SearchByBoW(KeyFrame*, KeyFrame*, ...), :278-366 --------------------------------------------------
extern "C" int orb_search_by_bow(orb_matcher* m, const uint8_t* desc1, const float* angle1, const uint8_t* has_mp1, int n1,
                                 const uint8_t* desc2, const float* angle2, const uint8_t* has_mp2, int n2, const orb_feature_vector* fv1,
                                 const orb_feature_vector* fv2, float nnratio, int check_ori, int32_t* matches12, int* nmatches) {
    if (!m || !nmatches) return orb_fail(ORB_ERR_INVALID, "null argument");
    if (n1 < 0 || n2 < 0) return orb_fail(ORB_ERR_INVALID, "bad count");
    int rc;
    if ((rc = check_fv(fv1, n1, "fv1")) || (rc = check_fv(fv2, n2, "fv2"))) return rc;
    *nmatches = 0;
    if (n1 && (!desc1 || !has_mp1 || !matches12 || (check_ori && !angle1))) return orb_fail(ORB_ERR_INVALID, "null argument");
    if (n2 && (!desc2 || !has_mp2 || (check_ori && !angle2))) return orb_fail(ORB_ERR_INVALID, "null argument");
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    if (n1 == 0 || n2 == 0) return ORB_OK;
    BowPairs P;  // rows: KF1 features with a map point; candidates: bucket mates of KF2 with a map point
    if ((rc = bow_distances(m, desc1, n1, desc2, n2, fv1, fv2, has_mp1, false, has_mp2, false, P))) return rc;
    std::vector<bool> vbMatched2((size_t)n2, false);
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    int count = 0;
    for (size_t row = 0; row < P.q1.size(); ++row) {
        const int idx1 = P.q1[row];
        int bestDist1 = INT_MAX, bestIdx2 = -1, bestDist2 = INT_MAX;
        for (int c = P.off[row]; c < P.off[row + 1]; ++c) {
            const int idx2 = P.cand[c];
            if (vbMatched2[idx2]) continue;  // :316
            const int dist = P.dist[c];
            if (dist < bestDist1) {
                bestDist2 = bestDist1;
                bestDist1 = dist;
                bestIdx2 = idx2;
            } else if (dist < bestDist2) {
                bestDist2 = dist;
            }
        }
        if (bestDist1 < TH_LOW && static_cast<float>(bestDist1) < nnratio * static_cast<float>(bestDist2)) {  // :329
            matches12[idx1] = bestIdx2;
            vbMatched2[bestIdx2] = true;
            ++count;
            if (check_ori) {
                float rot = angle1[idx1] - angle2[bestIdx2];
                if (rot < 0.0f) rot += 360.0f;
                int bin = (int)round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                if (bin < 0 || bin > HISTO_LENGTH) return orb_fail(ORB_ERR_SHAPE, "feature %d: rotation bin %d out of range (angles must lie in [0, 360))", idx1, bin);
                rotHist[bin].push_back(idx1);
            }
        }
    }
    if (check_ori) {  // :354-364: no `>= 0` guard here, a feature listed twice in fv1 is subtracted twice
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int b = 0; b < HISTO_LENGTH; ++b) {
            if (b == ind1 || b == ind2 || b == ind3) continue;
            for (int idx1 : rotHist[b]) {
                matches12[idx1] = -1;
                --count;
            }
        }
    }
    *nmatches = count;
    return ORB_OK;
}

// ---- SearchForTriangulation (bOnlyStereo = false), :368-467 with CheckDistEpipolarLine :71-85 ----------
extern "C" int orb_search_for_triangulation(orb_matcher* m, const orb_keypoint* keys1, const uint8_t* desc1, const uint8_t* has_mp1, int n1,
                                            const orb_keypoint* keys2, const uint8_t* desc2, const uint8_t* has_mp2, int n2,
                                            const orb_feature_vector* fv1, const orb_feature_vector* fv2, const float* f12,
                                            const float* sigma2, int nlevels, int check_ori, int32_t* matches12, int* nmatches) {
    if (!m || !nmatches || !f12 || !sigma2) return orb_fail(ORB_ERR_INVALID, "null argument");
    if (n1 < 0 || n2 < 0 || nlevels < 1) return orb_fail(ORB_ERR_INVALID, "bad count");
    int rc;
    if ((rc = check_fv(fv1, n1, "fv1")) || (rc = check_fv(fv2, n2, "fv2"))) return rc;
    *nmatches = 0;
    if (n1 && (!keys1 || !desc1 || !has_mp1 || !matches12)) return orb_fail(ORB_ERR_INVALID, "null argument");
    if (n2 && (!keys2 || !desc2 || !has_mp2)) return orb_fail(ORB_ERR_INVALID, "null argument");
    for (int i = 0; i < n2; ++i)
        if (keys2[i].octave < 0 || keys2[i].octave >= nlevels) return orb_fail(ORB_ERR_INVALID, "keypoint %d of KF2: octave out of range", i);
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    if (n1 == 0 || n2 == 0) return ORB_OK;
    BowPairs P;  // rows: KF1 features without a map point; candidates: bucket mates of KF2 without one (vbMatched2 is never set, D8)
    if ((rc = bow_distances(m, desc1, n1, desc2, n2, fv1, fv2, has_mp1, true, has_mp2, true, P))) return rc;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    int count = 0;
    for (size_t row = 0; row < P.q1.size(); ++row) {
        const int idx1 = P.q1[row];
        const orb_keypoint& kp1 = keys1[idx1];
        int bestDist = TH_LOW, bestIdx2 = -1;
        for (int c = P.off[row]; c < P.off[row + 1]; ++c) {
            const int idx2 = P.cand[c], dist = P.dist[c];
            if (dist > TH_LOW || dist > bestDist) continue;  // :414
            const orb_keypoint& kp2 = keys2[idx2];
            // CheckDistEpipolarLine: F12.at<float>(r, c) = f12[3*r + c]
            const float a = kp1.x * f12[0] + kp1.y * f12[3] + f12[6];
            const float b = kp1.x * f12[1] + kp1.y * f12[4] + f12[7];
            const float cc = kp1.x * f12[2] + kp1.y * f12[5] + f12[8];
            const float num = a * kp2.x + b * kp2.y + cc;
            const float den = a * a + b * b;
            if (den == 0) continue;
            const float dsqr = num * num / den;
            if (!(dsqr < 3.84 * sigma2[kp2.octave])) continue;
            bestIdx2 = idx2;
            bestDist = dist;
            if (check_ori) {
                float rot = kp1.angle - kp2.angle;
                if (rot < 0.0) rot += 360.0f;
                const int bin = static_cast<int>(round(rot * factor)) % HISTO_LENGTH;
                if (bin < 0) return orb_fail(ORB_ERR_SHAPE, "feature %d: negative rotation bin", idx1);
                rotHist[bin].push_back(idx1);
            }
        }
        if (bestIdx2 >= 0) {
            matches12[idx1] = bestIdx2;
            ++count;
        }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int b = 0; b < HISTO_LENGTH; ++b) {
            if (b == ind1 || b == ind2 || b == ind3) continue;
            for (int idx : rotHist[b])
                if (matches12[idx] >= 0) {
                    matches12[idx] = -1;
                    --count;
                }
        }
    }
    *nmatches = count;
    return ORB_OK;
}

// ---- DBoW2 vocabulary: TemplatedVocabulary::transform, TemplatedVocabulary.h:1126-1187, :1218-1260 --------------
struct orb_vocabulary {
    int device = 0;
    cudaStream_t stream = nullptr;
    int n_nodes = 0, depth_l = 0, weighting = 0, scoring = 0;
    int* d_child_off = nullptr;
    int* d_children = nullptr;
    uint8_t* d_desc = nullptr;
    std::vector<double> weight;
    std::vector<int32_t> word, nchild;
    // grow-only scratch for the features and the per-feature results
    uint8_t* d_feat = nullptr;
    int* d_out = nullptr;
    int* h_out = nullptr;  // pinned landing buffer of the per-feature results (2 ints per feature)
    size_t cap = 0;
    std::vector<unsigned long long> keys;  // host assembly scratch
};

extern "C" void orb_vocabulary_destroy(orb_vocabulary* v) {
    if (!v) return;
    cudaSetDevice(v->device);
    if (v->stream) cudaStreamSynchronize(v->stream);
    cudaFree(v->d_child_off);
    cudaFree(v->d_children);
    cudaFree(v->d_desc);
    cudaFree(v->d_feat);
    cudaFree(v->d_out);
    if (v->h_out) cudaFreeHost(v->h_out);
    if (v->stream) cudaStreamDestroy(v->stream);
    delete v;
}

static int vocabulary_upload(orb_vocabulary* v, const int32_t* child_off, const int32_t* children, const uint8_t* node_desc) {
    const int n = v->n_nodes, nc = child_off[n];
    CUDA_TRY(cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaMalloc((void**)&v->d_child_off, sizeof(int) * ((size_t)n + 1)));
    CUDA_TRY(cudaMalloc((void**)&v->d_children, sizeof(int) * (size_t)std::max(nc, 1)));
    CUDA_TRY(cudaMalloc((void**)&v->d_desc, (size_t)n * 32));
    CUDA_TRY(cudaMemcpyAsync(v->d_child_off, child_off, sizeof(int) * ((size_t)n + 1), cudaMemcpyHostToDevice, v->stream));
    if (nc) CUDA_TRY(cudaMemcpyAsync(v->d_children, children, sizeof(int) * (size_t)nc, cudaMemcpyHostToDevice, v->stream));
    CUDA_TRY(cudaMemcpyAsync(v->d_desc, node_desc, (size_t)n * 32, cudaMemcpyHostToDevice, v->stream));
    CUDA_TRY(cudaStreamSynchronize(v->stream));
    return ORB_OK;
}

extern "C" int orb_vocabulary_create(int device, int n_nodes, const int32_t* child_off, const int32_t* children, const uint8_t* node_desc,
                                     const double* node_weight, const int32_t* node_word, int depth_l, int weighting, int scoring,
                                     orb_vocabulary** out) {
    if (!out || !child_off || !node_desc || !node_weight || !node_word) return orb_fail(ORB_ERR_INVALID, "null argument");
    if (n_nodes < 1) return orb_fail(ORB_ERR_INVALID, "a vocabulary has at least the root node");
    if (weighting < 0 || weighting > 3 || scoring < 0 || scoring > 5) return orb_fail(ORB_ERR_INVALID, "bad weighting / scoring type");
    if (child_off[0] != 0) return orb_fail(ORB_ERR_INVALID, "child_off[0] != 0");
    for (int i = 0; i < n_nodes; ++i)
        if (child_off[i + 1] < child_off[i]) return orb_fail(ORB_ERR_INVALID, "child_off must not decrease");
    const int nc = child_off[n_nodes];
    if (nc && !children) return orb_fail(ORB_ERR_INVALID, "null children");
    // every child id in range and below... a tree: each non-root node is the child of exactly one node, no cycles
    std::vector<uint8_t> seen((size_t)n_nodes, 0);
    for (int c = 0; c < nc; ++c) {
        if (children[c] <= 0 || children[c] >= n_nodes) return orb_fail(ORB_ERR_INVALID, "child id out of range");
        if (seen[children[c]]++) return orb_fail(ORB_ERR_INVALID, "node %d has two parents", children[c]);
    }
    {  // reachability from the root without revisiting = no cycles among the listed nodes
        std::vector<int> stack(1, 0);
        size_t visited = 0;
        while (!stack.empty()) {
            const int u = stack.back();
            stack.pop_back();
            if (++visited > (size_t)n_nodes) return orb_fail(ORB_ERR_INVALID, "the node graph has a cycle");
            for (int c = child_off[u]; c < child_off[u + 1]; ++c) stack.push_back(children[c]);
        }
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return orb_fail(ORB_ERR_CUDA, "no CUDA device: %s (liborb_b200 has no CPU fallback)", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return orb_fail(ORB_ERR_INVALID, "device %d out of range", device);
    CUDA_TRY(cudaSetDevice(device));
    orb_vocabulary* v = new orb_vocabulary();
    v->device = device;
    v->n_nodes = n_nodes;
    v->depth_l = depth_l;
    v->weighting = weighting;
    v->scoring = scoring;
    v->weight.assign(node_weight, node_weight + n_nodes);
    v->word.assign(node_word, node_word + n_nodes);
    const int rc = vocabulary_upload(v, child_off, children, node_desc);
    if (rc != ORB_OK) {
        orb_vocabulary_destroy(v);
        return rc;
    }
    *out = v;
    return ORB_OK;
}

extern "C" int orb_vocabulary_transform(orb_vocabulary* v, const uint8_t* desc, int n, int levelsup, int32_t* word_of_feature,
                                        int32_t* node_of_feature, int32_t* bow_ids, double* bow_values, int bow_cap, int* bow_n,
                                        int32_t* fv_nodes, int32_t* fv_off, int32_t* fv_idx, int fv_cap, int* fv_n) {
    if (!v || !bow_n || !fv_n) return orb_fail(ORB_ERR_INVALID, "null argument");
    if (n < 0 || bow_cap < 0 || fv_cap < 0) return orb_fail(ORB_ERR_INVALID, "bad count");
    *bow_n = 0;
    *fv_n = 0;
    // empty(): a vocabulary without words (TemplatedVocabulary.h:1133-1136) -> both results stay empty
    if (v->n_nodes < 2 || n == 0) {
        if (fv_off && fv_cap >= 0) fv_off[0] = 0;
        return ORB_OK;
    }
    if (!desc) return orb_fail(ORB_ERR_INVALID, "null descriptors");
    CUDA_TRY(cudaSetDevice(v->device));
    if ((size_t)n > v->cap) {
        CUDA_TRY(cudaStreamSynchronize(v->stream));
        cudaFree(v->d_feat);
        cudaFree(v->d_out);
        if (v->h_out) cudaFreeHost(v->h_out);
        v->d_feat = nullptr;
        v->d_out = nullptr;
        v->h_out = nullptr;
        v->cap = 0;
        const size_t want = (size_t)n + (size_t)n / 4 + 64;
        CUDA_TRY(cudaMalloc((void**)&v->d_feat, want * 32));
        CUDA_TRY(cudaMalloc((void**)&v->d_out, want * 2 * sizeof(int)));
        CUDA_TRY(cudaMallocHost((void**)&v->h_out, want * 2 * sizeof(int)));
        v->cap = want;
    }
    CUDA_TRY(cudaMemcpyAsync(v->d_feat, desc, (size_t)n * 32, cudaMemcpyHostToDevice, v->stream));
    CUDA_TRY(orbk_voc_descent(v->d_feat, n, v->d_child_off, v->d_children, v->d_desc, v->depth_l - levelsup, v->d_out, v->d_out + n, v->stream));
    CUDA_TRY(cudaMemcpyAsync(v->h_out, v->d_out, sizeof(int) * 2 * (size_t)n, cudaMemcpyDeviceToHost, v->stream));
    CUDA_TRY(cudaStreamSynchronize(v->stream));
    const int32_t* leaf = v->h_out;
    const int32_t* nid = v->h_out + n;
    // ---- the two maps (:1143-1180).  The reference fills two std::maps feature by feature; the same contents come
    // out of sorting (key, feature index) pairs: a word's BoW value is its node weight added once per occurrence (every
    // occurrence adds the same double, so the sum does not depend on the order), a node's feature list is in ascending
    // feature index.
    const bool tf = v->weighting == ORB_VOC_TF_IDF || v->weighting == ORB_VOC_TF;
    const bool must = v->scoring != ORB_VOC_DOT_PRODUCT;       // ScoringObject.h:74-89
    const bool l2 = v->scoring == ORB_VOC_L2_NORM;
    std::vector<unsigned long long>& keys = v->keys;
    keys.clear();
    keys.reserve(2 * (size_t)n);
    for (int i = 0; i < n; ++i) {
        const int32_t id = v->word[leaf[i]];
        const double w = v->weight[leaf[i]];
        if (word_of_feature) word_of_feature[i] = id;
        if (node_of_feature) node_of_feature[i] = nid[i];
        if (w > 0) {  // not stopped
            if (nid[i] < 0)
                return orb_fail(ORB_ERR_SHAPE, "feature %d reaches a leaf above level L - levelsup (the reference stores an unset node id)", i);
            keys.push_back(((unsigned long long)(unsigned)id << 32) | (unsigned)i);
        }
    }
    const size_t m = keys.size();  // features that count
    for (size_t k = 0; k < m; ++k) keys.push_back(((unsigned long long)(unsigned)nid[(int)(keys[k] & 0xffffffffu)] << 32) | (keys[k] & 0xffffffffu));
    std::sort(keys.begin(), keys.begin() + m);
    std::sort(keys.begin() + m, keys.end());
    // distinct words / nodes
    int nb = 0, nf = 0;
    for (size_t k = 0; k < m; ++k) {
        nb += k == 0 || (keys[k] >> 32) != (keys[k - 1] >> 32);
        nf += k == 0 || (keys[m + k] >> 32) != (keys[m + k - 1] >> 32);
    }
    *bow_n = nb;
    *fv_n = nf;
    if (nb > bow_cap || nf > fv_cap) return orb_fail(ORB_ERR_CAPACITY, "needs %d BoW entries and %d feature-vector nodes", nb, nf);
    if ((nb && (!bow_ids || !bow_values)) || !fv_off || (nf && (!fv_nodes || !fv_idx))) return orb_fail(ORB_ERR_INVALID, "null output arrays");
    // BowVector: addWeight / addIfNotExist per occurrence, in ascending word id
    {
        int k = -1;
        for (size_t j = 0; j < m; ++j) {
            const int32_t id = (int32_t)(keys[j] >> 32);
            const double w = v->weight[leaf[(int)(keys[j] & 0xffffffffu)]];
            if (k < 0 || bow_ids[k] != id) {
                ++k;
                bow_ids[k] = id;
                bow_values[k] = w;
            } else if (tf) {
                bow_values[k] += w;  // addWeight; addIfNotExist leaves an existing entry alone
            }
        }
    }
    if (tf && nb > 0 && !must) {  // :1159-1165
        const double nd = (double)nb;
        for (int k = 0; k < nb; ++k) bow_values[k] /= nd;
    }
    if (must) {  // BowVector::normalize, BowVector.cpp:62-84
        double norm = 0.0;
        if (!l2) {
            for (int k = 0; k < nb; ++k) norm += fabs(bow_values[k]);
        } else {
            for (int k = 0; k < nb; ++k) norm += bow_values[k] * bow_values[k];
            norm = sqrt(norm);
        }
        if (norm > 0.0)
            for (int k = 0; k < nb; ++k) bow_values[k] /= norm;
    }
    // FeatureVector: node -> feature indices, ascending
    {
        int k = -1;
        fv_off[0] = 0;
        for (size_t j = 0; j < m; ++j) {
            const int32_t node = (int32_t)(keys[m + j] >> 32);
            if (k < 0 || fv_nodes[k] != node) {
                ++k;
                fv_nodes[k] = node;
            }
            fv_idx[j] = (int32_t)(keys[m + j] & 0xffffffffu);
            fv_off[k + 1] = (int)j + 1;
        }
    }
    return ORB_OK;
}
