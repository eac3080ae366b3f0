// Hand-written sm_100a Hamming-search kernels.
//
// Unit: ORBmatcher::DescriptorDistance (reference src/ORBmatcher.cc:896-908) = popcount of
// the XOR of two 256-bit descriptors; here 2 x 128-bit loads + 8 x POPC.
// Scan semantics (src/ORBmatcher.cc:49-55, :225-231, :321-327): candidates in list order,
// strict '<' updates of (best, second); best = first minimum, second = second smallest of
// the multiset.  Tensor cores are not used: this is not a dense contraction.
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

#include "match_kernels.h"

__device__ __forceinline__ int hamming256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// ------------------------------------------------------------------------------------------
// Brute force: one thread per query, train descriptors staged through shared memory in
// tiles and read back as warp-wide broadcasts.  blockIdx.y = (query set, train set) pair.
// ------------------------------------------------------------------------------------------
#define MA_THREADS 256
#define MA_TILE 256

__global__ void __launch_bounds__(MA_THREADS) k_match_all(const uint8_t* __restrict__ q, const int* __restrict__ nq,
                                                          size_t q_stride, const uint8_t* __restrict__ t,
                                                          const int* __restrict__ nt, size_t t_stride,
                                                          int* __restrict__ best_idx, int* __restrict__ best_dist,
                                                          int* __restrict__ second_dist, size_t out_stride) {
    __shared__ uint4 tile[MA_TILE][2];
    const int p = blockIdx.y;
    const int nQ = nq[p], nT = nt[p];
    if ((int)(blockIdx.x * MA_THREADS) >= nQ) return;
    const int qi = blockIdx.x * MA_THREADS + threadIdx.x;
    const uint4* Q = reinterpret_cast<const uint4*>(q + p * q_stride);
    const uint4* T = reinterpret_cast<const uint4*>(t + p * t_stride);
    uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0;
    if (qi < nQ) {
        q0 = __ldg(Q + 2 * (size_t)qi);
        q1 = __ldg(Q + 2 * (size_t)qi + 1);
    }
    int best = INT_MAX, second = INT_MAX, idx = -1;
    // This fork's descriptors have bits 182..255 always zero, i.e. words 6 and 7 are zero.
    // When that holds for every query of the warp and every train row of the tile (checked on the data,
    // so the result is exact for any input) the distance needs 6 POPC instead of 8 -- the POPC pipe
    // (quarter rate) is what bounds this kernel.
    const bool qUpperZero = __all_sync(0xffffffffu, (q1.z | q1.w) == 0u);
    for (int t0 = 0; t0 < nT; t0 += MA_TILE) {
        const int cnt = min(MA_TILE, nT - t0);
        __syncthreads();
        unsigned upper = 0;
        for (int i = threadIdx.x; i < cnt * 2; i += MA_THREADS) {
            const uint4 v = __ldg(T + 2 * (size_t)t0 + i);
            (&tile[0][0])[i] = v;
            if (i & 1) upper |= v.z | v.w;
        }
        const bool tUpperZero = __syncthreads_or(upper != 0u) == 0;
        if (qUpperZero && tUpperZero) {
#pragma unroll 4
            for (int j = 0; j < cnt; ++j) {
                const uint4 a = tile[j][0];
                const uint2 b = *reinterpret_cast<const uint2*>(&tile[j][1]);
                const int d = __popc(q0.x ^ a.x) + __popc(q0.y ^ a.y) + __popc(q0.z ^ a.z) + __popc(q0.w ^ a.w) +
                              __popc(q1.x ^ b.x) + __popc(q1.y ^ b.y);
                if (d < best) {
                    second = best;
                    best = d;
                    idx = t0 + j;
                } else if (d < second) {
                    second = d;
                }
            }
        } else {
#pragma unroll 4
            for (int j = 0; j < cnt; ++j) {
                const int d = hamming256(q0, q1, tile[j][0], tile[j][1]);
                if (d < best) {
                    second = best;
                    best = d;
                    idx = t0 + j;
                } else if (d < second) {
                    second = d;
                }
            }
        }
    }
    if (qi < nQ) {
        best_idx[p * out_stride + qi] = idx;
        best_dist[p * out_stride + qi] = best;
        second_dist[p * out_stride + qi] = second;
    }
}

// ------------------------------------------------------------------------------------------
// Windowed (CSR) search: one warp per query, lanes stride over the candidate list, then a
// warp-shuffle (min, second-min, position) merge that reproduces the sequential scan.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_match_csr(const uint8_t* __restrict__ q, int nq, const uint8_t* __restrict__ t,
                                                   const int* __restrict__ offsets, const int* __restrict__ cand,
                                                   int tie_last, int max_dist, int* __restrict__ best_idx,
                                                   int* __restrict__ best_dist, int* __restrict__ second_dist) {
    const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nq) return;
    const uint4* Q = reinterpret_cast<const uint4*>(q) + 2 * (size_t)warp;
    const uint4 q0 = __ldg(Q), q1 = __ldg(Q + 1);
    const uint4* T = reinterpret_cast<const uint4*>(t);
    const int beg = __ldg(offsets + warp), end = __ldg(offsets + warp + 1);
    int best = tie_last ? max_dist : INT_MAX, second = INT_MAX, idx = -1;
    int pos = tie_last ? -1 : INT_MAX;
    for (int c = beg + lane; c < end; c += 32) {
        const int j = __ldg(cand + c);
        const int d = hamming256(q0, q1, __ldg(T + 2 * (size_t)j), __ldg(T + 2 * (size_t)j + 1));
        if (tie_last) {
            if (d <= max_dist && d <= best) {
                best = d;
                idx = j;
                pos = c;
            }
        } else if (d < best) {
            second = best;
            best = d;
            idx = j;
            pos = c;
        } else if (d < second) {
            second = d;
        }
    }
    for (int s = 16; s > 0; s >>= 1) {
        const int ob = __shfl_xor_sync(0xffffffffu, best, s);
        const int os = __shfl_xor_sync(0xffffffffu, second, s);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, s);
        const int op = __shfl_xor_sync(0xffffffffu, pos, s);
        const bool take = tie_last ? (ob < best || (ob == best && op > pos)) : (ob < best || (ob == best && op < pos));
        if (!tie_last) second = min(max(best, ob), min(second, os));
        if (take) {
            best = ob;
            idx = oi;
            pos = op;
        }
    }
    if (lane == 0) {
        best_idx[warp] = idx;
        best_dist[warp] = best;
        second_dist[warp] = second;
    }
}

// ------------------------------------------------------------------------------------------
// Stereo: Hamming part of Frame::ComputeStereoMatches (reference src/Frame.cc:446-529).
// k_stereo_prep turns every right keypoint into (minr, maxr, x, octave) -- the rows of
// vRowIndices it would be listed in (:463-473).  k_stereo_match: one warp per left keypoint,
// right keypoints scanned in ascending index order (the order vRowIndices lists them in).
// ------------------------------------------------------------------------------------------
__global__ void k_stereo_prep(const orb_kp28* __restrict__ kr, int nr, const float* __restrict__ scale, int4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nr) return;
    const float y = kr[i].y;
    const float r = __fmul_rn(2.0f, scale[kr[i].octave]);
    const int maxr = (int)ceilf(__fadd_rn(y, r));
    const int minr = (int)floorf(__fsub_rn(y, r));
    out[i] = make_int4(minr, maxr, __float_as_int(kr[i].x), kr[i].octave);
}

__global__ void __launch_bounds__(256) k_stereo_match(const orb_kp28* __restrict__ kl, const uint8_t* __restrict__ dl, int nl,
                                                      const int4* __restrict__ rinfo, const uint8_t* __restrict__ dr, int nr,
                                                      float maxD, int* __restrict__ best_r, int* __restrict__ best_dist) {
    const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nl) return;
    const float uL = kl[warp].x, vL = kl[warp].y;
    const int levelL = kl[warp].octave;
    const int row = (int)vL;  // vRowIndices[(size_t)vL]
    const float minU = __fsub_rn(uL, maxD), maxU = uL;  // minD = 0
    int best = 100, idx = -1;                           // TH_HIGH, strict '<'
    if (!(maxU < 0.f)) {
        const uint4* Q = reinterpret_cast<const uint4*>(dl) + 2 * (size_t)warp;
        const uint4 q0 = __ldg(Q), q1 = __ldg(Q + 1);
        const uint4* T = reinterpret_cast<const uint4*>(dr);
        for (int i = lane; i < nr; i += 32) {
            const int4 ri = __ldg(rinfo + i);
            if (row < ri.x || row > ri.y) continue;
            if (ri.w < levelL - 1 || ri.w > levelL + 1) continue;
            const float uR = __int_as_float(ri.z);
            if (uR >= minU && uR <= maxU) {
                const int d = hamming256(q0, q1, __ldg(T + 2 * (size_t)i), __ldg(T + 2 * (size_t)i + 1));
                if (d < best) {
                    best = d;
                    idx = i;
                }
            }
        }
    }
    for (int s = 16; s > 0; s >>= 1) {
        const int ob = __shfl_xor_sync(0xffffffffu, best, s);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, s);
        // first minimum in ascending right index; lanes without a hit hold idx = -1, best = 100
        if (oi >= 0 && (ob < best || (ob == best && (idx < 0 || oi < idx)))) {
            best = ob;
            idx = oi;
        }
    }
    if (lane == 0) {
        best_r[warp] = idx;
        best_dist[warp] = best;
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
void orbk_count_launch(int n);

cudaError_t orbk_match_all(const uint8_t* q, const int* nq, size_t q_stride, const uint8_t* t, const int* nt,
                           size_t t_stride, int npairs, int max_nq, int* best_idx, int* best_dist, int* second_dist,
                           size_t out_stride, cudaStream_t st) {
    if (npairs <= 0 || max_nq <= 0) return cudaSuccess;
    dim3 grid((max_nq + MA_THREADS - 1) / MA_THREADS, npairs);
    k_match_all<<<grid, MA_THREADS, 0, st>>>(q, nq, q_stride, t, nt, t_stride, best_idx, best_dist, second_dist, out_stride);
    orbk_count_launch(1);
    return cudaGetLastError();
}

cudaError_t orbk_match_csr(const uint8_t* q, int nq, const uint8_t* t, const int* offsets, const int* cand, int tie_last,
                           int max_dist, int* best_idx, int* best_dist, int* second_dist, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    k_match_csr<<<(nq + 7) / 8, 256, 0, st>>>(q, nq, t, offsets, cand, tie_last, max_dist, best_idx, best_dist, second_dist);
    orbk_count_launch(1);
    return cudaGetLastError();
}

cudaError_t orbk_stereo(const orb_kp28* kl, const uint8_t* dl, int nl, const orb_kp28* kr, const uint8_t* dr, int nr,
                        const float* d_scale, int4* d_rinfo, float maxD, int* best_r, int* best_dist, cudaStream_t st) {
    if (nl <= 0) return cudaSuccess;
    if (nr > 0) {
        k_stereo_prep<<<(nr + 255) / 256, 256, 0, st>>>(kr, nr, d_scale, d_rinfo);
        orbk_count_launch(1);
    }
    k_stereo_match<<<(nl + 7) / 8, 256, 0, st>>>(kl, dl, nl, d_rinfo, dr, nr, maxD, best_r, best_dist);
    orbk_count_launch(1);
    return cudaGetLastError();
}
