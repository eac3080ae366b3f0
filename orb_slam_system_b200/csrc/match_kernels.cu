// Hand-written sm_100a Hamming-search kernels.
//
// Unit: ORBmatcher::DescriptorDistance (reference src/ORBmatcher.cc:896-908) = popcount of
// the XOR of two 256-bit descriptors; here 2 x 128-bit loads, XOR, and a carry-save adder tree that leaves 3 or 4 words to POPC (k_match_all) or 8 x POPC.
// Scan semantics (src/ORBmatcher.cc:49-55, :225-231, :321-327): candidates in list order,
// strict '<' updates of (best, second); best = first minimum, second = second smallest of
// the multiset.  Tensor cores are not used: this is not a dense contraction.
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "match_kernels.h"

__device__ __forceinline__ int hamming256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// Carry-save adder halves as single LOP3s.
__device__ __forceinline__ unsigned xor3(unsigned a, unsigned b, unsigned c) {
    unsigned r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ unsigned maj3(unsigned a, unsigned b, unsigned c) {
    unsigned r;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
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
                                                          int* __restrict__ second_dist, size_t out_stride, unsigned key_scale,
                                                          int only_huge) {
    __shared__ uint4 tile[MA_TILE][2];
    const int p = blockIdx.y;
    const int nQ = nq[p], nT = nt[p];
    if ((int)(blockIdx.x * MA_THREADS) >= nQ) return;
    if (only_huge && nT < (1 << 22)) return;  // the tensor-core kernel (match_mma.cu) has done this pair
    const int qi = blockIdx.x * MA_THREADS + threadIdx.x;
    const uint4* Q = reinterpret_cast<const uint4*>(q + p * q_stride);
    const uint4* T = reinterpret_cast<const uint4*>(t + p * t_stride);
    uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0;
    if (qi < nQ) {
        q0 = __ldg(Q + 2 * (size_t)qi);
        q1 = __ldg(Q + 2 * (size_t)qi + 1);
    }
    // This fork's descriptors have bits 182..255 always zero, i.e. words 6 and 7 are zero.
    // When that holds for every query of the warp and every train row of the tile (checked on the data,
    // so the result is exact for any input) the distance needs 6 words instead of 8.
    //
    // POPC is a quarter-rate pipe and bounds a popcount-per-word scan, so the XOR words of a pair first go through a
    // carry-save adder tree on the full-rate logic pipe (LOP3: sum = a^b^c, carry = maj(a,b,c)) that leaves one word
    // per binary weight: 6 words -> (ones, twos, fours), 8 words -> (ones, twos, fours, eights), i.e. 3 / 4 POPC instead
    // of 6 / 8, and d = p1 + 2 p2 + 4 p4 (+ 8 p8).
    // Scan state as keys d << 22 | j (j = train index < 2^22): key order refines distance order and the smallest key of
    // equal distances is the first one, so  best = min(best, key), second = min(second, max(key, best_before))  is the
    // reference's  if (d < best) {second = best; best = d; idx = j} else if (d < second) second = d.
    const bool qUpperZero = __all_sync(0xffffffffu, (q1.z | q1.w) == 0u);
    unsigned bestk = 0xffffffffu, seck = 0xffffffffu;
    int best = INT_MAX, second = INT_MAX, idx = -1;  // scan state of the plain path (train sets of 2^22 rows and more)
    const bool keyed = nT < (1 << 22);
    // key_scale = 1 << 22 arrives as a kernel argument so that the weighted sums stay multiply-adds on the FMA pipe
    // (as immediates ptxas turns them into shift-adds on the logic pipe, which is the busy one here)
    const unsigned C1 = key_scale, C2 = 2u * key_scale, C4 = 4u * key_scale, C8 = 8u * key_scale;
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
        if (!keyed) {
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
        } else if (qUpperZero && tUpperZero) {
#pragma unroll 4
            for (int j = 0; j < cnt; ++j) {
                const uint4 a = tile[j][0];
                const uint2 b = *reinterpret_cast<const uint2*>(&tile[j][1]);
                const unsigned x0 = q0.x ^ a.x, x1 = q0.y ^ a.y, x2 = q0.z ^ a.z, x3 = q0.w ^ a.w, x4 = q1.x ^ b.x, x5 = q1.y ^ b.y;
                const unsigned s1 = xor3(x0, x1, x2), c1 = maj3(x0, x1, x2);
                const unsigned s2 = xor3(x3, x4, x5), c2 = maj3(x3, x4, x5);
                const unsigned ones = s1 ^ s2, c3 = s1 & s2;
                const unsigned twos = xor3(c1, c2, c3), fours = maj3(c1, c2, c3);
                unsigned key = (unsigned)__popc(ones) * C1 + (unsigned)(t0 + j);
                key = (unsigned)__popc(twos) * C2 + key;
                key = (unsigned)__popc(fours) * C4 + key;
                seck = min(seck, max(key, bestk));
                bestk = min(bestk, key);
            }
        } else {
#pragma unroll 4
            for (int j = 0; j < cnt; ++j) {
                const uint4 a = tile[j][0], b = tile[j][1];
                const unsigned x0 = q0.x ^ a.x, x1 = q0.y ^ a.y, x2 = q0.z ^ a.z, x3 = q0.w ^ a.w;
                const unsigned x4 = q1.x ^ b.x, x5 = q1.y ^ b.y, x6 = q1.z ^ b.z, x7 = q1.w ^ b.w;
                const unsigned s1 = xor3(x0, x1, x2), c1 = maj3(x0, x1, x2);
                const unsigned s2 = xor3(x3, x4, x5), c2 = maj3(x3, x4, x5);
                const unsigned s3 = xor3(x6, x7, s1), c3 = maj3(x6, x7, s1);
                const unsigned ones = s2 ^ s3, c4 = s2 & s3;
                const unsigned t1 = xor3(c1, c2, c3), f1 = maj3(c1, c2, c3);
                const unsigned twos = t1 ^ c4, f2 = t1 & c4;
                const unsigned fours = f1 ^ f2, eights = f1 & f2;
                unsigned key = (unsigned)__popc(ones) * C1 + (unsigned)(t0 + j);
                key = (unsigned)__popc(twos) * C2 + key;
                key = (unsigned)__popc(fours) * C4 + key;
                key = (unsigned)__popc(eights) * C8 + key;
                seck = min(seck, max(key, bestk));
                bestk = min(bestk, key);
            }
        }
    }
    if (keyed) {
        if (bestk != 0xffffffffu) {
            best = (int)(bestk >> 22);
            idx = (int)(bestk & ((1u << 22) - 1u));
        }
        if (seck != 0xffffffffu) second = (int)(seck >> 22);
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
// All candidate distances of a windowed search: out[c] = distance(query i, train cand[c]) for
// c in [offsets[i], offsets[i+1]).  For the searches whose candidate eligibility depends on
// earlier accepts (vbMatched2 in SearchByBoW, vMatchedDistance in SearchForInitialization,
// claimed features in SearchByProjection): the host replays the reference's sequential scan
// over these numbers.  One warp per query, one candidate per lane and step.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dist_csr(const uint8_t* __restrict__ q, int nq, const uint8_t* __restrict__ t,
                                                  const int* __restrict__ offsets, const int* __restrict__ cand,
                                                  int* __restrict__ out) {
    const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nq) return;
    const uint4* Q = reinterpret_cast<const uint4*>(q) + 2 * (size_t)warp;
    const uint4 q0 = __ldg(Q), q1 = __ldg(Q + 1);
    const uint4* T = reinterpret_cast<const uint4*>(t);
    const int beg = __ldg(offsets + warp), end = __ldg(offsets + warp + 1);
    for (int c = beg + lane; c < end; c += 32) {
        const int j = __ldg(cand + c);
        out[c] = hamming256(q0, q1, __ldg(T + 2 * (size_t)j), __ldg(T + 2 * (size_t)j + 1));
    }
}

// ------------------------------------------------------------------------------------------
// Candidate windows on the device: Frame::AssignFeaturesToGrid / PosInGrid (reference
// src/Frame.cc:210-225, :362-372) and Frame::GetFeaturesInArea (:307-360; KeyFrame twin
// src/KeyFrame.cc:549-588).
//
// The 64 x 48 grid is kept in CSR form with cells numbered ix*48 + iy.  GetFeaturesInArea walks
// `for ix { for iy { members } }`, so for one ix the cells iy = minY..maxY are one contiguous
// slice of the member array: a query is at most 64 slices, each read with coalesced lanes.
// Members of a cell are in ascending keypoint index (the reference push_back's in index order).
// ------------------------------------------------------------------------------------------
#define GRID_COLS 64
#define GRID_ROWS 48
#define GRID_CELLS (GRID_COLS * GRID_ROWS)

__device__ __forceinline__ int pos_in_grid(const orb_kp28& k, float minX, float minY, float wInv, float hInv) {
    // posX = round((kp.pt.x - mnMinX) * mfGridElementWidthInv): float product, C round() (half away from zero)
    const float fx = roundf(__fmul_rn(__fsub_rn(k.x, minX), wInv));
    const float fy = roundf(__fmul_rn(__fsub_rn(k.y, minY), hInv));
    // NaN and out-of-range both fall out here; (int) of a huge float is UB in the reference, keys that far out never occur
    if (!(fx >= 0.0f && fx < (float)GRID_COLS && fy >= 0.0f && fy < (float)GRID_ROWS)) return -1;
    return (int)fx * GRID_ROWS + (int)fy;
}

// One CTA per frame view.  cell_start[GRID_CELLS + 1], members[n].
__global__ void __launch_bounds__(1024) k_grid_assign(const orb_kp28* __restrict__ keys, int n, float minX, float minY,
                                                      float wInv, float hInv, int* __restrict__ cell_start,
                                                      int* __restrict__ members) {
    __shared__ int cnt[GRID_CELLS];
    __shared__ int warp_sum[32];
    const int tid = threadIdx.x;
    for (int c = tid; c < GRID_CELLS; c += 1024) cnt[c] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 1024) {
        const int c = pos_in_grid(keys[i], minX, minY, wInv, hInv);
        if (c >= 0) atomicAdd(&cnt[c], 1);
    }
    __syncthreads();
    // exclusive scan of 3072 counters: 3 per thread
    const int a0 = cnt[3 * tid], a1 = cnt[3 * tid + 1], a2 = cnt[3 * tid + 2];
    int v = a0 + a1 + a2;
    const int lane = tid & 31, w = tid >> 5;
    int inc = v;
    for (int s = 1; s < 32; s <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, inc, s);
        if (lane >= s) inc += o;
    }
    if (lane == 31) warp_sum[w] = inc;
    __syncthreads();
    if (w == 0) {
        int ws = warp_sum[lane];
        for (int s = 1; s < 32; s <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, ws, s);
            if (lane >= s) ws += o;
        }
        warp_sum[lane] = ws;
    }
    __syncthreads();
    const int base = inc - v + (w ? warp_sum[w - 1] : 0);
    __syncthreads();
    cnt[3 * tid] = base;
    cnt[3 * tid + 1] = base + a0;
    cnt[3 * tid + 2] = base + a0 + a1;
    cell_start[3 * tid] = base;
    cell_start[3 * tid + 1] = base + a0;
    cell_start[3 * tid + 2] = base + a0 + a1;
    if (tid == 1023) cell_start[GRID_CELLS] = base + v;
    __syncthreads();
    // scatter (any order), then order every cell's few members by index
    for (int i = tid; i < n; i += 1024) {
        const int c = pos_in_grid(keys[i], minX, minY, wInv, hInv);
        if (c >= 0) members[atomicAdd(&cnt[c], 1)] = i;
    }
    __syncthreads();
    for (int c = tid; c < GRID_CELLS; c += 1024) {
        const int b = cell_start[c], e = cnt[c];  // cnt[c] now = end of cell c
        for (int i = b + 1; i < e; ++i) {
            const int key = members[i];
            int j = i - 1;
            while (j >= b && members[j] > key) {
                members[j + 1] = members[j];
                --j;
            }
            members[j + 1] = key;
        }
    }
}

// One warp per query.  FILL = false: counts[q] = number of candidates.  FILL = true: writes the
// candidate indices (reference order) and their Hamming distances at offsets[q].
template <bool FILL>
__global__ void __launch_bounds__(256) k_window(const orb_kp28* __restrict__ keys, const uint8_t* __restrict__ tdesc,
                                                const int* __restrict__ cell_start, const int* __restrict__ members,
                                                float minX, float minY, float wInv, float hInv, int nq,
                                                const uint8_t* __restrict__ qdesc, const float* __restrict__ qx,
                                                const float* __restrict__ qy, const float* __restrict__ qr,
                                                const int* __restrict__ qmin, const int* __restrict__ qmax,
                                                int* __restrict__ counts, const int* __restrict__ offsets,
                                                int* __restrict__ cand, int* __restrict__ dist) {
    const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nq) return;
    const float x = qx[warp], y = qy[warp], r = qr[warp];
    const int minLevel = qmin ? qmin[warp] : -1, maxLevel = qmax ? qmax[warp] : -1;
    int total = 0;
    // (int)floor((x - mnMinX - r) * inv), (int)ceil((x - mnMinX + r) * inv): float ops in source order
    const float fx0 = floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, minX), r), wInv));
    const float fx1 = ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, minX), r), wInv));
    const float fy0 = floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, minY), r), hInv));
    const float fy1 = ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, minY), r), hInv));
    // the comparisons below are the reference's early returns, done on floats so that huge values cannot overflow an int
    const bool empty = !(fx0 < (float)GRID_COLS) || !(fx1 >= 0.0f) || !(fy0 < (float)GRID_ROWS) || !(fy1 >= 0.0f);
    if (!empty) {
        const int cx0 = max(0, (int)fmaxf(fx0, -1.0f)), cx1 = min(GRID_COLS - 1, (int)fminf(fx1, (float)GRID_COLS));
        const int cy0 = max(0, (int)fmaxf(fy0, -1.0f)), cy1 = min(GRID_ROWS - 1, (int)fminf(fy1, (float)GRID_ROWS));
        const bool checkLevels = (minLevel > 0) || (maxLevel >= 0);
        uint4 q0, q1;
        int base = 0;
        if (FILL) {
            const uint4* Q = reinterpret_cast<const uint4*>(qdesc) + 2 * (size_t)warp;
            q0 = __ldg(Q);
            q1 = __ldg(Q + 1);
            base = offsets[warp];
        }
        const uint4* T = reinterpret_cast<const uint4*>(tdesc);
        for (int ix = cx0; ix <= cx1; ++ix) {
            const int b = __ldg(cell_start + ix * GRID_ROWS + cy0), e = __ldg(cell_start + ix * GRID_ROWS + cy1 + 1);
            for (int s = b; s < e; s += 32) {
                const int m = s + lane;
                bool ok = false;
                int j = 0;
                if (m < e) {
                    j = __ldg(members + m);
                    const orb_kp28 k = keys[j];
                    ok = true;
                    if (checkLevels) {
                        if (k.octave < minLevel) ok = false;
                        if (maxLevel >= 0 && k.octave > maxLevel) ok = false;
                    }
                    if (!(fabsf(__fsub_rn(k.x, x)) < r && fabsf(__fsub_rn(k.y, y)) < r)) ok = false;
                }
                const unsigned bal = __ballot_sync(0xffffffffu, ok);
                if (FILL && ok) {
                    const int o = base + total + __popc(bal & ((1u << lane) - 1u));
                    cand[o] = j;
                    dist[o] = hamming256(q0, q1, __ldg(T + 2 * (size_t)j), __ldg(T + 2 * (size_t)j + 1));
                }
                total += __popc(bal);
            }
        }
    }
    if (!FILL && lane == 0) counts[warp] = total;
}

// Exclusive scan of n ints by one CTA (n is a few thousand queries): out[0..n], out[n] = total.
__global__ void __launch_bounds__(1024) k_scan_counts(const int* __restrict__ in, int n, int* __restrict__ out) {
    __shared__ int warp_sum[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int b = 0; b < n; b += 1024) {
        const int i = b + tid;
        const int v = i < n ? in[i] : 0;
        int inc = v;
        for (int s = 1; s < 32; s <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, inc, s);
            if (lane >= s) inc += o;
        }
        if (lane == 31) warp_sum[w] = inc;
        __syncthreads();
        if (w == 0) {
            int ws = warp_sum[lane];
            for (int s = 1; s < 32; s <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, ws, s);
                if (lane >= s) ws += o;
            }
            warp_sum[lane] = ws;
        }
        __syncthreads();
        const int c = carry;
        if (i < n) out[i] = c + inc - v + (w ? warp_sum[w - 1] : 0);
        __syncthreads();
        if (tid == 0) carry = c + warp_sum[31];
        __syncthreads();
    }
    if (tid == 0) out[n] = carry;
}

// ------------------------------------------------------------------------------------------
// Stereo: Hamming part of Frame::ComputeStereoMatches (reference src/Frame.cc:446-529).
// k_stereo_prep turns every right keypoint into (minr, maxr, x, octave) -- the rows of
// vRowIndices it would be listed in (:463-473).  k_stereo_match: one warp per left keypoint,
// right keypoints scanned in ascending index order (the order vRowIndices lists them in).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void stereo_prep_one(const orb_kp28* __restrict__ kr, const float* __restrict__ scale, int4* __restrict__ out, int i) {
    const float y = kr[i].y;
    const float r = __fmul_rn(2.0f, scale[kr[i].octave]);
    const int maxr = (int)ceilf(__fadd_rn(y, r));
    const int minr = (int)floorf(__fsub_rn(y, r));
    out[i] = make_int4(minr, maxr, __float_as_int(kr[i].x), kr[i].octave);
}

__global__ void k_stereo_prep(const orb_kp28* __restrict__ kr, int nr, const float* __restrict__ scale, int4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nr) return;
    stereo_prep_one(kr, scale, out, i);
}

__device__ __forceinline__ void stereo_match_one(const orb_kp28* __restrict__ kl, const uint8_t* __restrict__ dl,
                                                 const int4* __restrict__ rinfo, const uint8_t* __restrict__ dr, int nr, float maxD,
                                                 int* __restrict__ best_r, int* __restrict__ best_dist, int warp, int lane) {
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

__global__ void __launch_bounds__(256) k_stereo_match(const orb_kp28* __restrict__ kl, const uint8_t* __restrict__ dl, int nl,
                                                      const int4* __restrict__ rinfo, const uint8_t* __restrict__ dr, int nr,
                                                      float maxD, int* __restrict__ best_r, int* __restrict__ best_dist) {
    const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nl) return;
    stereo_match_one(kl, dl, rinfo, dr, nr, maxD, best_r, best_dist, warp, lane);
}

// ------------------------------------------------------------------------------------------
// Stereo refinement: the rest of Frame::ComputeStereoMatches after the Hamming search
// (reference src/Frame.cc:531-603): 11x11 SAD of the centre-subtracted patches over 11 shifts on the
// pyramid level of the left keypoint, first minimum, parabola fit, disparity gate.  One warp per left
// keypoint; the left patch and the 11 x 21 right strip are staged in shared memory, lane s < 11 sums
// shift s.  Pixels are integers, so the float sums of the reference are exact integers here.
// Outputs per left keypoint: uRight / depth (-1 when rejected) and the best SAD (-1 when rejected) for
// the median cut (:606-619), which the caller does once it has all of them.
// flags[0] |= 1 when a rowRange / colRange would leave the level image (cv::Exception in the reference).
// ------------------------------------------------------------------------------------------
// frameL / frameR: which frames of the level buffers the pair lives in (lv.left[l] + frameL * lv.plane[l], ...)
__device__ __forceinline__ void stereo_refine_one(const orb_kp28* __restrict__ kl, const orb_kp28* __restrict__ kr,
                                                  const int* __restrict__ best_r, const int* __restrict__ best_dist,
                                                  const OrbStereoLevels& lv, size_t frameL, size_t frameR, float mbf, float maxD,
                                                  float* __restrict__ u_right, float* __restrict__ depth, int* __restrict__ sad,
                                                  int* __restrict__ flags, int i, int (&s_l)[8][121], int (&s_r)[8][11 * 21],
                                                  int (&s_d)[8][12]) {
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float outU = -1.0f, outD = -1.0f;
    int outS = -1;
    const int bi = best_r[i];
    if (bi >= 0 && best_dist[i] < 75) {  // thOrbDist = (TH_HIGH + TH_LOW) / 2, :451, :532
        const orb_kp28 kpL = kl[i];
        const int oct = kpL.octave;
        const float uR0 = kr[bi].x;
        const float sf = lv.inv_scale[oct];
        const float scaleduL = roundf(__fmul_rn(kpL.x, sf));
        const float scaledvL = roundf(__fmul_rn(kpL.y, sf));
        const float scaleduR0 = roundf(__fmul_rn(uR0, sf));
        const int R = lv.rows[oct], C = lv.cols[oct];
        const int r0 = (int)(scaledvL - 5.0f), c0 = (int)(scaleduL - 5.0f), q0 = (int)(scaleduR0 - 10.0f);
        bool ok = true;
        if (r0 < 0 || r0 + 11 > R || c0 < 0 || c0 + 11 > C) {  // IL's rowRange / colRange (:542)
            if (lane == 0) atomicOr(flags, 1);
            ok = false;
        }
        // iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1 (:553-556)
        if (ok && (scaleduR0 < 0.0f || scaleduR0 + 11.0f >= (float)C)) ok = false;
        if (ok && q0 < 0) {  // first colRange of the loop starts at scaleduR0 - L - w (:560)
            if (lane == 0) atomicOr(flags, 1);
            ok = false;
        }
        if (ok) {
            const uint8_t* PL = lv.left[oct] + frameL * lv.plane[oct] + (size_t)r0 * lv.pitch[oct] + c0;
            const uint8_t* PR = lv.right[oct] + frameR * lv.plane[oct] + (size_t)r0 * lv.pitch[oct] + q0;
            for (int e = lane; e < 121; e += 32) s_l[wib][e] = PL[(e / 11) * lv.pitch[oct] + (e % 11)];
            for (int e = lane; e < 231; e += 32) s_r[wib][e] = PR[(e / 21) * lv.pitch[oct] + (e % 21)];
            __syncwarp();
            if (lane < 11) {
                const int cL = s_l[wib][5 * 11 + 5], cR = s_r[wib][5 * 21 + lane + 5];
                const int dc = cL - cR;
                int acc = 0;
                for (int a = 0; a < 11; ++a)
#pragma unroll
                    for (int b = 0; b < 11; ++b) acc += abs(s_l[wib][a * 11 + b] - s_r[wib][a * 21 + b + lane] - dc);
                s_d[wib][lane] = acc;
            }
            __syncwarp();
            if (lane == 0) {
                int bestS = 2147483647, bestinc = 0;
                for (int s2 = 0; s2 < 11; ++s2)
                    if (s_d[wib][s2] < bestS) {  // float dist < int bestDist, exact
                        bestS = s_d[wib][s2];
                        bestinc = s2 - 5;
                    }
                if (bestinc != -5 && bestinc != 5) {
                    const float dist1 = (float)s_d[wib][5 + bestinc - 1], dist2 = (float)s_d[wib][5 + bestinc];
                    const float dist3 = (float)s_d[wib][5 + bestinc + 1];
                    const float deltaR = __fdiv_rn(__fsub_rn(dist1, dist3),
                                                   __fmul_rn(2.0f, __fsub_rn(__fadd_rn(dist1, dist3), __fmul_rn(2.0f, dist2))));
                    if (!(deltaR < -1.0f || deltaR > 1.0f)) {
                        float bestuR = __fmul_rn(lv.scale[oct], __fadd_rn(__fadd_rn(scaleduR0, (float)bestinc), deltaR));
                        float disparity = __fsub_rn(kpL.x, bestuR);
                        if (disparity >= 0.0f && disparity < maxD) {
                            if (disparity <= 0.0f) {
                                disparity = 0.01f;                       // float(0.01)
                                bestuR = (float)((double)kpL.x - 0.01);  // uL - 0.01 in double, then to float
                            }
                            outD = __fdiv_rn(mbf, disparity);
                            outU = bestuR;
                            outS = bestS;
                        }
                    }
                }
            }
            __syncwarp();
        }
    }
    if (lane == 0) {
        u_right[i] = outU;
        depth[i] = outD;
        sad[i] = outS;
    }
}

__global__ void __launch_bounds__(256) k_stereo_refine(const orb_kp28* __restrict__ kl, int nl, const orb_kp28* __restrict__ kr,
                                                       const int* __restrict__ best_r, const int* __restrict__ best_dist,
                                                       const OrbStereoLevels lv, float mbf, float maxD, float* __restrict__ u_right,
                                                       float* __restrict__ depth, int* __restrict__ sad, int* __restrict__ flags) {
    __shared__ int s_l[8][121];
    __shared__ int s_r[8][11 * 21];
    __shared__ int s_d[8][12];
    const int i = (blockIdx.x * 256 + threadIdx.x) >> 5;
    if (i >= nl) return;
    stereo_refine_one(kl, kr, best_r, best_dist, lv, 0, 0, mbf, maxD, u_right, depth, sad, flags, i, s_l, s_r, s_d);
}

// ------------------------------------------------------------------------------------------
// Frame::ComputeStereoMatches for every stereo pair of one extractor batch (pair p = frames 2p / 2p + 1, keypoints and
// descriptors in the batch layout [frame][cap]): the same three steps with blockIdx.y = pair, then the median cut
// (src/Frame.cc:606-619) on the device, one CTA per pair.  flags[p]: bit 0 = a row band / SAD window leaves the image
// (the reference faults), bit 1 = an octave out of range.
// ------------------------------------------------------------------------------------------
__global__ void k_stereo_prep_batch(const orb_kp28* __restrict__ kps, const int* __restrict__ counts, int cap, int nlevels, int rows,
                                    const float* __restrict__ scale, int4* __restrict__ rinfo, int* __restrict__ flags) {
    const int p = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    const orb_kp28* kl = kps + (size_t)(2 * p) * cap;
    const orb_kp28* kr = kps + (size_t)(2 * p + 1) * cap;
    const int nl = min(counts[2 * p], cap), nr = min(counts[2 * p + 1], cap);
    if (i < nr) {
        const int oct = kr[i].octave;
        if (oct < 0 || oct >= nlevels) {
            atomicOr(flags + p, 2);
            rinfo[(size_t)p * cap + i] = make_int4(1, 0, 0, 0);  // an empty row band: never a candidate
        } else {
            stereo_prep_one(kr, scale, rinfo + (size_t)p * cap, i);
            const int4 ri = rinfo[(size_t)p * cap + i];
            if (ri.x < 0 || ri.y >= rows) atomicOr(flags + p, 1);  // vRowIndices[yi] out of bounds in the reference (:463-473)
        }
    }
    if (i < nl) {
        const int oct = kl[i].octave;
        if (oct < 0 || oct >= nlevels) atomicOr(flags + p, 2);
        if (!(kl[i].y >= 0.0f && kl[i].y < (float)rows)) atomicOr(flags + p, 1);
    }
}

__global__ void __launch_bounds__(256) k_stereo_match_batch(const orb_kp28* __restrict__ kps, const uint8_t* __restrict__ desc,
                                                            const int* __restrict__ counts, int cap, const int4* __restrict__ rinfo,
                                                            float maxD, const int* __restrict__ flags, int* __restrict__ best_r,
                                                            int* __restrict__ best_dist) {
    const int p = blockIdx.y;
    const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nl = min(counts[2 * p], cap), nr = min(counts[2 * p + 1], cap);
    if (warp >= nl || flags[p]) return;
    const size_t L = (size_t)(2 * p) * cap, R = (size_t)(2 * p + 1) * cap;
    stereo_match_one(kps + L, desc + L * 32, rinfo + (size_t)p * cap, desc + R * 32, nr, maxD, best_r + (size_t)p * cap,
                     best_dist + (size_t)p * cap, warp, lane);
}

__global__ void __launch_bounds__(256) k_stereo_refine_batch(const orb_kp28* __restrict__ kps, const int* __restrict__ counts, int cap,
                                                             const int* __restrict__ best_r, const int* __restrict__ best_dist,
                                                             const OrbStereoLevels lv, float mbf, float maxD, float* __restrict__ u_right,
                                                             float* __restrict__ depth, int* __restrict__ sad, int* __restrict__ flags) {
    __shared__ int s_l[8][121];
    __shared__ int s_r[8][11 * 21];
    __shared__ int s_d[8][12];
    const int p = blockIdx.y;
    const int i = (blockIdx.x * 256 + threadIdx.x) >> 5;
    const int nl = min(counts[2 * p], cap);
    if (i >= nl || flags[p]) return;  // a refused pair: its rows are reset by the median kernel
    const size_t L = (size_t)(2 * p) * cap, R = (size_t)(2 * p + 1) * cap, O = (size_t)p * cap;
    stereo_refine_one(kps + L, kps + R, best_r + O, best_dist + O, lv, (size_t)(2 * p), (size_t)(2 * p + 1), mbf, maxD, u_right + O, depth + O,
                      sad + O, flags + p, i, s_l, s_r, s_d);
}

// Median cut of one pair per CTA: median = element size / 2 of the sorted SAD values of the accepted matches, everything
// with !(sad < 1.5f * 1.4f * median) is reset.  Rows of refused pairs (flags) are all reset.
#define STEREO_MED_THREADS 1024
__global__ void __launch_bounds__(STEREO_MED_THREADS) k_stereo_median_batch(const int* __restrict__ counts, int cap, const int* __restrict__ sad,
                                                                            const int* __restrict__ flags, unsigned npad,
                                                                            float* __restrict__ u_right, float* __restrict__ depth) {
    extern __shared__ int keys[];  // npad ints
    __shared__ int nvalid;
    const int p = blockIdx.x, tid = threadIdx.x;
    const int nl = min(counts[2 * p], cap);
    const size_t O = (size_t)p * cap;
    if (flags[p]) {
        for (int i = tid; i < nl; i += STEREO_MED_THREADS) {
            u_right[O + i] = -1.0f;
            depth[O + i] = -1.0f;
        }
        return;
    }
    if (tid == 0) nvalid = 0;
    __syncthreads();
    int mine = 0;
    for (unsigned i = tid; i < npad; i += STEREO_MED_THREADS) {
        const int v = (int)i < nl ? sad[O + i] : -1;
        keys[i] = v >= 0 ? v : 0x7fffffff;
        mine += v >= 0;
    }
    if (mine) atomicAdd(&nvalid, mine);
    __syncthreads();
    for (unsigned k = 2; k <= npad; k <<= 1)
        for (unsigned j = k >> 1; j > 0; j >>= 1) {
            for (unsigned i = tid; i < (npad >> 1); i += STEREO_MED_THREADS) {
                const unsigned lo = ((i & ~(j - 1)) << 1) | (i & (j - 1)), hi = lo | j;
                const bool up = (lo & k) == 0;
                const int x = keys[lo], y = keys[hi];
                if ((x > y) == up) {
                    keys[lo] = y;
                    keys[hi] = x;
                }
            }
            __syncthreads();
        }
    const int nv = nvalid;
    if (nv == 0) return;  // an empty list is left alone (the reference reads past an empty vector)
    const float thDist = 1.5f * 1.4f * (float)keys[nv / 2];
    for (int i = tid; i < nl; i += STEREO_MED_THREADS) {
        const int v = sad[O + i];
        if (v >= 0 && !((float)v < thDist)) {
            u_right[O + i] = -1.0f;
            depth[O + i] = -1.0f;
        }
    }
}

// ------------------------------------------------------------------------------------------
// DBoW2 vocabulary descent: TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup)
// (reference Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1218-1260) with F::distance = FORB::distance
// (FORB.cpp:81-101, the same 256-bit Hamming distance).  16 lanes per feature: lane c takes child c
// (c, c+16, ... for wider nodes), the 16-lane argmin keeps the FIRST minimum in child order (`d < best_d`).
// node_at_level = the node reached when current_level == L - levelsup (0 = root when that level is <= 0,
// -1 when the descent ends above it: the reference then leaves *nid unset).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_voc_descent(const uint8_t* __restrict__ feat, int n, const int* __restrict__ child_off,
                                                     const int* __restrict__ children, const uint8_t* __restrict__ node_desc,
                                                     int nid_level, int* __restrict__ leaf_node, int* __restrict__ node_at_level) {
    const int g = (blockIdx.x * 256 + threadIdx.x) >> 4, sub = threadIdx.x & 15;
    const unsigned gmask = 0xffffu << (threadIdx.x & 16);
    if (g >= n) return;  // a whole 16-lane group leaves together; the shuffles below name only the own group
    const uint4* Fp = reinterpret_cast<const uint4*>(feat) + 2 * (size_t)g;
    const uint4 f0 = __ldg(Fp), f1 = __ldg(Fp + 1);
    const uint4* D = reinterpret_cast<const uint4*>(node_desc);
    int node = 0, level = 0, nid = nid_level <= 0 ? 0 : -1;
    int b = __ldg(child_off), e = __ldg(child_off + 1);
    while (b < e) {  // do { ... } while (!isLeaf()): the root of a non-empty vocabulary has children
        ++level;
        int best = 0x7fffffff, pos = 0x7fffffff;
        for (int c = b + sub; c < e; c += 16) {
            const int id = __ldg(children + c);
            const int d = hamming256(f0, f1, __ldg(D + 2 * (size_t)id), __ldg(D + 2 * (size_t)id + 1));
            if (d < best) {
                best = d;
                pos = c;
            }
        }
#pragma unroll
        for (int s2 = 8; s2 > 0; s2 >>= 1) {
            const int ob = __shfl_xor_sync(gmask, best, s2), op = __shfl_xor_sync(gmask, pos, s2);
            if (ob < best || (ob == best && op < pos)) {
                best = ob;
                pos = op;
            }
        }
        node = __ldg(children + pos);
        if (level == nid_level) nid = node;
        b = __ldg(child_off + node);
        e = __ldg(child_off + node + 1);
    }
    if (sub == 0) {
        leaf_node[g] = node;
        node_at_level[g] = nid;
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
void orbk_count_launch(int n);

cudaError_t orbk_match_all_popc(const uint8_t* q, const int* nq, size_t q_stride, const uint8_t* t, const int* nt, size_t t_stride,
                                int npairs, int max_nq, int* best_idx, int* best_dist, int* second_dist, size_t out_stride,
                                int only_huge, cudaStream_t st) {
    if (npairs <= 0 || max_nq <= 0) return cudaSuccess;
    dim3 grid((max_nq + MA_THREADS - 1) / MA_THREADS, npairs);
    k_match_all<<<grid, MA_THREADS, 0, st>>>(q, nq, q_stride, t, nt, t_stride, best_idx, best_dist, second_dist, out_stride, 1u << 22,
                                             only_huge);
    orbk_count_launch(1);
    return cudaGetLastError();
}

cudaError_t orbk_match_all_mma(const uint8_t* q, const int* nq, size_t q_stride, const uint8_t* t, const int* nt, size_t t_stride,
                               int npairs, int max_nq, int* best_idx, int* best_dist, int* second_dist, size_t out_stride, int kind,
                               int variant, cudaStream_t st);

// Brute-force scan of npairs (query set, train set) pairs.  The tensor-core kernel (match_mma.cu) does every pair whose
// train set has fewer than 2^22 rows; k_match_all's plain path takes the rest (max_nt < 0: not known on the host, both
// kernels are enqueued and each leaves the other's pairs alone).  ORB_B200_MATCH=popc selects the POPC / LOP3 kernel for
// everything (the comparator of the parity tests); ORB_B200_MMA_KIND=f8 the e4m3 form of the contraction.
cudaError_t orbk_match_all(const uint8_t* q, const int* nq, size_t q_stride, const uint8_t* t, const int* nt,
                           size_t t_stride, int npairs, int max_nq, int max_nt, int* best_idx, int* best_dist, int* second_dist,
                           size_t out_stride, cudaStream_t st) {
    if (npairs <= 0 || max_nq <= 0) return cudaSuccess;
    const char* sel = getenv("ORB_B200_MATCH");
    if (sel && !strcmp(sel, "popc"))
        return orbk_match_all_popc(q, nq, q_stride, t, nt, t_stride, npairs, max_nq, best_idx, best_dist, second_dist, out_stride, 0, st);
    const char* kind = getenv("ORB_B200_MMA_KIND");
    const char* var = getenv("ORB_B200_MMA_VARIANT");
    cudaError_t e = orbk_match_all_mma(q, nq, q_stride, t, nt, t_stride, npairs, max_nq, best_idx, best_dist, second_dist, out_stride,
                                       kind && !strcmp(kind, "f8") ? 1 : 0, var ? atoi(var) : 0, st);
    if (e != cudaSuccess) return e;
    if (max_nt < 0 || max_nt >= (1 << 22))
        return orbk_match_all_popc(q, nq, q_stride, t, nt, t_stride, npairs, max_nq, best_idx, best_dist, second_dist, out_stride, 1, st);
    return cudaSuccess;
}

cudaError_t orbk_match_csr(const uint8_t* q, int nq, const uint8_t* t, const int* offsets, const int* cand, int tie_last,
                           int max_dist, int* best_idx, int* best_dist, int* second_dist, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    k_match_csr<<<(nq + 7) / 8, 256, 0, st>>>(q, nq, t, offsets, cand, tie_last, max_dist, best_idx, best_dist, second_dist);
    orbk_count_launch(1);
    return cudaGetLastError();
}

cudaError_t orbk_dist_csr(const uint8_t* q, int nq, const uint8_t* t, const int* offsets, const int* cand, int* out, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    k_dist_csr<<<(nq + 7) / 8, 256, 0, st>>>(q, nq, t, offsets, cand, out);
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

cudaError_t orbk_grid_assign(const orb_kp28* keys, int n, float minX, float minY, float wInv, float hInv, int* cell_start, int* members,
                             cudaStream_t st) {
    k_grid_assign<<<1, 1024, 0, st>>>(keys, n, minX, minY, wInv, hInv, cell_start, members);
    orbk_count_launch(1);
    return cudaGetLastError();
}

cudaError_t orbk_window_count(const orb_kp28* keys, const int* cell_start, const int* members, float minX, float minY, float wInv,
                              float hInv, int nq, const float* qx, const float* qy, const float* qr, const int* qmin, const int* qmax,
                              int* counts, int* offsets, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    k_window<false><<<(nq + 7) / 8, 256, 0, st>>>(keys, nullptr, cell_start, members, minX, minY, wInv, hInv, nq, nullptr, qx, qy, qr,
                                                   qmin, qmax, counts, nullptr, nullptr, nullptr);
    k_scan_counts<<<1, 1024, 0, st>>>(counts, nq, offsets);
    orbk_count_launch(2);
    return cudaGetLastError();
}

cudaError_t orbk_window_fill(const orb_kp28* keys, const uint8_t* tdesc, const int* cell_start, const int* members, float minX,
                             float minY, float wInv, float hInv, int nq, const uint8_t* qdesc, const float* qx, const float* qy,
                             const float* qr, const int* qmin, const int* qmax, const int* offsets, int* cand, int* dist,
                             cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    k_window<true><<<(nq + 7) / 8, 256, 0, st>>>(keys, tdesc, cell_start, members, minX, minY, wInv, hInv, nq, qdesc, qx, qy, qr, qmin,
                                                  qmax, nullptr, offsets, cand, dist);
    orbk_count_launch(1);
    return cudaGetLastError();
}

cudaError_t orbk_stereo_refine(const orb_kp28* kl, int nl, const orb_kp28* kr, const int* best_r, const int* best_dist,
                               const OrbStereoLevels& lv, float mbf, float maxD, float* u_right, float* depth, int* sad, int* flags,
                               cudaStream_t st) {
    if (nl <= 0) return cudaSuccess;
    k_stereo_refine<<<(nl + 7) / 8, 256, 0, st>>>(kl, nl, kr, best_r, best_dist, lv, mbf, maxD, u_right, depth, sad, flags);
    orbk_count_launch(1);
    return cudaGetLastError();
}

cudaError_t orbk_match_mma_init();

// Per-device function attributes of the matcher kernels (dynamic shared memory above 48 KB); orb_matcher_create calls it.
cudaError_t orbk_match_init_device() {
    cudaError_t e = cudaFuncSetAttribute(k_stereo_median_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) return e;
    return orbk_match_mma_init();
}

cudaError_t orbk_stereo_batch(const orb_kp28* kps, const uint8_t* desc, const int* counts, int cap, int npairs, int nlevels, int rows,
                              const float* d_scale, const OrbStereoLevels& lv, float mbf, float maxD, int4* d_rinfo, int* d_best_r,
                              int* d_best_dist, int* d_sad, int* d_flags, float* u_right, float* depth, cudaStream_t st) {
    if (npairs <= 0 || cap <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(d_flags, 0, sizeof(int) * npairs, st);
    if (e != cudaSuccess) return e;
    k_stereo_prep_batch<<<dim3((cap + 255) / 256, npairs), 256, 0, st>>>(kps, counts, cap, nlevels, rows, d_scale, d_rinfo, d_flags);
    k_stereo_match_batch<<<dim3((cap + 7) / 8, npairs), 256, 0, st>>>(kps, desc, counts, cap, d_rinfo, maxD, d_flags, d_best_r, d_best_dist);
    k_stereo_refine_batch<<<dim3((cap + 7) / 8, npairs), 256, 0, st>>>(kps, counts, cap, d_best_r, d_best_dist, lv, mbf, maxD, u_right, depth,
                                                                       d_sad, d_flags);
    unsigned npad = 2;
    while (npad < (unsigned)cap) npad <<= 1;
    if ((size_t)npad * 4 > 160 * 1024) return cudaErrorInvalidValue;  // cap above 40960 keypoints per frame
    k_stereo_median_batch<<<npairs, STEREO_MED_THREADS, (size_t)npad * 4, st>>>(counts, cap, d_sad, d_flags, npad, u_right, depth);
    orbk_count_launch(4);
    return cudaGetLastError();
}

cudaError_t orbk_voc_descent(const uint8_t* feat, int n, const int* child_off, const int* children, const uint8_t* node_desc,
                             int nid_level, int* leaf_node, int* node_at_level, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    k_voc_descent<<<(n + 15) / 16, 256, 0, st>>>(feat, n, child_off, children, node_desc, nid_level, leaf_node, node_at_level);
    orbk_count_launch(1);
    return cudaGetLastError();
}
