// k_match_mma: brute-force Hamming search (BASELINE config 4) on the 5th-generation tensor cores.
//
// Unit: ORBmatcher::DescriptorDistance (reference src/ORBmatcher.cc:896-908) = popcount(a ^ b) over 256 bits.
// With the query bits as a' = 2a - 1 in {-1, +1} and the train bits as b in {0, 1}
//     a' . b = 2 (a . b) - |b|     =>     popcount(a ^ b) = |a| + |b| - 2 (a . b) = |a| - a' . b
// so all nq x nt distances of a pair are one (nq x 256) x (256 x nt) contraction plus a per-query constant.  -1, 0, +1
// are exact in int8 (and in e4m3), and a sum of at most 256 of them is exact in an int32 (and in a binary32)
// accumulator: the distances are bit-identical to the reference's.  The scan semantics on top of them are the shared
// ones (src/ORBmatcher.cc:49-55): best = first minimum in train order, second = second smallest of the multiset.
//
// One CTA = 128 queries (the M of the MMA, one TMEM lane each) of one (query set, train set) pair; it walks the train
// rows in tiles of 128 (the N of the MMA):
//   expand    every thread takes half a train row (one 128-bit load) and spreads its bits to bytes straight into the
//             canonical K-major no-swizzle UMMA layout in shared memory (8 x 16-byte core matrices, PRMT with the bit
//             nibbles as selectors; the bit -> K position map is a fixed permutation, the same for both operands, which
//             a dot product does not see).  Nothing expanded ever touches HBM.
//   mma       one thread issues 6 or 8 tcgen05.mma 128 x 128 x 32 (K = 32 bytes per instruction; 6 when words 6-7 of
//             every train row of the tile are zero -- this fork's descriptors, SURVEY D2 -- checked on the data) into
//             one of two 128-column TMEM accumulator stages and commits to an mbarrier.
//   epilogue  all 8 warps: tcgen05.ld 64 accumulator columns per thread, two IMADs turn two of them into one packed pair
//             of 16-bit keys ((256 - acc) << 6 | column), a min / second-min tournament on packed lanes
//             (VIMNMX.U16x2, 1.2 instructions per distance) reduces the 64 columns, and one 32-bit (distance, train
//             index) key update per 64 columns keeps the running best / second.
// The MMA of tile t runs while the threads do the epilogue of tile t-1 and the expansion of tile t+1; two CTAs share an
// SM (2 x 256 TMEM columns, 2 x 97 KB of shared memory), so the phases of one fill the gaps of the other.
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

#include <utility>

#include "match_kernels.h"

void orbk_count_launch(int n);

#define MM_M 128
#define MM_N 128
#define MM_THREADS 256
#define MM_TILE_BYTES (128 * 256)  // 128 rows x 256 expanded bytes
#define MM_SBO 2048                // bytes between 8-row groups (16 K-chunks of 128 bytes each)
#define MM_LBO 128                 // bytes between the 16-byte K-chunks of a row group (one 8 x 16 B core matrix)
#define MM_TMEM_COLS 256           // two accumulator stages of MM_N columns
#define MM_OFF_BAR (3 * MM_TILE_BYTES)
#define MM_OFF_TMEM (MM_OFF_BAR + 16)
#define MM_OFF_MERGE (MM_OFF_BAR + 32)
#define MM_SMEM (MM_OFF_MERGE + 128 * 8)

__device__ __forceinline__ unsigned mm_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mm_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// Waits for the phase with the given parity.  A wait that never ends would hang the device, so after ~10^7 failed polls
// (seconds; a tile's MMA takes a microsecond) the kernel traps and the launch reports an error instead.
__device__ __forceinline__ void mm_mbar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    for (unsigned spins = 0; !done; ++spins) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void mm_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mm_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] . B[smem]^T, one K = 32-byte step.  KIND 0: kind::i8 (s8 x s8 -> s32); KIND 1: kind::f8f6f4
// (e4m3 x e4m3 -> f32).  Issued by one thread for the whole CTA.
template <int KIND>
__device__ __forceinline__ void mm_mma(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned idesc, unsigned accumulate) {
    if (KIND == 0)
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
}
__device__ __forceinline__ void mm_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 consecutive accumulator columns of this thread's TMEM lane (32 lanes x 32 bit, repeated 32 times along the columns)
__device__ __forceinline__ void mm_tmem_ld32(unsigned taddr, unsigned (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void mm_tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// One 32-bit descriptor word -> 32 bytes (two 16-byte K-chunks), one byte per bit: byte = lut.byte[bit].
// (w >> m) & 0x11111111 leaves bit 4n + m in nibble n; PRMT takes its selectors from the low four nibbles.
__device__ __forceinline__ void mm_expand_word(unsigned w, unsigned lut, uint4& c0, uint4& c1) {
    const unsigned m0 = w & 0x11111111u, m1 = (w >> 1) & 0x11111111u, m2 = (w >> 2) & 0x11111111u, m3 = (w >> 3) & 0x11111111u;
    c0.x = __byte_perm(lut, 0u, m0);
    c0.y = __byte_perm(lut, 0u, m1);
    c0.z = __byte_perm(lut, 0u, m2);
    c0.w = __byte_perm(lut, 0u, m3);
    c1.x = __byte_perm(lut, 0u, m0 >> 16);
    c1.y = __byte_perm(lut, 0u, m1 >> 16);
    c1.z = __byte_perm(lut, 0u, m2 >> 16);
    c1.w = __byte_perm(lut, 0u, m3 >> 16);
}

// words [w0, w0 + nw) of one row into its place of a tile: row r, K-chunk kc at (r / 8) * SBO + kc * LBO + (r % 8) * 16
__device__ __forceinline__ void mm_expand_half_row(uint8_t* tile, int r, int hf, const uint4& v, unsigned lut, int nw) {
    uint8_t* base = tile + (r >> 3) * MM_SBO + (r & 7) * 16 + hf * 8 * MM_LBO;
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < nw) {
            uint4 c0, c1;
            mm_expand_word(w[k], lut, c0, c1);
            *reinterpret_cast<uint4*>(base + (2 * k) * MM_LBO) = c0;
            *reinterpret_cast<uint4*>(base + (2 * k + 1) * MM_LBO) = c1;
        }
    }
}

// What the kernel needs besides the data; the descriptor words arrive as arguments so that a test can probe encodings.
struct MmParams {
    unsigned neg_lo, neg_hi;       // -64 and -64 << 16 as run-time values: keeps the key arithmetic on IMAD (FMA pipe)
    unsigned idesc;                // tcgen05 instruction descriptor (M = 128, N = 128, K-major A and B)
    unsigned long long desc_base;  // shared-memory matrix descriptor without its start address
    unsigned lut_a, lut_b;         // byte values of a query bit (0, 1) / of a train bit (0, 1)
};

template <int KIND>
__global__ void __launch_bounds__(MM_THREADS, 2) k_match_mma(const uint8_t* __restrict__ q, const int* __restrict__ nq, size_t q_stride,
                                                             const uint8_t* __restrict__ t, const int* __restrict__ nt, size_t t_stride,
                                                             int* __restrict__ best_idx, int* __restrict__ best_dist,
                                                             int* __restrict__ second_dist, size_t out_stride, MmParams prm) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int p = blockIdx.y;
    const int nQ = nq[p], nT = nt[p];
    const int q0 = blockIdx.x * MM_M;
    if (q0 >= nQ || nT >= (1 << 22)) return;  // train sets of 2^22 rows and more: the plain kernel (orbk_match_all)
    const uint4* Q = reinterpret_cast<const uint4*>(q + p * q_stride);
    const uint4* T = reinterpret_cast<const uint4*>(t + p * t_stride);
    uint8_t* tileA = smem;
    const unsigned sA = mm_smem_u32(smem);
    const unsigned bar0 = sA + MM_OFF_BAR;
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(smem + MM_OFF_TMEM);
    uint2* merge = reinterpret_cast<uint2*>(smem + MM_OFF_MERGE);
    const int ntiles = (nT + MM_N - 1) / MM_N;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(mm_smem_u32(tmem_slot)), "r"(MM_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        mm_mbar_init(bar0, 1);
        mm_mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // this thread's half row of every operand tile: rows so that 8 consecutive lanes write 8 consecutive 16-byte slots
    const int er = (tid & 7) | ((tid >> 4) << 3), ehf = (tid >> 3) & 1;
    {
        const int row = q0 + er;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (row < nQ) v = __ldg(Q + 2 * (size_t)row + ehf);
        mm_expand_half_row(tileA, er, ehf, v, prm.lut_a, 4);
    }
    uint4 vnext = make_uint4(0, 0, 0, 0);
    if (er < nT) vnext = __ldg(T + 2 * (size_t)er + ehf);
    mm_fence_before();
    __syncthreads();
    mm_fence_after();
    const unsigned tmem = *tmem_slot;
    const unsigned trow = tmem + ((unsigned)(warp & 3) << 21);  // TMEM lane of this thread's warp quarter (lane field << 16)
    const int ch = warp >> 2;                                    // which 64 columns of a tile this thread reduces

    unsigned bestk = 0xffffffffu, seck = 0xffffffffu;

    // min / second-min of 64 accumulator columns of tile `tt`, merged into the running keys
    auto epilogue = [&](int tt) {
        mm_mbar_wait(bar0 + 8 * (tt & 1), (unsigned)(tt >> 1) & 1u);
        mm_fence_after();
        unsigned acc0[32], acc1[32];
        const unsigned ta = trow + (unsigned)((tt & 1) * MM_N + ch * 64);
        mm_tmem_ld32(ta, acc0);
        mm_tmem_ld32(ta + 32, acc1);
        mm_tmem_ld_wait();
        unsigned P[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            if (KIND == 1) {  // binary32 accumulators holding integers: + 1.5 * 2^23 leaves the integer in the low mantissa bits
                acc0[i] = __float_as_uint(__fadd_rn(__uint_as_float(acc0[i]), 12582912.0f)) - 0x4B400000u;
                acc1[i] = __float_as_uint(__fadd_rn(__uint_as_float(acc1[i]), 12582912.0f)) - 0x4B400000u;
            }
            // low lane: column i, high lane: column 32 + i; key = (256 - acc) << 6 | column
            const unsigned c = (256u * 64u + (unsigned)i) | ((256u * 64u + 32u + (unsigned)i) << 16);
            P[i] = acc1[i] * prm.neg_hi + (acc0[i] * prm.neg_lo + c);
        }
        const int lim = nT - (tt * MM_N + ch * 64);  // columns of this slab that are train rows
        if (lim < 64) {
#pragma unroll
            for (int i = 0; i < 32; ++i) P[i] |= (i < lim ? 0u : 0x0000ffffu) | (i + 32 < lim ? 0u : 0xffff0000u);
        }
        // tournament: (lo, hi) = (best, second) of a set of keys, lane-wise
        unsigned lo[16], hi[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            lo[i] = __vminu2(P[2 * i], P[2 * i + 1]);
            hi[i] = __vmaxu2(P[2 * i], P[2 * i + 1]);
        }
#pragma unroll
        for (int n = 8; n >= 1; n >>= 1) {
#pragma unroll
            for (int i = 0; i < n; ++i) {
                const unsigned b = __vminu2(lo[2 * i], lo[2 * i + 1]);
                const unsigned m = __vmaxu2(lo[2 * i], lo[2 * i + 1]);
                hi[i] = __vimin3_u16x2(m, hi[2 * i], hi[2 * i + 1]);
                lo[i] = b;
            }
        }
        // the two lanes against each other; afterwards both halves hold the slab's best / second key
        const unsigned bs = __byte_perm(lo[0], 0u, 0x1032), ss = __byte_perm(hi[0], 0u, 0x1032);
        const unsigned b2 = __vminu2(lo[0], bs);
        const unsigned s2 = __vimin3_u16x2(__vmaxu2(lo[0], bs), hi[0], ss);
        // 32-bit keys: (256 - acc) << 22 | train index (slab * 64 + column); the second key only carries its distance
        const unsigned kb = (b2 & 0xffc0003fu) | ((unsigned)(tt * 2 + ch) << 6);
        const unsigned mx = max(kb, bestk);
        seck = min(seck, min(mx, s2));
        bestk = min(bestk, kb);
    };

    for (int tt = 0; tt < ntiles; ++tt) {
        const int st = tt & 1;
        uint8_t* tileB = smem + (1 + st) * MM_TILE_BYTES;
        // tile tt (loaded one iteration ahead); words 6-7 of every row zero => 6 K-steps
        const uint4 v = vnext;
        {
            const int row = (tt + 1) * MM_N + er;
            vnext = make_uint4(0, 0, 0, 0);
            if (row < nT) vnext = __ldg(T + 2 * (size_t)row + ehf);
        }
        const int upper = __syncthreads_or(ehf ? (int)((v.z | v.w) != 0u) : 0);
        // B[st] was last read by the MMA of tile tt - 2, whose completion every thread waited for in epilogue(tt - 2)
        mm_expand_half_row(tileB, er, ehf, v, prm.lut_b, (ehf && !upper) ? 2 : 4);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mm_fence_before();  // orders the tcgen05.ld of epilogue(tt - 2) (same accumulator stage) before the barrier
        __syncthreads();
        if (tid == 0) {
            mm_fence_after();
            const unsigned long long da = prm.desc_base | (unsigned long long)((sA & 0x3ffffu) >> 4);
            const unsigned long long db = prm.desc_base | (unsigned long long)(((sA + (1 + st) * MM_TILE_BYTES) & 0x3ffffu) >> 4);
            const int ksteps = upper ? 8 : 6;
            for (int k = 0; k < ksteps; ++k)  // one K-step = 32 expanded bytes = two K-chunks = 2 * LBO bytes further on
                mm_mma<KIND>(tmem + st * MM_N, da + (unsigned long long)((k * 2 * MM_LBO) >> 4), db + (unsigned long long)((k * 2 * MM_LBO) >> 4),
                             prm.idesc, k > 0 ? 1u : 0u);
            mm_commit(bar0 + 8 * st);
        }
        if (tt > 0) epilogue(tt - 1);
    }
    if (ntiles > 0) epilogue(ntiles - 1);

    // the two column halves of a row meet in shared memory; warps 0-3 write the row's result
    mm_fence_before();
    if (ch == 1) merge[(warp & 3) * 32 + lane] = make_uint2(bestk, seck);
    __syncthreads();
    if (ch == 0) {
        const uint2 o = merge[warp * 32 + lane];
        const unsigned mx = max(bestk, o.x);
        seck = min(min(seck, o.y), mx);
        bestk = min(bestk, o.x);
        const int qi = q0 + warp * 32 + lane;
        if (qi < nQ) {
            const uint4 a0 = __ldg(Q + 2 * (size_t)qi), a1 = __ldg(Q + 2 * (size_t)qi + 1);
            const int na = __popc(a0.x) + __popc(a0.y) + __popc(a0.z) + __popc(a0.w) + __popc(a1.x) + __popc(a1.y) + __popc(a1.z) + __popc(a1.w);
            int idx = -1, bd = INT_MAX, sd = INT_MAX;
            if ((bestk >> 22) <= 512u) {
                idx = (int)(bestk & 0x3fffffu);
                bd = (int)(bestk >> 22) - 256 + na;
            }
            if ((seck >> 22) <= 512u) sd = (int)(seck >> 22) - 256 + na;
            best_idx[p * out_stride + qi] = idx;
            best_dist[p * out_stride + qi] = bd;
            second_dist[p * out_stride + qi] = sd;
        }
    }
    if (warp == 0) {
        mm_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(MM_TMEM_COLS) : "memory");
    }
}

// Train sets the tensor-core kernel leaves alone (2^22 rows and more): k_match_all's plain path, see match_kernels.cu.
cudaError_t orbk_match_all_popc(const uint8_t* q, const int* nq, size_t q_stride, const uint8_t* t, const int* nt, size_t t_stride,
                                int npairs, int max_nq, int* best_idx, int* best_dist, int* second_dist, size_t out_stride,
                                int only_huge, cudaStream_t st);

cudaError_t orbk_match_mma_init() {
    cudaError_t e = cudaFuncSetAttribute(k_match_mma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, MM_SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_match_mma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, MM_SMEM);
}

// kind: 0 = kind::i8, 1 = kind::f8f6f4 (e4m3).  variant: 0 = the documented encoding; other values exist for
// tools/probes/mma_probe.py only (1: leading / stride byte offsets swapped, 2: descriptor version bits clear).
cudaError_t orbk_match_all_mma(const uint8_t* q, const int* nq, size_t q_stride, const uint8_t* t, const int* nt, size_t t_stride,
                               int npairs, int max_nq, int* best_idx, int* best_dist, int* second_dist, size_t out_stride, int kind,
                               int variant, cudaStream_t st) {
    if (npairs <= 0 || max_nq <= 0) return cudaSuccess;
    static const cudaError_t init = orbk_match_mma_init();
    if (init != cudaSuccess) return init;
    MmParams prm;
    prm.neg_lo = (unsigned)-64;
    prm.neg_hi = (unsigned)(-64 * 65536);
    unsigned lbo = MM_LBO, sbo = MM_SBO;
    if (variant == 1) std::swap(lbo, sbo);
    // matrix descriptor: start address >> 4 [0,14), leading byte offset >> 4 [16,30), stride byte offset >> 4 [32,46),
    // descriptor version 1 [46,48), base offset 0, layout type 0 = no swizzle [61,64)
    prm.desc_base = ((unsigned long long)(lbo >> 4) << 16) | ((unsigned long long)(sbo >> 4) << 32) | (variant == 2 ? 0ull : (1ull << 46));
    // instruction descriptor: D format [4,6), A format [7,10), B format [10,13), A / B major [15], [16] = 0 (K-major),
    // N >> 3 [17,23), M >> 4 [24,29)
    const unsigned shape = ((unsigned)(MM_N >> 3) << 17) | ((unsigned)(MM_M >> 4) << 24);
    if (kind == 0) {
        prm.idesc = (2u << 4) | (1u << 7) | (1u << 10) | shape;  // s32 accumulators, signed 8-bit A and B
        prm.lut_a = 0x000001ffu;                                  // query bit 0 -> -1, 1 -> +1
        prm.lut_b = 0x00000100u;                                  // train bit 0 -> 0, 1 -> 1
    } else {
        prm.idesc = (1u << 4) | shape;  // f32 accumulators, e4m3 A and B
        prm.lut_a = 0x000038b8u;        // e4m3: 0xb8 = -1.0, 0x38 = +1.0
        prm.lut_b = 0x00003800u;
    }
    dim3 grid((max_nq + MM_M - 1) / MM_M, npairs);
    if (kind == 0)
        k_match_mma<0><<<grid, MM_THREADS, MM_SMEM, st>>>(q, nq, q_stride, t, nt, t_stride, best_idx, best_dist, second_dist, out_stride, prm);
    else
        k_match_mma<1><<<grid, MM_THREADS, MM_SMEM, st>>>(q, nq, q_stride, t, nt, t_stride, best_idx, best_dist, second_dist, out_stride, prm);
    orbk_count_launch(1);
    return cudaGetLastError();
}
