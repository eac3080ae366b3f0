// Brute-force Hamming search (BASELINE config 4) on the 5th-generation tensor cores.  Three generations of the same
// contraction live here: k_match_mma3 (persistent, warp-specialised, bias K-step: the kernel orbk_match_all launches),
// k_match_mma2 (warp-specialised, one CTA per 256 queries) and k_match_mma (the first form, described below); the two older
// ones stay selectable (ORB_B200_MMA_VARIANT) as comparators for the numbers in DESIGN.md section 4 and 9.
//
// Unit: ORBmatcher::DescriptorDistance (reference src/ORBmatcher.cc:896-908) = popcount(a ^ b) over 256 bits.
// With the query bits as a' = 2a - 1 in {-1, +1} and the train bits as b in {0, 1}
//     a' . b = 2 (a . b) - |b|     =>     popcount(a ^ b) = |a| + |b| - 2 (a . b) = |a| - a' . b
// so all nq x nt distances of a pair are one (nq x 256) x (256 x nt) contraction plus a per-query constant.  -1, 0, +1
// are exact in int8 (and in e4m3), and a sum of at most 256 of them is exact in an int32 (and in a binary32)
// accumulator: the distances are bit-identical to the reference's.  The scan semantics on top of them are the shared
// ones (src/ORBmatcher.cc:49-55): best = first minimum in train order, second = second smallest of the multiset.
//
// k_match_mma, the first form.  One CTA = 128 queries (the M of the MMA, one TMEM lane each) of one (query set, train set)
// pair; it walks the train rows in tiles of 128 (the N of the MMA):
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
// The same with a suspend-time hint (nanoseconds): the roles that run ahead of the bottleneck park in the barrier unit instead
// of re-issuing the poll, which would take issue slots from the epilogue warps on their scheduler.
__device__ __forceinline__ void mm_mbar_wait_parked(unsigned bar, unsigned parity, unsigned hint_ns) {
    unsigned done = 0;
    for (unsigned spins = 0; !done; ++spins) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(hint_ns)
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
// The same with the descriptors' words apart: only the low word (start address, leading byte offset) changes from K-step to
// K-step and from stage to stage, so the issuing thread keeps 32-bit values and no 64-bit arithmetic between two MMAs.
template <int KIND, int ACC>
__device__ __forceinline__ void mm_mma_w(unsigned tmem_d, unsigned alo, unsigned blo, unsigned hi, unsigned idesc) {
    if (KIND == 0)
        asm volatile(
            "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %5, 0;\nmov.b64 da, {%1, %3};\nmov.b64 db, {%2, %3};\n"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %4, p;\n}\n" ::"r"(tmem_d), "r"(alo), "r"(blo), "r"(hi), "r"(idesc), "n"(ACC)
            : "memory");
    else
        asm volatile(
            "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %5, 0;\nmov.b64 da, {%1, %3};\nmov.b64 db, {%2, %3};\n"
            "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %4, p;\n}\n" ::"r"(tmem_d), "r"(alo), "r"(blo), "r"(hi), "r"(idesc), "n"(ACC)
            : "memory");
}
// accumulate one more K-step whose two descriptors have high words of their own
template <int KIND>
__device__ __forceinline__ void mm_mma_x(unsigned tmem_d, unsigned alo, unsigned blo, unsigned ahi, unsigned bhi, unsigned idesc) {
    if (KIND == 0)
        asm volatile(
            "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, 1, 0;\nmov.b64 da, {%1, %3};\nmov.b64 db, {%2, %4};\n"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %5, p;\n}\n" ::"r"(tmem_d), "r"(alo), "r"(blo), "r"(ahi), "r"(bhi), "r"(idesc)
            : "memory");
    else
        asm volatile(
            "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, 1, 0;\nmov.b64 da, {%1, %3};\nmov.b64 db, {%2, %4};\n"
            "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %5, p;\n}\n" ::"r"(tmem_d), "r"(alo), "r"(blo), "r"(ahi), "r"(bhi), "r"(idesc)
            : "memory");
}
__device__ __forceinline__ bool mm_elect_one() {
    unsigned pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
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

// One 32-bit descriptor word -> 32 bytes (two 16-byte K-chunks), one byte per bit.  PRMT reads four selector nibbles from
// the low 16 bits of its third operand; a nibble's low three bits pick one of the eight bytes of (a, b) and its top bit asks
// for sign replication.  With the top bits cleared (w & 0x77777777) the tables {v0 v1 v0 v1 | v0 v1 v0 v1}, {v0 v0 v1 v1 | ...}
// and {v0 v0 v0 v0 | v1 v1 v1 v1} read bit 0, 1 and 2 of every nibble without isolating it first; bit 3 is bit 2 of w >> 1.
// 13 integer instructions per word (2 LOP3, 3 SHF, 8 PRMT).  Byte j of c0 / c1 is NOT bit j: queries and train rows go through
// the same permutation of the K positions, which a dot product does not see.
struct MmLut {
    unsigned t0, t1, t2a, t2b;
};
__device__ __forceinline__ MmLut mm_lut(unsigned lut) {
    MmLut L;
    L.t0 = __byte_perm(lut, 0u, 0x1010);
    L.t1 = __byte_perm(lut, 0u, 0x1100);
    L.t2a = __byte_perm(lut, 0u, 0x0000);
    L.t2b = __byte_perm(lut, 0u, 0x1111);
    return L;
}
__device__ __forceinline__ void mm_expand_word(unsigned w, const MmLut& L, uint4& c0, uint4& c1) {
    const unsigned x = w & 0x77777777u, y = (w >> 1) & 0x77777777u;
    const unsigned xh = x >> 16, yh = y >> 16;
    c0.x = __byte_perm(L.t0, L.t0, x);
    c0.y = __byte_perm(L.t1, L.t1, x);
    c0.z = __byte_perm(L.t2a, L.t2b, x);
    c0.w = __byte_perm(L.t2a, L.t2b, y);
    c1.x = __byte_perm(L.t0, L.t0, xh);
    c1.y = __byte_perm(L.t1, L.t1, xh);
    c1.z = __byte_perm(L.t2a, L.t2b, xh);
    c1.w = __byte_perm(L.t2a, L.t2b, yh);
}

// words [w0, w0 + nw) of one row into its place of a tile: row r, K-chunk kc at (r / 8) * SBO + kc * LBO + (r % 8) * 16
__device__ __forceinline__ void mm_expand_half_row(uint8_t* tile, int r, int hf, const uint4& v, unsigned lut_bytes, int nw) {
    const MmLut lut = mm_lut(lut_bytes);
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
    unsigned park_ns;              // suspend-time hint of the producer / MMA waits (0: plain polling)
    unsigned shl16;                // 65536 as a run-time value (IMAD instead of a shift + add on the ALU pipe)
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

// ------------------------------------------------------------------------------------------
// k_match_mma2: the same contraction, warp-specialised.  One CTA per SM = 256 queries (two 128-row A tiles that share
// every expanded train tile) and three roles that meet only at mbarriers:
//   producers  4 warps, one train row per thread and tile: 2 x 128-bit loads, expansion into one of three B stages,
//              fence.proxy.async, arrive on full[stage]; wait on empty[stage] (committed by the MMAs that read it)
//   MMA        one thread: waits for full[stage] and for the epilogue to have drained the accumulator stage, issues
//              2 x (6 | 8) tcgen05.mma 128 x 128 x 32 (A tile 0 and A tile 1 against the same B stage), commits to
//              empty[stage] and to tfull[accumulator stage]
//   epilogue   8 warps, one query row per thread (all 128 columns of a tile: two 64-column slabs): tcgen05.ld, packed key
//              tournament, running keys; releases the accumulator stage as soon as its second slab is in registers
// No block-wide barrier inside the tile loop; the train-tile expansion is paid once per 256 queries instead of once per 128.
// ------------------------------------------------------------------------------------------
#define M2_PROD_WARPS 4
#define M2_THREADS(EW) (((EW) + M2_PROD_WARPS + 1) * 32)
#define M2_STAGES 3
#define M2_OFF_B (2 * MM_TILE_BYTES)
#define M2_OFF_BAR (M2_OFF_B + M2_STAGES * MM_TILE_BYTES)  // full[3], empty[3], tfull[2], tempty[2]
#define M2_OFF_TMEM (M2_OFF_BAR + 10 * 8)
#define M2_OFF_FLAGS (M2_OFF_TMEM + 8)
#define M2_OFF_MERGE (M2_OFF_FLAGS + 32)  // 256 rows x (best, second) of the upper column half (16 epilogue warps)
#define M2_SMEM (M2_OFF_MERGE + 2 * MM_M * 8)

__device__ __forceinline__ void mm_mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// 64 accumulator columns (two tcgen05.ld of 32) -> 32 packed key pairs; low lane: column i, high lane: column 32 + i,
// key = (256 - acc) << 6 | column; columns >= lim (past the last train row) become 0xffff
template <int KIND>
__device__ __forceinline__ void mm_slab_keys(unsigned (&acc0)[32], unsigned (&acc1)[32], const MmParams& prm, int lim, unsigned (&P)[32]) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        if (KIND == 1) {  // binary32 accumulators holding integers: + 1.5 * 2^23 leaves the integer in the low mantissa bits
            acc0[i] = __float_as_uint(__fadd_rn(__uint_as_float(acc0[i]), 12582912.0f)) - 0x4B400000u;
            acc1[i] = __float_as_uint(__fadd_rn(__uint_as_float(acc1[i]), 12582912.0f)) - 0x4B400000u;
        }
        const unsigned c = (256u * 64u + (unsigned)i) | ((256u * 64u + 32u + (unsigned)i) << 16);
        P[i] = acc1[i] * prm.neg_hi + (acc0[i] * prm.neg_lo + c);
    }
    if (lim < 64) {
#pragma unroll
        for (int i = 0; i < 32; ++i) P[i] |= (i < lim ? 0u : 0x0000ffffu) | (i + 32 < lim ? 0u : 0xffff0000u);
    }
}

// BK = 1: the accumulators already are the keys (bias K-step, see k_match_mma3): one IMAD packs two of them
template <int KIND, int BK>
__device__ __forceinline__ void mm_slab_keys_any(unsigned (&acc0)[32], unsigned (&acc1)[32], const MmParams& prm, int lim, unsigned (&P)[32]) {
    if (BK) {
#pragma unroll
        for (int i = 0; i < 32; ++i) P[i] = acc1[i] * prm.shl16 + acc0[i];
        if (lim < 64) {
#pragma unroll
            for (int i = 0; i < 32; ++i) P[i] |= (i < lim ? 0u : 0x0000ffffu) | (i + 32 < lim ? 0u : 0xffff0000u);
        }
    } else {
        mm_slab_keys<KIND>(acc0, acc1, prm, lim, P);
    }
}

// The same result in two passes, 0.9 instead of 1.3 ALU-pipe instructions per key pair.  Pass 1: lane-wise minimum M of the
// 32 words (VIMNMX3).  Pass 2: lane-wise minimum of key + ~M (16-bit wrap-around add, done by VIADDMNMX together with the
// running minimum): the lane's own minimum becomes 0xffff and drops out -- keys of a lane are distinct, except for the
// 0xffff of masked columns, which can only be the minimum when the whole lane is masked -- and every other key becomes
// key - M - 1 >= 0, order kept.  The lane's second key is that minimum + M + 1, again with wrap-around adds (all-masked lane:
// 0xffff + 0xffff + 1 = 0xffff).
__device__ __forceinline__ void mm_min2(const unsigned (&P)[32], unsigned& b2, unsigned& s2) {
    unsigned m[11];
#pragma unroll
    for (int i = 0; i < 10; ++i) m[i] = __vimin3_u16x2(P[3 * i], P[3 * i + 1], P[3 * i + 2]);
    m[10] = __vminu2(P[30], P[31]);
    const unsigned m0 = __vimin3_u16x2(m[0], m[1], m[2]), m1 = __vimin3_u16x2(m[3], m[4], m[5]), m2 = __vimin3_u16x2(m[6], m[7], m[8]);
    const unsigned M = __vminu2(__vimin3_u16x2(m0, m1, m2), __vminu2(m[9], m[10]));
    const unsigned nM = ~M;
    unsigned R[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
#pragma unroll
    for (int i = 0; i < 32; ++i) R[i & 3] = __viaddmin_u16x2(P[i], nM, R[i & 3]);
    const unsigned r = __vminu2(__vimin3_u16x2(R[0], R[1], R[2]), R[3]);
    const unsigned S = __viaddmin_u16x2(__viaddmin_u16x2(r, M, 0xffffffffu), 0x00010001u, 0xffffffffu);
    const unsigned bs = __byte_perm(M, 0u, 0x1032), ss = __byte_perm(S, 0u, 0x1032);
    b2 = __vminu2(M, bs);
    s2 = __vimin3_u16x2(__vmaxu2(M, bs), S, ss);
}

// min / second-min tournament over 32 packed key pairs; both halves of b2 / s2 hold the slab's best / second 16-bit key
__device__ __forceinline__ void mm_tournament(const unsigned (&P)[32], unsigned& b2, unsigned& s2) {
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
    const unsigned bs = __byte_perm(lo[0], 0u, 0x1032), ss = __byte_perm(hi[0], 0u, 0x1032);
    b2 = __vminu2(lo[0], bs);
    s2 = __vimin3_u16x2(__vmaxu2(lo[0], bs), hi[0], ss);
}

// mm_unit_keys in two parts, so that a caller can put other straight-line work between the arithmetic and the (rare) masking
template <int KIND>
__device__ __forceinline__ void mm_unit_keys_raw(unsigned (&acc)[32], const MmParams& prm, unsigned (&P)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if (KIND == 1) {
            acc[i] = __float_as_uint(__fadd_rn(__uint_as_float(acc[i]), 12582912.0f)) - 0x4B400000u;
            acc[i + 16] = __float_as_uint(__fadd_rn(__uint_as_float(acc[i + 16]), 12582912.0f)) - 0x4B400000u;
        }
        const unsigned c = (256u * 64u + (unsigned)i) | ((256u * 64u + 16u + (unsigned)i) << 16);
        P[i] = acc[i + 16] * prm.neg_hi + (acc[i] * prm.neg_lo + c);
    }
}
__device__ __forceinline__ void mm_unit_mask(unsigned (&P)[16], int lim) {
    if (lim < 32) {
#pragma unroll
        for (int i = 0; i < 16; ++i) P[i] |= (i < lim ? 0u : 0x0000ffffu) | (i + 16 < lim ? 0u : 0xffff0000u);
    }
}

// 32 accumulator columns (one tcgen05.ld) -> 16 packed key pairs; low lane: column i, high lane: column 16 + i
template <int KIND>
__device__ __forceinline__ void mm_unit_keys(unsigned (&acc)[32], const MmParams& prm, int lim, unsigned (&P)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if (KIND == 1) {
            acc[i] = __float_as_uint(__fadd_rn(__uint_as_float(acc[i]), 12582912.0f)) - 0x4B400000u;
            acc[i + 16] = __float_as_uint(__fadd_rn(__uint_as_float(acc[i + 16]), 12582912.0f)) - 0x4B400000u;
        }
        const unsigned c = (256u * 64u + (unsigned)i) | ((256u * 64u + 16u + (unsigned)i) << 16);
        P[i] = acc[i + 16] * prm.neg_hi + (acc[i] * prm.neg_lo + c);
    }
    if (lim < 32) {
#pragma unroll
        for (int i = 0; i < 16; ++i) P[i] |= (i < lim ? 0u : 0x0000ffffu) | (i + 16 < lim ? 0u : 0xffff0000u);
    }
}

__device__ __forceinline__ void mm_tournament16(const unsigned (&P)[16], unsigned& b2, unsigned& s2) {
    unsigned lo[8], hi[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        lo[i] = __vminu2(P[2 * i], P[2 * i + 1]);
        hi[i] = __vmaxu2(P[2 * i], P[2 * i + 1]);
    }
#pragma unroll
    for (int n = 4; n >= 1; n >>= 1) {
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const unsigned b = __vminu2(lo[2 * i], lo[2 * i + 1]);
            const unsigned m = __vmaxu2(lo[2 * i], lo[2 * i + 1]);
            hi[i] = __vimin3_u16x2(m, hi[2 * i], hi[2 * i + 1]);
            lo[i] = b;
        }
    }
    const unsigned bs = __byte_perm(lo[0], 0u, 0x1032), ss = __byte_perm(hi[0], 0u, 0x1032);
    b2 = __vminu2(lo[0], bs);
    s2 = __vimin3_u16x2(__vmaxu2(lo[0], bs), hi[0], ss);
}

// EW = 8 (shipped): one query row per epilogue thread (all 128 columns of a tile).  EW = 16: two threads per query row, 64
// columns each in units of 32, 80 registers per thread so that 21 warps fit the register file; measured on B200: 3.59 T
// pairs/s against 3.83 T with EW = 8 (more warps do not help: the epilogue is bound by the ALU pipe's issue rate, and
// the 32-column units cost more instructions per distance), kept as a comparator.
template <int KIND, int EW>
__global__ void __launch_bounds__(M2_THREADS(EW), 1) k_match_mma2(const uint8_t* __restrict__ q, const int* __restrict__ nq, size_t q_stride,
                                                              const uint8_t* __restrict__ t, const int* __restrict__ nt, size_t t_stride,
                                                              int* __restrict__ best_idx, int* __restrict__ best_dist,
                                                              int* __restrict__ second_dist, size_t out_stride, MmParams prm) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int p = blockIdx.y;
    const int nQ = nq[p], nT = nt[p];
    const int q0 = blockIdx.x * 2 * MM_M;
    if (q0 >= nQ || nT >= (1 << 22)) return;
    const uint4* Q = reinterpret_cast<const uint4*>(q + p * q_stride);
    const uint4* T = reinterpret_cast<const uint4*>(t + p * t_stride);
    const unsigned sA = mm_smem_u32(smem);
    const unsigned bFull = sA + M2_OFF_BAR, bEmpty = bFull + 8 * M2_STAGES, bTfull = bEmpty + 8 * M2_STAGES, bTempty = bTfull + 16;
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(smem + M2_OFF_TMEM);
    volatile int* flags = reinterpret_cast<volatile int*>(smem + M2_OFF_FLAGS);
    const int ntiles = (nT + MM_N - 1) / MM_N;

    if (warp == EW + M2_PROD_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(mm_smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < M2_STAGES; ++s) {
            mm_mbar_init(bFull + 8 * s, M2_PROD_WARPS * 32);
            mm_mbar_init(bEmpty + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mm_mbar_init(bTfull + 8 * a, 1);
            mm_mbar_init(bTempty + 8 * a, EW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // both A tiles (256 query rows x 2 halves), +-1 bytes
    for (int item = tid; item < 4 * MM_M; item += M2_THREADS(EW)) {
        const int er = (item & 7) | ((item >> 4) << 3), ehf = (item >> 3) & 1;
        const int row = q0 + er;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (row < nQ) v = __ldg(Q + 2 * (size_t)row + ehf);
        mm_expand_half_row(smem + (er >> 7) * MM_TILE_BYTES, er & 127, ehf, v, prm.lut_a, 4);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mm_fence_before();
    __syncthreads();
    mm_fence_after();
    const unsigned tmem = *tmem_slot;

    if (EW == 16 && warp < EW) {
        // ---------------- epilogue, 16 warps: (A tile, lane quarter, column half) per warp ----------------
        const int tileA = (warp >> 2) & 1, half = warp >> 3;
        const unsigned trow = tmem + ((unsigned)(warp & 3) << 21) + (unsigned)(tileA * MM_N + half * 64);
        unsigned bestk = 0xffffffffu, seck = 0xffffffffu;
        auto merge = [&](unsigned b2, unsigned s2, int unit) {  // train index = unit * 32 + column
            const unsigned kb = (b2 & 0xffc0001fu) | ((unsigned)unit << 5);
            const unsigned mx = max(kb, bestk);
            seck = min(seck, min(mx, s2));
            bestk = min(bestk, kb);
        };
        unsigned acc[32];
        if (ntiles > 0) {
            mm_mbar_wait(bTfull, 0u);
            mm_fence_after();
            mm_tmem_ld32(trow, acc);
        }
        for (int tt = 0; tt < ntiles; ++tt) {
            const int a = tt & 1;
            const unsigned ta = trow + (unsigned)(a * 2 * MM_N);
            unsigned P[16], b2, s2;
            mm_tmem_ld_wait();
            mm_unit_keys<KIND>(acc, prm, nT - tt * MM_N - half * 64, P);
            mm_tmem_ld32(ta + 32, acc);
            mm_tournament16(P, b2, s2);
            merge(b2, s2, tt * 4 + half * 2);
            mm_tmem_ld_wait();
            mm_fence_before();
            __syncwarp();
            if (lane == 0) mm_mbar_arrive(bTempty + 8 * a);
            mm_unit_keys<KIND>(acc, prm, nT - tt * MM_N - half * 64 - 32, P);
            if (tt + 1 < ntiles) {
                mm_mbar_wait(bTfull + 8 * (a ^ 1), (unsigned)((tt + 1) >> 1) & 1u);
                mm_fence_after();
                mm_tmem_ld32(trow + (unsigned)((a ^ 1) * 2 * MM_N), acc);
            }
            mm_tournament16(P, b2, s2);
            merge(b2, s2, tt * 4 + half * 2 + 1);
        }
        // the two column halves of a row meet in shared memory (named barrier 2: the epilogue warps only)
        uint2* mrg = reinterpret_cast<uint2*>(smem + M2_OFF_MERGE);
        const int rowInCta = tileA * MM_M + (warp & 3) * 32 + lane;
        if (half == 1) mrg[rowInCta] = make_uint2(bestk, seck);
        asm volatile("bar.sync 2, %0;" ::"n"(EW * 32) : "memory");
        if (half == 0) {
            const uint2 o = mrg[rowInCta];
            const unsigned mx = max(bestk, o.x);
            seck = min(min(seck, o.y), mx);
            bestk = min(bestk, o.x);
            const int qi = q0 + rowInCta;
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
        mm_fence_before();
    } else if (warp < EW) {
        // ---------------- epilogue, 8 warps: one query row per thread ----------------
        const int tileA = warp >> 2;
        const unsigned trow = tmem + ((unsigned)(warp & 3) << 21) + (unsigned)(tileA * MM_N);
        unsigned bestk = 0xffffffffu, seck = 0xffffffffu;
        auto merge = [&](unsigned b2, unsigned s2, int slab) {
            // 32-bit keys: (256 - acc) << 22 | train index (slab * 64 + column); the second key only carries its distance
            const unsigned kb = (b2 & 0xffc0003fu) | ((unsigned)slab << 6);
            const unsigned mx = max(kb, bestk);
            seck = min(seck, min(mx, s2));
            bestk = min(bestk, kb);
        };
        // The accumulator loads run one slab ahead of the arithmetic: the tcgen05.ld of the next 64 columns is in flight
        // while the tournament of the previous 64 runs.
        unsigned acc0[32], acc1[32];
        if (ntiles > 0) {
            mm_mbar_wait(bTfull, 0u);
            mm_fence_after();
            mm_tmem_ld32(trow, acc0);
            mm_tmem_ld32(trow + 32, acc1);
        }
        for (int tt = 0; tt < ntiles; ++tt) {
            const int a = tt & 1;
            const unsigned ta = trow + (unsigned)(a * 2 * MM_N);
            unsigned P[32], b2, s2;
            mm_tmem_ld_wait();  // columns 0..63 of tile tt
            mm_slab_keys<KIND>(acc0, acc1, prm, nT - tt * MM_N, P);
            mm_tmem_ld32(ta + 64, acc0);
            mm_tmem_ld32(ta + 96, acc1);
            mm_tournament(P, b2, s2);
            merge(b2, s2, tt * 2);
            mm_tmem_ld_wait();  // columns 64..127: this warp is done with the accumulator stage, the MMAs of tile tt + 2 may overwrite it
            mm_fence_before();
            __syncwarp();
            if (lane == 0) mm_mbar_arrive(bTempty + 8 * a);
            mm_slab_keys<KIND>(acc0, acc1, prm, nT - tt * MM_N - 64, P);
            if (tt + 1 < ntiles) {
                mm_mbar_wait(bTfull + 8 * (a ^ 1), (unsigned)((tt + 1) >> 1) & 1u);
                mm_fence_after();
                const unsigned tn = trow + (unsigned)((a ^ 1) * 2 * MM_N);
                mm_tmem_ld32(tn, acc0);
                mm_tmem_ld32(tn + 32, acc1);
            }
            mm_tournament(P, b2, s2);
            merge(b2, s2, tt * 2 + 1);
        }
        const int qi = q0 + tileA * MM_M + (warp & 3) * 32 + lane;
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
        mm_fence_before();
    } else if (warp < EW + M2_PROD_WARPS) {
        // ---------------- producers: one train row per thread and tile ----------------
        const int r = tid - EW * 32;
        uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0;
        if (r < nT) {
            n0 = __ldg(T + 2 * (size_t)r);
            n1 = __ldg(T + 2 * (size_t)r + 1);
        }
        for (int tt = 0; tt < ntiles; ++tt) {
            const int s = tt % M2_STAGES;
            const uint4 v0 = n0, v1 = n1;
            {
                const int row = (tt + 1) * MM_N + r;
                n0 = n1 = make_uint4(0, 0, 0, 0);
                if (row < nT) {
                    n0 = __ldg(T + 2 * (size_t)row);
                    n1 = __ldg(T + 2 * (size_t)row + 1);
                }
            }
            // words 6-7 of every row of the tile zero => the MMAs skip the last two K-steps and nobody expands them
            unsigned upper;
            asm volatile(
                "{\n.reg .pred p, q;\nsetp.ne.u32 p, %1, 0;\nbar.red.or.pred q, 1, %2, p;\nselp.u32 %0, 1, 0, q;\n}\n"
                : "=r"(upper)
                : "r"(v1.z | v1.w), "n"(M2_PROD_WARPS * 32)
                : "memory");
            mm_mbar_wait(bEmpty + 8 * s, ((unsigned)(tt / M2_STAGES) & 1u) ^ 1u);
            uint8_t* tileB = smem + M2_OFF_B + s * MM_TILE_BYTES;
            mm_expand_half_row(tileB, r, 0, v0, prm.lut_b, 4);
            mm_expand_half_row(tileB, r, 1, v1, prm.lut_b, upper ? 4 : 2);
            if (r == 0) flags[s] = (int)upper;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mm_mbar_arrive(bFull + 8 * s);
        }
    } else if (lane == 0) {
        // ---------------- MMA issuer ----------------
        const unsigned long long dA0 = prm.desc_base | (unsigned long long)((sA & 0x3ffffu) >> 4);
        const unsigned long long dA1 = prm.desc_base | (unsigned long long)(((sA + MM_TILE_BYTES) & 0x3ffffu) >> 4);
        for (int tt = 0; tt < ntiles; ++tt) {
            const int s = tt % M2_STAGES, a = tt & 1;
            mm_mbar_wait(bFull + 8 * s, (unsigned)(tt / M2_STAGES) & 1u);
            mm_mbar_wait(bTempty + 8 * a, ((unsigned)(tt >> 1) & 1u) ^ 1u);
            mm_fence_after();
            const unsigned long long dB = prm.desc_base | (unsigned long long)(((sA + M2_OFF_B + s * MM_TILE_BYTES) & 0x3ffffu) >> 4);
            const int ksteps = flags[s] ? 8 : 6;
            const unsigned d0 = tmem + (unsigned)(a * 2 * MM_N);
            for (int k = 0; k < ksteps; ++k) {
                const unsigned long long o = (unsigned long long)((k * 2 * MM_LBO) >> 4);
                mm_mma<KIND>(d0, dA0 + o, dB + o, prm.idesc, k > 0 ? 1u : 0u);
            }
            for (int k = 0; k < ksteps; ++k) {
                const unsigned long long o = (unsigned long long)((k * 2 * MM_LBO) >> 4);
                mm_mma<KIND>(d0 + MM_N, dA1 + o, dB + o, prm.idesc, k > 0 ? 1u : 0u);
            }
            mm_commit(bEmpty + 8 * s);
            mm_commit(bTfull + 8 * a);
        }
    }
    __syncthreads();
    if (warp == EW + M2_PROD_WARPS) {
        mm_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// k_match_mma3 (the kernel that ships): k_match_mma2 made persistent, and what the timeline of its hand-overs
// (tools/probes/mma_timeline.py) and knock-out timings (tools/probes/match_sweep.py on a -DORB_B200_MMA_KNOCKOUT build) asked for:
//  * one CTA per SM walks the work items (pair, block of 256 queries) with stride gridDim.x; the roles keep running barrier
//    counters across the items, the two A tiles of the NEXT item are expanded into a second A buffer while the MMAs of the
//    current one still read the first; TMEM allocation, barrier set-up and the fill / drain of the pipeline are paid once per
//    CTA instead of once per 256 queries;
//  * two MMA-issuing warps, one per accumulator stage: tcgen05.mma blocks its thread while the tensor core's queue is full and
//    a barrier poll takes ~150 cycles even when it succeeds at once, so a single issuing warp left the tensor pipe idle 40 %
//    of the time; the issuing warps stay converged (elect.sync), which lets ptxas keep the descriptors in uniform registers;
//  * the bias K-step: query bits become -+64, and one more K = 32 step per A tile multiplies {8, 1, 64, 64, 64, 64} by
//    {column >> 3, column & 7, 64, 64, 64, 64}, so the accumulator IS the 16-bit key (256 - a'.b) * 64 + column and one IMAD
//    packs two of them (IMAD issues at half rate on B200, like the integer min / max: tools/probes/epi_pipe_probe.cu).  Its
//    operands take 2.3 KB: the query side is one 8-row group read by all 16 groups (stride byte offset 0), the train side one
//    K-chunk read twice (leading byte offset 0) against a zero second chunk on the query side;
//  * min / second-min in two passes (mm_min2) instead of the tournament.
// Shared memory: 2 x 64 KB of A, 3 x 32 KB of B, 2.3 KB of bias operands = 226.4 KB.
// ------------------------------------------------------------------------------------------
#define M3_EW 8
#define M3_MMA_WARPS 2
#define M3_THREADS ((M3_EW + M2_PROD_WARPS + M3_MMA_WARPS) * 32)
#define M3_STAGES 3
#define M3_OFF_B (4 * MM_TILE_BYTES)
#define M3_OFF_AX (M3_OFF_B + M3_STAGES * MM_TILE_BYTES)
#define M3_AX_BYTES 256   // bias K-step, query side: 8 rows x 2 K-chunks, every row group reads the same rows
#define M3_BX_BYTES 2048  // bias K-step, train side: 128 rows x 1 K-chunk, both K-chunks read the same bytes
#define M3_OFF_BX (M3_OFF_AX + M3_AX_BYTES)
#define M3_OFF_BAR (M3_OFF_BX + M3_BX_BYTES)  // full[3], empty[3], tfull[2], tempty[2], afull[2], aempty[2]
#define M3_OFF_TMEM (M3_OFF_BAR + 14 * 8)
#define M3_OFF_FLAGS (M3_OFF_TMEM + 8)
#define M3_SMEM (M3_OFF_FLAGS + 32)
static_assert(M3_OFF_AX % 128 == 0 && M3_SMEM <= 232448, "k_match_mma3 shared memory");

// DBG (timing experiments, -DORB_B200_MMA_KNOCKOUT builds only): 1 = no tcgen05.mma issued, 2 = producers skip the expansion,
// 4 = epilogue skips the key arithmetic, 8 = and the TMEM loads; the barrier traffic stays, the results are garbage.
template <int KIND, int DBG, int BK>
__global__ void __launch_bounds__(M3_THREADS, 1) k_match_mma3(const uint8_t* __restrict__ q, const int* __restrict__ nq, size_t q_stride,
                                                              const uint8_t* __restrict__ t, const int* __restrict__ nt, size_t t_stride,
                                                              int* __restrict__ best_idx, int* __restrict__ best_dist,
                                                              int* __restrict__ second_dist, size_t out_stride, int npairs, int qblocks,
                                                              MmParams prm) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned sA = mm_smem_u32(smem);
    const unsigned bFull = sA + M3_OFF_BAR, bEmpty = bFull + 8 * M3_STAGES, bTfull = bEmpty + 8 * M3_STAGES, bTempty = bTfull + 16,
                   bAfull = bTempty + 16, bAempty = bAfull + 16;
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(smem + M3_OFF_TMEM);
    volatile int* flags = reinterpret_cast<volatile int*>(smem + M3_OFF_FLAGS);
    const int nitems = npairs * qblocks;
    // DBG & 16: CTA 0 writes clock() of its hand-over events into second_dist (role * 4 + event, 256 tiles each)
    auto trace = [&](int slot, unsigned n) {
        if ((DBG & 16) && blockIdx.x == 0 && n < 256u) second_dist[slot * 256 + n] = (int)clock();
    };
    auto wait_ahead = [&](unsigned bar, unsigned parity) {  // producers and the MMA thread: ahead of the epilogue most of the time
        if (prm.park_ns)
            mm_mbar_wait_parked(bar, parity, prm.park_ns);
        else
            mm_mbar_wait(bar, parity);
    };

    if (warp == M3_EW + M2_PROD_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(mm_smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < M3_STAGES; ++s) {
            mm_mbar_init(bFull + 8 * s, M2_PROD_WARPS * 32);
            mm_mbar_init(bEmpty + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mm_mbar_init(bTfull + 8 * a, 1);
            mm_mbar_init(bTempty + 8 * a, M3_EW);
            mm_mbar_init(bAfull + 8 * a, M2_PROD_WARPS * 32);
            mm_mbar_init(bAempty + 8 * a, M3_MMA_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (BK && tid < 16 + 128) {
        // the bias K-step: 32 more K positions whose products add 64 * 256 + (train row & 63) to every accumulator, so that with
        // the query bits as -+64 the accumulator IS the 16-bit key (256 - a'.b) * 64 + column.  Query side: bytes {8, 1, 64, 64,
        // 64, 64, 0 ...} in K-chunk 0 and zeros in K-chunk 1, the same for every row (stride byte offset 0).  Train side: bytes
        // {column >> 3, column & 7, 64, 64, 64, 64, 0 ...}, read for both K-chunks (leading byte offset 0; chunk 1 meets zeros).
        uint4 v = make_uint4(0, 0, 0, 0);
        if (tid < 8) v = make_uint4(0x40400108u, 0x00004040u, 0, 0);
        if (tid >= 16) {
            const unsigned col = (unsigned)(tid - 16) & 63u;
            v = make_uint4(0x40400000u | (col >> 3) | ((col & 7u) << 8), 0x00004040u, 0, 0);
        }
        *reinterpret_cast<uint4*>(smem + M3_OFF_AX + tid * 16) = v;  // rows 8g + r of the train side sit at g * 128 + r * 16
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    mm_fence_before();
    __syncthreads();
    mm_fence_after();
    const unsigned tmem = *tmem_slot;

    // every role enumerates the same work items and skips the same ones
#define M3_FOR_ITEMS                                                            \
    for (int it = blockIdx.x; it < nitems; it += gridDim.x) {                   \
        const int p = it / qblocks, q0 = (it - p * qblocks) * 2 * MM_M;          \
        const int nQ = nq[p], nT = nt[p];                                       \
        if (q0 >= nQ || nT >= (1 << 22)) continue;                              \
        const int ntiles = (nT + MM_N - 1) / MM_N;                              \
        const uint4* Q = reinterpret_cast<const uint4*>(q + p * q_stride);      \
        const uint4* T = reinterpret_cast<const uint4*>(t + p * t_stride);      \
        (void)Q;                                                                \
        (void)T;

    if (warp < M3_EW) {
        // ---------------- epilogue: one query row per thread ----------------
        const int tileA = warp >> 2;
        const unsigned trow = tmem + ((unsigned)(warp & 3) << 21) + (unsigned)(tileA * MM_N);
        unsigned acount = 0;  // accumulator tiles consumed so far (stage = acount & 1, phase = (acount >> 1) & 1)
        M3_FOR_ITEMS
            unsigned bestk = 0xffffffffu, seck = 0xffffffffu;
            auto merge = [&](unsigned b2, unsigned s2, int slab) {
                const unsigned kb = (b2 & 0xffc0003fu) | ((unsigned)slab << 6);
                const unsigned mx = max(kb, bestk);
                seck = min(seck, min(mx, s2));
                bestk = min(bestk, kb);
            };
            unsigned acc0[32], acc1[32];
            if (ntiles > 0) {
                mm_mbar_wait(bTfull + 8 * (acount & 1u), (acount >> 1) & 1u);
                mm_fence_after();
                const unsigned ta = trow + (acount & 1u) * 2 * MM_N;
                mm_tmem_ld32(ta, acc0);
                mm_tmem_ld32(ta + 32, acc1);
            }
            for (int tt = 0; tt < ntiles; ++tt) {
                const unsigned a = acount & 1u;
                const unsigned ta = trow + a * 2 * MM_N;
                unsigned P[32], b2, s2;
                mm_tmem_ld_wait();
                if (!(DBG & 4)) mm_slab_keys_any<KIND, BK>(acc0, acc1, prm, nT - tt * MM_N, P);
                if (!(DBG & 8)) {
                    mm_tmem_ld32(ta + 64, acc0);
                    mm_tmem_ld32(ta + 96, acc1);
                }
                if (!(DBG & 4)) {
                    mm_min2(P, b2, s2);
                    merge(b2, s2, tt * 2);
                }
                mm_tmem_ld_wait();
                mm_fence_before();
                __syncwarp();
                if (lane == 0) mm_mbar_arrive(bTempty + 8 * a);
                if (tid == 0) trace(9, acount);
                if (!(DBG & 4)) mm_slab_keys_any<KIND, BK>(acc0, acc1, prm, nT - tt * MM_N - 64, P);
                ++acount;
                if (tt + 1 < ntiles) {
                    mm_mbar_wait(bTfull + 8 * (acount & 1u), (acount >> 1) & 1u);
                    if (tid == 0) trace(8, acount);
                    mm_fence_after();
                    const unsigned tn = trow + (acount & 1u) * 2 * MM_N;
                    if (!(DBG & 8)) {
                        mm_tmem_ld32(tn, acc0);
                        mm_tmem_ld32(tn + 32, acc1);
                    }
                }
                if (!(DBG & 4)) {
                    mm_min2(P, b2, s2);
                    merge(b2, s2, tt * 2 + 1);
                } else {
                    bestk ^= acc0[0] ^ acc1[31];
                }
            }
            const int qi = q0 + tileA * MM_M + (warp & 3) * 32 + lane;
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
        mm_fence_before();
    } else if (warp < M3_EW + M2_PROD_WARPS) {
        // ---------------- producers ----------------
        const int r = tid - M3_EW * 32;
        unsigned tcount = 0, icount = 0;  // B tiles / A buffers produced so far
        M3_FOR_ITEMS
            if (ntiles == 0) continue;
            uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0;
            if (r < nT) {
                n0 = __ldg(T + 2 * (size_t)r);
                n1 = __ldg(T + 2 * (size_t)r + 1);
            }
            // the item's two A tiles (256 query rows x 2 halves, +-1 bytes) into the A buffer the MMAs of item - 2 have left
            const unsigned ab = icount & 1u;
            wait_ahead(bAempty + 8 * ab, ((icount >> 1) & 1u) ^ 1u);
            uint8_t* bufA = smem + ab * 2 * MM_TILE_BYTES;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int j = r + 128 * k;
                const int er = (j & 7) | ((j >> 4) << 3), ehf = (j >> 3) & 1;
                const int row = q0 + er;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (row < nQ) v = __ldg(Q + 2 * (size_t)row + ehf);
                mm_expand_half_row(bufA + (er >> 7) * MM_TILE_BYTES, er & 127, ehf, v, prm.lut_a, 4);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mm_mbar_arrive(bAfull + 8 * ab);
            ++icount;
            for (int tt = 0; tt < ntiles; ++tt) {
                const unsigned s = tcount % M3_STAGES;
                const uint4 v0 = n0, v1 = n1;
                {
                    const int row = (tt + 1) * MM_N + r;
                    n0 = n1 = make_uint4(0, 0, 0, 0);
                    if (row < nT) {
                        n0 = __ldg(T + 2 * (size_t)row);
                        n1 = __ldg(T + 2 * (size_t)row + 1);
                    }
                }
                unsigned upper;
                asm volatile(
                    "{\n.reg .pred p, q;\nsetp.ne.u32 p, %1, 0;\nbar.red.or.pred q, 1, %2, p;\nselp.u32 %0, 1, 0, q;\n}\n"
                    : "=r"(upper)
                    : "r"(v1.z | v1.w), "n"(M2_PROD_WARPS * 32)
                    : "memory");
                if (r == 0) trace(0, tcount);
                wait_ahead(bEmpty + 8 * s, ((tcount / M3_STAGES) & 1u) ^ 1u);
                if (r == 0) trace(1, tcount);
                uint8_t* tileB = smem + M3_OFF_B + s * MM_TILE_BYTES;
                if (!(DBG & 2)) {
                    mm_expand_half_row(tileB, r, 0, v0, prm.lut_b, 4);
                    mm_expand_half_row(tileB, r, 1, v1, prm.lut_b, upper ? 4 : 2);
                }
                if (r == 0) flags[s] = (int)upper;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mm_mbar_arrive(bFull + 8 * s);
                if (r == 0) trace(2, tcount);
                ++tcount;
            }
        }
    } else {
        // ---------------- MMA issuers: two warps, one per accumulator stage (tile parity) ----------------
        // tcgen05.mma blocks its thread while the tensor core's queue is full, and a barrier poll that succeeds at once still
        // takes ~150 cycles; with a single issuing warp those polls sat between the last MMA of one tile and the first of the
        // next and the tensor pipe idled ~500 of every 1300 cycles (tools/probes/mma_timeline.py).  With two warps one polls
        // while the other is blocked in its issue.  Each warp walks every tile and acts on its own parity; the whole warp
        // stays converged and one elected lane issues.
        const unsigned mw = (unsigned)(warp - (M3_EW + M2_PROD_WARPS));
        unsigned tcount = 0, icount = 0;  // tiles (= accumulator tiles) / A buffers so far
        const unsigned dhi = (unsigned)(prm.desc_base >> 32), dlo = (unsigned)prm.desc_base;
        constexpr unsigned KSTEP = (2 * MM_LBO) >> 4;  // one K-step further on, in descriptor units of 16 bytes
        // bias K-step descriptors: query side LBO 128 / SBO 0, train side LBO 0 / SBO 128; version bits as in desc_base
        const unsigned ver = dhi & 0xffffc000u;
        const unsigned ax_lo = (((sA + M3_OFF_AX) & 0x3ffffu) >> 4) | ((128u >> 4) << 16), x_hi = ver;
        const unsigned bx_lo = ((sA + M3_OFF_BX) & 0x3ffffu) >> 4, bx_hi = ver | (128u >> 4);
        M3_FOR_ITEMS
            if (ntiles == 0) continue;
            const unsigned ab = icount & 1u;
            wait_ahead(bAfull + 8 * ab, (icount >> 1) & 1u);
            const unsigned a0 = dlo | (((sA + ab * 2 * MM_TILE_BYTES) & 0x3ffffu) >> 4), a1 = a0 + (MM_TILE_BYTES >> 4);
            for (int tt = 0; tt < ntiles; ++tt, ++tcount) {
                if ((tcount & 1u) != mw) continue;
                const unsigned s = tcount % M3_STAGES, a = mw;
                wait_ahead(bFull + 8 * s, (tcount / M3_STAGES) & 1u);
                if (lane == 0) trace(4, tcount);
                wait_ahead(bTempty + 8 * a, ((tcount >> 1) & 1u) ^ 1u);
                if (lane == 0) trace(5, tcount);
                mm_fence_after();
                const unsigned b0 = dlo | (((sA + M3_OFF_B + s * MM_TILE_BYTES) & 0x3ffffu) >> 4);
                const bool upper = flags[s] != 0;
                const unsigned d0 = tmem + a * 2 * MM_N, d1 = d0 + MM_N;
                if (mm_elect_one()) {
                    if (!(DBG & 1)) {
                        mm_mma_w<KIND, 0>(d0, a0, b0, dhi, prm.idesc);
#pragma unroll
                        for (int k = 1; k < 6; ++k) mm_mma_w<KIND, 1>(d0, a0 + k * KSTEP, b0 + k * KSTEP, dhi, prm.idesc);
                        mm_mma_w<KIND, 0>(d1, a1, b0, dhi, prm.idesc);
#pragma unroll
                        for (int k = 1; k < 6; ++k) mm_mma_w<KIND, 1>(d1, a1 + k * KSTEP, b0 + k * KSTEP, dhi, prm.idesc);
                        if (BK) {
                            mm_mma_x<KIND>(d0, ax_lo, bx_lo, x_hi, bx_hi, prm.idesc);
                            mm_mma_x<KIND>(d1, ax_lo, bx_lo, x_hi, bx_hi, prm.idesc);
                        }
                        if (upper) {
#pragma unroll
                            for (int k = 6; k < 8; ++k) mm_mma_w<KIND, 1>(d0, a0 + k * KSTEP, b0 + k * KSTEP, dhi, prm.idesc);
#pragma unroll
                            for (int k = 6; k < 8; ++k) mm_mma_w<KIND, 1>(d1, a1 + k * KSTEP, b0 + k * KSTEP, dhi, prm.idesc);
                        }
                    }
                    mm_commit(bEmpty + 8 * s);
                    mm_commit(bTfull + 8 * a);
                }
                __syncwarp();
                if (lane == 0) trace(6, tcount);
            }
            // this A buffer is free once the MMAs both warps have issued so far are complete (an arrival from each)
            if (mm_elect_one()) mm_commit(bAempty + 8 * ab);
            __syncwarp();
            ++icount;
        }
    }
#undef M3_FOR_ITEMS
    __syncthreads();
    if (warp == M3_EW + M2_PROD_WARPS) {
        mm_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

// Train sets the tensor-core kernel leaves alone (2^22 rows and more): k_match_all's plain path, see match_kernels.cu.
cudaError_t orbk_match_all_popc(const uint8_t* q, const int* nq, size_t q_stride, const uint8_t* t, const int* nt, size_t t_stride,
                                int npairs, int max_nq, int* best_idx, int* best_dist, int* second_dist, size_t out_stride,
                                int only_huge, cudaStream_t st);

// Function attributes are per device: called by orb_matcher_create for the handle's device.
cudaError_t orbk_match_mma_init() {
    cudaError_t e = cudaFuncSetAttribute(k_match_mma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, MM_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_match_mma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, MM_SMEM);
    if (e != cudaSuccess) return e;
#define M3_ATTR(K, F)                                                                                             \
    e = cudaFuncSetAttribute(k_match_mma3<K, F, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, M3_SMEM); \
    if (e != cudaSuccess) return e;                                                                      \
    if (K == 0) e = cudaFuncSetAttribute(k_match_mma3<0, F, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, M3_SMEM); \
    if (e != cudaSuccess) return e;
    M3_ATTR(0, 0) M3_ATTR(1, 0)
#ifdef ORB_B200_MMA_KNOCKOUT
    M3_ATTR(0, 1) M3_ATTR(0, 2) M3_ATTR(0, 3) M3_ATTR(0, 4) M3_ATTR(0, 12) M3_ATTR(0, 13) M3_ATTR(0, 14) M3_ATTR(0, 15) M3_ATTR(0, 16) M3_ATTR(0, 30) M3_ATTR(0, 31)
#endif
#undef M3_ATTR
    e = cudaFuncSetAttribute(k_match_mma2<0, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, M2_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_match_mma2<0, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, M2_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_match_mma2<1, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, M2_SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_match_mma2<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, M2_SMEM);
}

// kind: 0 = kind::i8, 1 = kind::f8f6f4 (e4m3).  variant: 0 = the persistent warp-specialised kernel (k_match_mma3) with the
// documented encoding; comparators: 10 = the first form of the kernel (k_match_mma), 20 = warp-specialised, one CTA per 256
// queries, 16 epilogue warps, 30 = the same with 8 epilogue warps; 1 / 2 (11 / 12) exist for tools/probes/mma_probe.py only
// (leading / stride byte offsets swapped, descriptor version bits clear).
cudaError_t orbk_match_all_mma(const uint8_t* q, const int* nq, size_t q_stride, const uint8_t* t, const int* nt, size_t t_stride,
                               int npairs, int max_nq, int* best_idx, int* best_dist, int* second_dist, size_t out_stride, int kind,
                               int variant, cudaStream_t st) {
    if (npairs <= 0 || max_nq <= 0) return cudaSuccess;
    MmParams prm;
    prm.neg_lo = (unsigned)-64;
    prm.neg_hi = (unsigned)(-64 * 65536);
    {
        const char* e = getenv("ORB_B200_MMA_PARK");
        prm.park_ns = e ? (unsigned)atoi(e) : 0u;
    }
    prm.shl16 = 65536u;
    unsigned lbo = MM_LBO, sbo = MM_SBO;
    if (variant % 10 == 1) std::swap(lbo, sbo);
    // matrix descriptor: start address >> 4 [0,14), leading byte offset >> 4 [16,30), stride byte offset >> 4 [32,46),
    // descriptor version 1 [46,48), base offset 0, layout type 0 = no swizzle [61,64)
    prm.desc_base = ((unsigned long long)(lbo >> 4) << 16) | ((unsigned long long)(sbo >> 4) << 32) | (variant % 10 == 2 ? 0ull : (1ull << 46));
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
    if (variant < 10) {  // one CTA per SM walks the (pair, 256 queries) items
        static int smsOf[64];
        int dev = 0;
        cudaGetDevice(&dev);
        int sms = dev < 64 ? smsOf[dev] : 0;
        if (sms == 0) {
            if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
            if (dev < 64) smsOf[dev] = sms;
        }
        const int qblocks = (max_nq + 2 * MM_M - 1) / (2 * MM_M);
        const long long nitems = (long long)npairs * qblocks;
        const int grid = (int)(nitems < sms ? nitems : sms);
        const char* ebk = getenv("ORB_B200_MMA_BK");
        const bool bk = kind == 0 && !(ebk && atoi(ebk) == 0);
        if (bk) prm.lut_a = 0x0000c040u;  // query bit 0 -> +64, 1 -> -64: the accumulator is -64 a'.b (+ the bias K-step)
#define M3_GO(K, F)                                                                                                                        \
    do {                                                                                                                                   \
        if (bk)                                                                                                                            \
            k_match_mma3<0, F, 1><<<grid, M3_THREADS, M3_SMEM, st>>>(q, nq, q_stride, t, nt, t_stride, best_idx, best_dist, second_dist, \
                                                                       out_stride, npairs, qblocks, prm);                                  \
        else                                                                                                                               \
            k_match_mma3<K, F, 0><<<grid, M3_THREADS, M3_SMEM, st>>>(q, nq, q_stride, t, nt, t_stride, best_idx, best_dist, second_dist, \
                                                                       out_stride, npairs, qblocks, prm);                                  \
    } while (0)
#ifdef ORB_B200_MMA_KNOCKOUT  // timing experiments only (results are garbage): see the DBG bits in k_match_mma3
        const char* ed = getenv("ORB_B200_MMA_DEBUG");
        switch (kind == 0 && ed ? atoi(ed) : 0) {
#define M3_CASE(F)   \
    case F:          \
        M3_GO(0, F); \
        break;
            M3_CASE(1) M3_CASE(2) M3_CASE(3) M3_CASE(4) M3_CASE(12) M3_CASE(13) M3_CASE(14) M3_CASE(15) M3_CASE(16) M3_CASE(30) M3_CASE(31)
#undef M3_CASE
            default:
                if (kind != 0)
                    M3_GO(1, 0);
                else
                    M3_GO(0, 0);
        }
#else
        if (kind != 0)
            M3_GO(1, 0);
        else
            M3_GO(0, 0);
#endif
#undef M3_GO
    } else if (variant >= 10 && variant < 20) {  // the first form of the kernel (2 CTAs per SM, block barriers): kept as a comparator
        dim3 grid((max_nq + MM_M - 1) / MM_M, npairs);
        if (kind == 0)
            k_match_mma<0><<<grid, MM_THREADS, MM_SMEM, st>>>(q, nq, q_stride, t, nt, t_stride, best_idx, best_dist, second_dist, out_stride, prm);
        else
            k_match_mma<1><<<grid, MM_THREADS, MM_SMEM, st>>>(q, nq, q_stride, t, nt, t_stride, best_idx, best_dist, second_dist, out_stride, prm);
    } else {
        dim3 grid((max_nq + 2 * MM_M - 1) / (2 * MM_M), npairs);
        const bool e8 = variant >= 30;  // 20: sixteen epilogue warps (two threads per query row): measured slower still
        if (kind == 0 && !e8)
            k_match_mma2<0, 16><<<grid, M2_THREADS(16), M2_SMEM, st>>>(q, nq, q_stride, t, nt, t_stride, best_idx, best_dist, second_dist, out_stride, prm);
        else if (kind == 0)
            k_match_mma2<0, 8><<<grid, M2_THREADS(8), M2_SMEM, st>>>(q, nq, q_stride, t, nt, t_stride, best_idx, best_dist, second_dist, out_stride, prm);
        else if (!e8)
            k_match_mma2<1, 16><<<grid, M2_THREADS(16), M2_SMEM, st>>>(q, nq, q_stride, t, nt, t_stride, best_idx, best_dist, second_dist, out_stride, prm);
        else
            k_match_mma2<1, 8><<<grid, M2_THREADS(8), M2_SMEM, st>>>(q, nq, q_stride, t, nt, t_stride, best_idx, best_dist, second_dist, out_stride, prm);
    }
    orbk_count_launch(1);
    return cudaGetLastError();
}
