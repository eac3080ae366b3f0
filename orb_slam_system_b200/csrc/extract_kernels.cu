// Hand-written sm_100a kernels of the ORB extract + describe path.
//
// Path (reference src/ORBextractor.cc): ComputePyramid :497-515 -> k_resize;
// ComputeKeyPointsOctTree :288-357 (per-cell cv::FAST + retry) -> k_detect;
// DistributeOctTree :228-286 -> k_octree (closed form of the list surgery);
// IC_Angle :21-48, GaussianBlur :479, computeOrbDescriptor :57-73 -> k_blur, k_describe.
// All integer results are bit-exact with the reference semantics; float math is
// compiled with -fmad=false and IEEE div so it follows the oracle step by step.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>

#include <atomic>
#include <type_traits>

#include "orb_plan.h"
#include "extract_kernels.h"

static std::atomic<unsigned long long> g_launches{0};  // handles are used from several threads (src/Frame.cc:58-61)

// Programmatic dependent launch: a kernel launched with launch_pdl may be scheduled while its predecessor in the stream is
// still draining; it calls pdl_enter() before touching anything the predecessor wrote (griddepcontrol.wait returns once
// the predecessor grid has completed and its writes are visible) and, by issuing launch_dependents right away, lets its
// own successor do the same.  The launch latency of the dependent chains (six pyramid levels, detect -> octree -> describe)
// then overlaps the tail of the previous kernel.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
    pdl_wait();
    pdl_release();
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned vmax3(unsigned a, unsigned b, unsigned c) { return __vimax3_u16x2(a, b, c); }
__device__ __forceinline__ unsigned vmin3(unsigned a, unsigned b, unsigned c) { return __vimin3_u16x2(a, b, c); }
// four unsigned bytes (pixels) times four signed bytes (weights), accumulated into a signed int
// n / d for n < 2^16 as a multiply-high by magic = div_magic(d); d == 1 has no 32-bit magic and is passed through.
__device__ __forceinline__ unsigned div_magic(unsigned d) { return d <= 1u ? 0u : 0xffffffffu / d + 1u; }
__device__ __forceinline__ unsigned div_by(unsigned n, unsigned magic) { return magic ? __umulhi(n, magic) : n; }

__device__ __forceinline__ int dp4a_us(unsigned a, int b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// ------------------------------------------------------------------------------------------
// k_resize: cv::resize INTER_LINEAR 8UC1 (fixed point), one level from the
// previous one.  4 output pixels per thread, one 32-bit store.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_resize(const uint8_t* __restrict__ src, int spitch,
                                                unsigned long long splane, uint8_t* __restrict__ dst,
                                                int dpitch, unsigned long long dplane, int drows, int dcols,
                                                const int* __restrict__ xtab, const int* __restrict__ xcoef,
                                                const int* __restrict__ ytab, const int* __restrict__ ycoef) {
    pdl_enter();
    const int x4 = (blockIdx.x * 64 + threadIdx.x) * 4;
    const int y = blockIdx.y * 4 + threadIdx.y;
    if (x4 >= dcols || y >= drows) return;
    const int f = blockIdx.z;
    const int yt = __ldg(ytab + y), yc = __ldg(ycoef + y);
    const int b0 = yc & 0xffff, b1 = yc >> 16;
    const uint8_t* S0 = src + f * splane + (size_t)(yt & 0xffff) * spitch;
    const uint8_t* S1 = src + f * splane + (size_t)(yt >> 16) * spitch;
    unsigned out = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int x = x4 + k;
        if (x < dcols) {
            const int xt = __ldg(xtab + x), xc = __ldg(xcoef + x);
            const int s = xt & 0xffff, s1 = xt >> 16;
            const int a0 = xc & 0xffff, a1 = xc >> 16;
            const int r0 = __ldg(S0 + s) * a0 + __ldg(S0 + s1) * a1;
            const int r1 = __ldg(S1 + s) * a0 + __ldg(S1 + s1) * a1;
            const int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
            out |= (unsigned)(v & 0xff) << (8 * k);
        }
    }
    *reinterpret_cast<unsigned*>(dst + f * dplane + (size_t)y * dpitch + x4) = out;
}

// ------------------------------------------------------------------------------------------
// k_resize4: the same fixed-point resize, 4 output pixels x 4 output rows per thread.
// Per group of 4 output columns the host precomputes (xgrp): the first aligned source word wb
// and PRMT selectors that gather the 8 source bytes (S[s_k], S[s1_k], k = 0..3) out of three
// aligned words into two words laid out [A0 B0 A1 B1] [A2 B2 A3 B3]; the horizontal pass is then
// one IDP.2A per pixel and row (16-bit coefficient pair x 8-bit pixel pair).  Used when every
// group's bytes fit the 12-byte window (scale factors up to ~2.3); k_resize is the general form.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned gather8(unsigned w0, unsigned w1, unsigned w2, unsigned s01, unsigned s2) {
    const unsigned t0 = __byte_perm(w0, w1, s01 & 0xffffu), t1 = __byte_perm(w1, w2, s01 >> 16);
    return __byte_perm(t0, t1, s2);
}

__global__ void __launch_bounds__(256) k_resize4(const uint8_t* __restrict__ src, int spitch, unsigned long long splane,
                                                 uint8_t* __restrict__ dst, int dpitch, unsigned long long dplane, int drows,
                                                 int dcols, const int4* __restrict__ xgrp, const int4* __restrict__ xcoef4,
                                                 const int* __restrict__ ytab, const int* __restrict__ ycoef) {
    pdl_enter();
    const int g = blockIdx.x * 64 + threadIdx.x;
    const int ngroups = (dcols + 3) >> 2;
    const int y0 = (blockIdx.y * 4 + threadIdx.y) * 4;
    if (g >= ngroups || y0 >= drows) return;
    const int f = blockIdx.z;
    const int4 xg = __ldg(xgrp + g);   // wb, sel01 of word 0, sel01 of word 1, sel2 word0 | sel2 word1 << 16
    const int4 xc = __ldg(xcoef4 + g); // a0 | a1 << 16 for the 4 columns
    const unsigned pw = (unsigned)spitch >> 2;
    const unsigned i0 = (unsigned)xg.x, i1 = min(i0 + 1u, pw - 1u), i2 = min(i0 + 2u, pw - 1u);
    const unsigned* S = reinterpret_cast<const unsigned*>(src + f * splane);
    uint8_t* D = dst + f * dplane + 4 * g + (size_t)y0 * dpitch;
    // Horizontal pass of one source row for the 4 columns: R = S[s] * a0 + S[s1] * a1 (A.2), one IDP.2A each.
    auto hpass = [&](unsigned row, int (&r)[4]) {
        const unsigned* R = S + row * pw;
        const unsigned u0 = __ldg(R + i0), u1 = __ldg(R + i1), u2 = __ldg(R + i2);
        const unsigned p0 = gather8(u0, u1, u2, (unsigned)xg.y, (unsigned)xg.w & 0xffffu);
        const unsigned p1 = gather8(u0, u1, u2, (unsigned)xg.z, (unsigned)xg.w >> 16);
        r[0] = (int)__dp2a_lo((unsigned)xc.x, p0, 0u);
        r[1] = (int)__dp2a_hi((unsigned)xc.y, p0, 0u);
        r[2] = (int)__dp2a_lo((unsigned)xc.z, p1, 0u);
        r[3] = (int)__dp2a_hi((unsigned)xc.w, p1, 0u);
    };
    // (Reusing the horizontal sums of a source row shared by consecutive output rows was tried: the row-dependent branches
    // keep the compiler from issuing all 24 loads of the thread up front, and the kernel, which waits on loads, got slower.)
#pragma unroll
    for (int dy = 0; dy < 4; ++dy) {
        const int y = y0 + dy;
        if (y < drows) {
            const int yt = __ldg(ytab + y), yc = __ldg(ycoef + y);
            // ((b * (R >> 4)) >> 16) as one multiply-high with b << 16
            const unsigned b0 = (unsigned)(yc & 0xffff) << 16, b1 = (unsigned)(yc >> 16) << 16;
            int ra[4], rb[4];
            hpass((unsigned)(yt & 0xffff), ra);
            hpass((unsigned)(yt >> 16), rb);
            unsigned o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = (__umulhi(b0, (unsigned)ra[k] >> 4) + __umulhi(b1, (unsigned)rb[k] >> 4) + 2u) >> 2;
            *reinterpret_cast<unsigned*>(D) = __byte_perm(__byte_perm(o[0], o[1], 0x0040), __byte_perm(o[2], o[3], 0x0040), 0x5410);
        }
        D += dpitch;
    }
}

// ------------------------------------------------------------------------------------------
// k_resize_tile: k_resize4's arithmetic on a source tile staged in shared memory by one TMA box load.
// k_resize4 waits on its global loads (each thread's 24 loads follow a table load); here one elected thread
// issues the bulk copy of the 256 x RSZ_BOX_H source box while the others fetch their table entries, and the
// gathers become shared-memory reads.  One CTA = RSZ_W x RSZ_H output pixels; thread = 4 columns x 4 rows.
// The host checks (build_plan, rszTiled) that every tile's source words lie inside its box, whose x start is the
// 16-byte aligned column of the tile's first gather word and whose y start is the tile's first source row.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count);
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes);
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity);
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* map, int x, int y, int z, unsigned bar);
__device__ __forceinline__ unsigned smem_u32(const void* p);

__global__ void __launch_bounds__(RSZ_THREADS) k_resize_tile(const CUtensorMap* __restrict__ map, int frameBase, uint8_t* __restrict__ dst,
                                                             int dpitch, unsigned long long dplane, int drows, int dcols,
                                                             const int4* __restrict__ xgrp, const int4* __restrict__ xcoef4,
                                                             const int* __restrict__ ytab, const int* __restrict__ ycoef) {
    pdl_release();  // the tables read below are static; the source level is first touched by the TMA load
    __shared__ __align__(128) unsigned box[RSZ_BOX_H][64];
    __shared__ unsigned long long barMem;
    const int tid = threadIdx.x, tg = tid % (RSZ_W / 4), tr = tid / (RSZ_W / 4);
    const int ngroups = (dcols + 3) >> 2;
    const int g0 = blockIdx.x * (RSZ_W / 4), Y0 = blockIdx.y * RSZ_H, f = blockIdx.z;
    const int bx0 = (4 * __ldg(&xgrp[g0].x)) & ~15, by0 = __ldg(ytab + Y0) & 0xffff;
    const unsigned bar = smem_u32(&barMem);
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        pdl_wait();  // the previous level is complete and visible
        mbar_expect_tx(bar, (unsigned)sizeof(box));
        tma_load_3d(smem_u32(&box[0][0]), map, bx0, by0, f + frameBase, bar);
    }
    const int g = g0 + tg, y0 = Y0 + 4 * tr;
    const bool live = g < ngroups && y0 < drows;
    int4 xg = make_int4(0, 0, 0, 0), xc = xg;
    int yt[4] = {0, 0, 0, 0}, yc[4] = {0, 0, 0, 0};
    if (live) {
        xg = __ldg(xgrp + g);
        xc = __ldg(xcoef4 + g);
#pragma unroll
        for (int dy = 0; dy < 4; ++dy)
            if (y0 + dy < drows) {
                yt[dy] = __ldg(ytab + y0 + dy);
                yc[dy] = __ldg(ycoef + y0 + dy);
            }
    }
    mbar_wait(bar, 0);
    if (!live) return;
    const unsigned* B = &box[0][0] + (xg.x - (bx0 >> 2));
    uint8_t* D = dst + f * dplane + 4 * g + (size_t)y0 * dpitch;
    auto hpass = [&](int row, int (&r)[4]) {
        const unsigned* R = B + (row - by0) * 64;
        const unsigned u0 = R[0], u1 = R[1], u2 = R[2];
        const unsigned p0 = gather8(u0, u1, u2, (unsigned)xg.y, (unsigned)xg.w & 0xffffu);
        const unsigned p1 = gather8(u0, u1, u2, (unsigned)xg.z, (unsigned)xg.w >> 16);
        r[0] = (int)__dp2a_lo((unsigned)xc.x, p0, 0u);
        r[1] = (int)__dp2a_hi((unsigned)xc.y, p0, 0u);
        r[2] = (int)__dp2a_lo((unsigned)xc.z, p1, 0u);
        r[3] = (int)__dp2a_hi((unsigned)xc.w, p1, 0u);
    };
#pragma unroll
    for (int dy = 0; dy < 4; ++dy) {
        if (y0 + dy < drows) {
            const unsigned b0 = (unsigned)(yc[dy] & 0xffff) << 16, b1 = (unsigned)(yc[dy] >> 16) << 16;
            int ra[4], rb[4];
            hpass(yt[dy] & 0xffff, ra);
            hpass(yt[dy] >> 16, rb);
            unsigned o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = (__umulhi(b0, (unsigned)ra[k] >> 4) + __umulhi(b1, (unsigned)rb[k] >> 4) + 2u) >> 2;
            *reinterpret_cast<unsigned*>(D) = __byte_perm(__byte_perm(o[0], o[1], 0x0040), __byte_perm(o[2], o[3], 0x0040), 0x5410);
        }
        D += dpitch;
    }
}

// ------------------------------------------------------------------------------------------
// k_detect: per-cell FAST-9/16 + cell-local NMS + iniTh/minTh retry.
//
// One CTA = one tile = `tileCells` FAST cells of one cell row of one level of one frame,
// staged in shared memory (image tile <= 256 x 65 px).  Scores never go to HBM.
//   pass A  exact corner score m = max(c - min_arcs max_arc r, max_arcs min_arc r - c) for
//           4 pixels per thread on packed u16x2 lanes (VIMNMX3.U16x2), stored as
//           u = max(m, lowTh) - lowTh in a byte tile (0 = not a corner at lowTh).
//   pass B  strict 8-neighbour maximum inside the pixel's own cell, 4 pixels per thread on the
//           same packed lanes (no per-pixel branches); survivors go to a shared-memory list
//           and set the per-cell flag "cv::FAST + NMS at iniTh returned non-empty".
//   pass C  emit survivors at iniTh, or at minTh where the cell flag is clear
//           (reference ORBextractor.cc:293-296,330-331).
// cv::FAST semantics: candidates exist 3 px inside the cell image, NMS
// neighbours outside the cell's candidate area count as 0, response = m - 1.
// ------------------------------------------------------------------------------------------
#define DET_MAX_SURV 4096  // >= ceil(250/2) * ceil(59/2): NMS survivors are never 8-adjacent
#define DET_WLIST_PER_WARP 512  // >= ceil(59/4) rows x 32 words handled by one warp in pass B

// Dynamic shared memory of k_detect, sized by the plan's tallest tile (plan.detRows = max boxH):
//   E, O  2 x detRows x DET_EP words: the image tile expanded to one pixel per 16-bit lane, twice: word i of a row of
//         E holds pixels (2i, 2i + 1), of O pixels (2i + 1, 2i + 2), so the pixel pair starting at ANY column is one
//         aligned word and pass A gathers its ring with plain loads (no byte permutes on the ALU pipe).  After pass A
//         the E rows hold F, the NMS survivors (4 candidate columns per word)
//   sc    first the TMA landing buffer (dense 256-byte rows, the box), then (detRows - 4) x DET_SP score tile with a
//         one-word / one-row zero border; later the survivor list (u | r<<8 | cx<<16 | cell<<24)
//   tail  DetectTail
#define DET_EP 136   // words per E / O row: 4 words of left padding (ring columns left of the tile), 128 pixel pairs, 4 right
#define DET_EPAD 4
struct DetectTail {
    unsigned short wlist[(DET_THREADS / 32) * DET_WLIST_PER_WARP];  // per warp: (row << 6 | word) of its non-zero F words
    int wcount[DET_THREADS / 32];
    unsigned qmask[64];                        // per 4-column group: byte mask of the columns that are candidates
    unsigned char cellOf[DET_TILE_W + 8];      // candidate column -> cell index inside the tile
    int cellHasIni[16];
    int nWords;
    int nSurv;
    int nEmit;
    int emitBase;
    int emitFill;
    unsigned long long bar;                    // mbarrier the TMA load completes on
};

// (rounded to 128 bytes: the TMA landing buffer follows)
__host__ __device__ inline size_t det_img_bytes(int detRows) { return ((size_t)2 * detRows * DET_EP * 4 + 127) & ~(size_t)127; }
__host__ __device__ inline size_t det_sc_bytes(int detRows) {
    const size_t a = (size_t)(detRows - 4) * DET_SP, b = (size_t)DET_MAX_SURV * 4, c = (size_t)detRows * DET_TILE_W + 16;
    const size_t m = a > b ? (a > c ? a : c) : (b > c ? b : c);
    return (m + 15) & ~(size_t)15;
}
static size_t detect_smem_bytes(int detRows) {
    return det_img_bytes(detRows) + det_sc_bytes(detRows) + sizeof(DetectTail);
}

// ---- TMA / mbarrier primitives (PTX; SASS: UTMALDG, SYNCS) ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* map, int x, int y, int z, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"((unsigned long long)map), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}

// Minimum and maximum of two packed lanes on the FMA pipe.  A lane holding a pixel value 0..255 in 16 bits is, read as
// binary16, the subnormal p * 2^-24; sums and differences of such values are exact (|result| < 2^10 ulps), the clamp of
// sub.sat is max(a - b, 0) because everything is far below 1.0, and no .ftz is applied to f16 arithmetic.  So
//   d = max(a - b, 0);  hi = b + d = max(a, b);  lo = a - d = min(a, b)
// are three HADD2 that produce exactly the bit patterns VIMNMX.U16x2 would -- on the pipe k_detect leaves idle.
__device__ __forceinline__ void minmax_fma(unsigned a, unsigned b, unsigned& lo, unsigned& hi) {
    unsigned d;
    asm("sub.sat.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    asm("add.f16x2 %0, %1, %2;" : "=r"(hi) : "r"(b), "r"(d));
    asm("sub.f16x2 %0, %1, %2;" : "=r"(lo) : "r"(a), "r"(d));
}

// Per 16-bit lane: 0xffff where a > b, else 0 (lanes hold small non-negative integers = fp16 subnormals; HSET2.BM).
__device__ __forceinline__ unsigned gt_mask2(unsigned a, unsigned b) {
    unsigned r;
    asm("set.gt.u32.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}

__device__ __forceinline__ void fast_score_pairs(const unsigned (&r)[16], unsigned c2, unsigned low2,
                                                 unsigned neglow2, unsigned& u) {
    // A = min over the 16 arcs of 9 contiguous ring pixels of the arc's maximum, B = max over arcs of the arc's
    // minimum.  The arcs starting at k and k + 1 (k even) share the 8 pixels W = r[k+1 .. k+8]:
    //   min(max(r[k], W), max(W, r[k+9])) = max(W, min(r[k], r[k+9])),   W = max(M4[k+1], M4[k+5]),
    // with M4[j] = max(r[j .. j+3]) built from pair maxima at the odd positions.  36 min/max per polarity; the 16
    // (min, max) pairs of two ring pixels run on the FMA pipe (minmax_fma), the rest on the ALU pipe.
    unsigned M2[8], m2[8], M4[8], m4[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)  // odd position j = 2i + 1
        minmax_fma(r[2 * i + 1], r[(2 * i + 2) & 15], m2[i], M2[i]);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        M4[i] = __vmaxu2(M2[i], M2[(i + 1) & 7]);
        m4[i] = __vminu2(m2[i], m2[(i + 1) & 7]);
    }
    unsigned tA[8], tB[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {  // even position k = 2i: W = M4[k+1] u M4[k+5] = M4 index i and i + 2
        unsigned mn, mx;
        minmax_fma(r[2 * i], r[(2 * i + 9) & 15], mn, mx);
        tA[i] = vmax3(M4[i], M4[(i + 2) & 7], mn);
        tB[i] = vmin3(m4[i], m4[(i + 2) & 7], mx);
    }
    const unsigned A = __vminu2(vmin3(tA[0], tA[1], tA[2]), vmin3(vmin3(tA[3], tA[4], tA[5]), tA[6], tA[7]));
    const unsigned B = __vmaxu2(vmax3(tB[0], tB[1], tB[2]), vmax3(vmax3(tB[3], tB[4], tB[5]), tB[6], tB[7]));
    // m = max(c - A, B - c); u = max(m, low) - low = sat(max(sat(c - A), sat(B - c)) - low) for low >= 0: the
    // saturating differences are HADD2.SAT on the FMA pipe (fp16 subnormal lanes, exact), one VIMNMX joins them
    unsigned d1, d2;
    asm("sub.sat.f16x2 %0, %1, %2;" : "=r"(d1) : "r"(c2), "r"(A));
    asm("sub.sat.f16x2 %0, %1, %2;" : "=r"(d2) : "r"(B), "r"(c2));
    const unsigned m = __vmaxu2(d1, d2);
    asm("sub.sat.f16x2 %0, %1, %2;" : "=r"(u) : "r"(m), "r"(low2));
    (void)neglow2;
}

// bytes 0 and 2 / bytes 1 and 3 of a word as two zero-extended 16-bit lanes
__device__ __forceinline__ unsigned even_lanes(unsigned w) { return w & 0x00ff00ffu; }
__device__ __forceinline__ unsigned odd_lanes(unsigned w) { return __byte_perm(w, 0u, 0x4341); }

__global__ void __launch_bounds__(DET_THREADS) k_detect(const __grid_constant__ OrbPlan plan,
                                                        const CUtensorMap* __restrict__ maps, int tileOffset) {
    pdl_release();  // nothing before the TMA load below depends on the pyramid kernels
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int detRows = plan.detRows;
    unsigned (*E)[DET_EP] = reinterpret_cast<unsigned (*)[DET_EP]>(smem_raw);
    unsigned (*O)[DET_EP] = E + detRows;
    unsigned (*sc)[DET_SP / 4] = reinterpret_cast<unsigned (*)[DET_SP / 4]>(smem_raw + det_img_bytes(detRows));
    unsigned* surv = reinterpret_cast<unsigned*>(sc);
    DetectTail& sm = *reinterpret_cast<DetectTail*>(smem_raw + det_img_bytes(detRows) + det_sc_bytes(detRows));

    const int f = blockIdx.y;
    const int tile = blockIdx.x + tileOffset;
    if (tile >= plan.totalTiles) return;
    const unsigned te = __ldg(plan.detTileTab + tile);  // level | cell row << 4 | tile column << 18
    const int l = (int)(te & 15u), ci = (int)((te >> 4) & 0x3fffu), tx = (int)(te >> 18);
    const OrbLevel& L = plan.lv[l];
    const int j0 = tx * L.tileCells, j1 = min(j0 + L.tileCells, L.nCols);
    const int maxBX = L.cols - ORB_MINB, maxBY = L.rows - ORB_MINB;
    const int X0 = ORB_MINB + j0 * L.wCell;
    const int X1 = min(ORB_MINB + j1 * L.wCell + 6, maxBX);
    const int Y0 = ORB_MINB + ci * L.hCell;
    const int Y1 = min(Y0 + L.hCell + 6, maxBY);
    const int TW = X1 - X0, TH = Y1 - Y0;
    const int CW = TW - 6, CH = TH - 6;  // candidate area
    if (CW <= 0 || CH <= 0) return;
    const int tid = threadIdx.x;
    const int wCell = L.wCell;

    // ---- stage the image tile with one TMA box load.  The TMA unit needs a 16-byte aligned start
    // address, so the 256 x boxH box starts at X0a = X0 & ~15: box byte (r, c) = level pixel (Y0 + r, X0a + c).
    // Candidate column cx (level x = X0 + 3 + cx) is box column a0 + cx with a0 = (X0 & 15) + 3; pass A works on
    // aligned groups of four box columns 4(wo + q) .. + 3, so group q, byte k <-> cx = 4q + k - ph.
    // Out-of-image parts of the box are zero-filled by the TMA unit.
    const int a0 = (X0 & 15) + 3, wo = a0 >> 2, ph = a0 & 3;
    const int QR = (CW + ph + 3) >> 2;  // 4-pixel groups per candidate row
    const unsigned bar = smem_u32(&sm.bar);
    unsigned* stage = reinterpret_cast<unsigned*>(sc);  // the landing buffer shares the score tile's memory
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        pdl_wait();  // the level is complete and visible
        mbar_expect_tx(bar, (unsigned)(DET_TILE_W * L.boxH));
        tma_load_3d(smem_u32(stage), maps + l, X0 - (X0 & 15), Y0, f + plan.frameBase, bar);
    }
    {
        // cellOf[cx + 4] for cx in [-4, CW + 4): cell index of candidate column cx (255 left of the tile)
        const unsigned wcMagic = 0xffffffffu / (unsigned)wCell + 1u;  // (i - 4) / wCell as a multiply-high
        for (int i = tid; i < CW + 8; i += DET_THREADS) sm.cellOf[i] = i < 4 ? (unsigned char)255 : (unsigned char)__umulhi((unsigned)(i - 4), wcMagic);
        if (tid < 16) sm.cellHasIni[tid] = 0;
        if (tid < 64) {  // group q, byte k <-> cx = 4q + k - ph: keep 0 <= cx < CW
            unsigned mk = 0xffffffffu;
            const int rem = CW + ph - 4 * tid;  // bytes of this group left of the candidate area's end
            if (rem < 4) mk = rem > 0 ? (1u << (8 * rem)) - 1u : 0u;
            if (tid == 0) mk &= 0xffffffffu << (8 * ph);  // bytes before candidate column 0
            sm.qmask[tid] = mk;
        }
        if (tid == 0) {
            sm.nWords = 0;
            sm.nSurv = 0;
            sm.nEmit = 0;
            sm.emitFill = 0;
        }
    }
    mbar_wait(bar, 0);  // the tile has landed (async-proxy writes are visible after the wait)
    // ---- expand: box words wo - 1 .. wo + QR of the TH tile rows -> E / O pixel-pair words
    {
        const int nw = QR + 2;  // one word of ring halo on each side
        const unsigned nwMagic = 0xffffffffu / (unsigned)nw + 1u;
        for (int item = tid; item < TH * nw; item += DET_THREADS) {
            const int r = (int)__umulhi((unsigned)item, nwMagic);
            const int wv = wo - 1 + (item - r * nw);
            const unsigned w = stage[r * (DET_TILE_W / 4) + wv], wn = stage[r * (DET_TILE_W / 4) + wv + 1];
            const unsigned v = __funnelshift_r(w, wn, 8);  // pixels 4wv + 1 .. 4wv + 4
            *reinterpret_cast<uint2*>(&E[r][DET_EPAD + 2 * wv]) = make_uint2(__byte_perm(w, 0u, 0x4140), __byte_perm(w, 0u, 0x4342));
            *reinterpret_cast<uint2*>(&O[r][DET_EPAD + 2 * wv]) = make_uint2(__byte_perm(v, 0u, 0x4140), __byte_perm(v, 0u, 0x4342));
        }
    }
    __syncthreads();  // E / O complete, the landing buffer is dead: its memory becomes the score tile
    {
        // zero border of the score tile: rows 0 and CH+1, words 0 and QR+1
        for (int i = tid; i < 2 * (QR + 2); i += DET_THREADS) {
            const int r = i < QR + 2 ? 0 : CH + 1;
            sc[r][i < QR + 2 ? i : i - (QR + 2)] = 0u;
        }
        for (int i = tid; i < 2 * (CH + 2); i += DET_THREADS) {
            const int r = i >> 1;
            sc[r][(i & 1) ? QR + 1 : 0] = 0u;
        }
    }

    // ---- pass A: scores.  item (r, q): candidate row r, box columns 4(wo + q) .. + 3 as the pixel pairs P0 = (0, 1), P1 = (2, 3)
    const unsigned low2 = (unsigned)plan.lowTh * 0x00010001u;
    const unsigned neglow2 = ((unsigned)(-plan.lowTh) & 0xffffu) * 0x00010001u;
    const int q = tid & 63, grp = tid >> 6;  // (pass B mapping)
    {
        // items (row, group) flattened over all threads so that no lane idles when QR < 64
        const unsigned qrMagic = div_magic((unsigned)QR);  // item / QR for item < 2^16
        for (int item = tid; item < CH * QR; item += DET_THREADS) {
            const int r = (int)div_by((unsigned)item, qrMagic);
            const int q = item - r * QR;
            // h = word of pixel pair P0 in E; ring offset dx: even -> E[h + dx/2], odd -> O[h + (dx-1)/2]; P1 is the next word
            const unsigned* e = &E[r][DET_EPAD + 2 * (wo + q)];
            const unsigned* o = &O[r][DET_EPAD + 2 * (wo + q)];
#define ROWE(rr, i) e[(rr) * DET_EP + (i)]
#define ROWO(rr, i) o[(rr) * DET_EP + (i)]
            unsigned r0[16], r1[16];  // ring of P0, ring of P1
            // rows 6 and 0 (dy = +3, -3): dx = 0, 1, -1
            {
                const uint2 e6 = *reinterpret_cast<const uint2*>(&ROWE(6, 0)), o6 = *reinterpret_cast<const uint2*>(&ROWO(6, 0));
                const unsigned o6m = ROWO(6, -1);
                r0[0] = e6.x; r1[0] = e6.y;      // (0, 3)
                r0[1] = o6.x; r1[1] = o6.y;      // (1, 3)
                r0[15] = o6m; r1[15] = o6.x;     // (-1, 3)
                const uint2 e0 = *reinterpret_cast<const uint2*>(&ROWE(0, 0)), o0 = *reinterpret_cast<const uint2*>(&ROWO(0, 0));
                const unsigned o0m = ROWO(0, -1);
                r0[8] = e0.x; r1[8] = e0.y;      // (0, -3)
                r0[7] = o0.x; r1[7] = o0.y;      // (1, -3)
                r0[9] = o0m; r1[9] = o0.x;       // (-1, -3)
            }
            // rows 5 and 1 (dy = +2, -2): dx = 2, -2
            {
                const uint2 e5 = *reinterpret_cast<const uint2*>(&ROWE(5, 0));
                const unsigned e5m = ROWE(5, -1), e5p = ROWE(5, 2);
                r0[2] = e5.y; r1[2] = e5p;       // (2, 2)
                r0[14] = e5m; r1[14] = e5.x;     // (-2, 2)
                const uint2 e1 = *reinterpret_cast<const uint2*>(&ROWE(1, 0));
                const unsigned e1m = ROWE(1, -1), e1p = ROWE(1, 2);
                r0[6] = e1.y; r1[6] = e1p;       // (2, -2)
                r0[10] = e1m; r1[10] = e1.x;     // (-2, -2)
            }
            // rows 4, 3, 2 (dy = +1, 0, -1): dx = 3, -3
            {
                const uint2 o4m = *reinterpret_cast<const uint2*>(&ROWO(4, -2));
                const unsigned o4a = ROWO(4, 1), o4b = ROWO(4, 2);
                r0[3] = o4a; r1[3] = o4b;        // (3, 1)
                r0[13] = o4m.x; r1[13] = o4m.y;  // (-3, 1)
                const uint2 o3m = *reinterpret_cast<const uint2*>(&ROWO(3, -2));
                const unsigned o3a = ROWO(3, 1), o3b = ROWO(3, 2);
                r0[4] = o3a; r1[4] = o3b;        // (3, 0)
                r0[12] = o3m.x; r1[12] = o3m.y;  // (-3, 0)
                const uint2 o2m = *reinterpret_cast<const uint2*>(&ROWO(2, -2));
                const unsigned o2a = ROWO(2, 1), o2b = ROWO(2, 2);
                r0[5] = o2a; r1[5] = o2b;        // (3, -1)
                r0[11] = o2m.x; r1[11] = o2m.y;  // (-3, -1)
            }
            const uint2 cen = *reinterpret_cast<const uint2*>(&ROWE(3, 0));
#undef ROWE
#undef ROWO
            unsigned u01, u23;
            fast_score_pairs(r0, cen.x, low2, neglow2, u01);
            fast_score_pairs(r1, cen.y, low2, neglow2, u23);
            const unsigned word = __byte_perm(u01, u23, 0x6420) & sm.qmask[q];  // bytes: p0, p1, p2, p3, candidates only
            sc[r + 1][q + 1] = word;
        }
    }
    __syncthreads();

    // ---- pass B: cell-local NMS on packed lanes.  Thread (q, grp) walks a strip of rows with a
    // 3-row sliding window.  For candidate col 4q+k the left / right neighbours are dropped when
    // they belong to another cell (maskL / maskR).  Non-zero survivor words are compacted with
    // one ballot per row step (no per-pixel branches).
    const int iniU = plan.iniTh - plan.lowTh + 1;  // u >= iniU  <=>  m > iniTh
    const int minU = plan.minTh - plan.lowTh + 1;
    {
        const bool active = q < QR;
        const int lane = tid & 31;
        unsigned mLe = 0, mLo = 0, mRe = 0, mRo = 0;  // 0x00ff per lane where the neighbour counts
        if (active) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int cx = 4 * q + k - ph;
                const bool first = sm.cellOf[cx + 3] != sm.cellOf[cx + 4];
                const bool last = sm.cellOf[cx + 5] != sm.cellOf[cx + 4];
                const unsigned ln = 0xffu << (16 * (k >> 1));
                if (!first) { if (k & 1) mLo |= ln; else mLe |= ln; }
                if (!last) { if (k & 1) mRo |= ln; else mRe |= ln; }
            }
        }
        const int rpg = (CH + 3) >> 2;
        const int r0 = grp * rpg, r1 = min(r0 + rpg, CH);  // identical for the whole warp
        const int wbase = (tid >> 5) * DET_WLIST_PER_WARP;
        int wcount = 0;  // warp-uniform
        if (r0 < r1) {
            // window rows in bordered coordinates: up = r, mid = r + 1, dn = r + 2
            unsigned upAo, upBe, upBo, upCe, midAo, midBe, midBo, midCe;
            {
                const unsigned a = sc[r0][q], b = sc[r0][q + 1], c = sc[r0][q + 2];
                upAo = odd_lanes(a); upBe = even_lanes(b); upBo = odd_lanes(b); upCe = even_lanes(c);
                const unsigned a2 = sc[r0 + 1][q], b2 = sc[r0 + 1][q + 1], c2 = sc[r0 + 1][q + 2];
                midAo = odd_lanes(a2); midBe = even_lanes(b2); midBo = odd_lanes(b2); midCe = even_lanes(c2);
            }
            for (int r = r0; r < r1; ++r) {
                const unsigned a = sc[r + 2][q], b = sc[r + 2][q + 1], c = sc[r + 2][q + 2];
                const unsigned dnAo = odd_lanes(a), dnBe = even_lanes(b), dnBo = odd_lanes(b), dnCe = even_lanes(c);
                // vertical maxima of the neighbour columns (centre column without the centre row)
                const unsigned vAo = vmax3(upAo, midAo, dnAo), vBe = vmax3(upBe, midBe, dnBe);
                const unsigned vBo = vmax3(upBo, midBo, dnBo), vCe = vmax3(upCe, midCe, dnCe);
                const unsigned vCenE = __vmaxu2(upBe, dnBe), vCenO = __vmaxu2(upBo, dnBo);
                // even pixels (k=0,2): left = bytes 4q-1, 4q+1 ; right = bytes 4q+1, 4q+3
                const unsigned Le = __byte_perm(vAo, vBo, 0x5432) & mLe;
                const unsigned Re = vBo & mRe;
                // odd pixels (k=1,3): left = bytes 4q, 4q+2 ; right = bytes 4q+2, 4q+4
                const unsigned Lo = vBe & mLo;
                const unsigned Ro = __byte_perm(vBe, vCe, 0x5432) & mRo;
                const unsigned nbE = vmax3(Le, Re, vCenE), nbO = vmax3(Lo, Ro, vCenO);
                // keep u where u > nb: one packed fp16 compare per lane pair (the lanes hold 0..255, exact as subnormals)
                const unsigned kE = gt_mask2(midBe, nbE), kO = gt_mask2(midBo, nbO);
                const unsigned outw = active ? ((midBe & kE) | ((midBo & kO) << 8)) : 0u;
                const unsigned nz = __ballot_sync(0xffffffffu, outw != 0u);
                if (outw) {  // this warp's private list region: no atomics
                    E[r][q] = outw;
                    sm.wlist[wbase + wcount + __popc(nz & ((1u << lane) - 1u))] = (unsigned short)((r << 6) | q);
                }
                wcount += __popc(nz);
                upAo = midAo; upBe = midBe; upBo = midBo; upCe = midCe;
                midAo = dnAo; midBe = dnBe; midBo = dnBo; midCe = dnCe;
            }
        }
        if (lane == 0) sm.wcount[tid >> 5] = wcount;
    }
    __syncthreads();  // F / wlist complete, score tile dead (its memory becomes the survivor list)
    // ---- survivors out of the compacted words; per-cell "non-empty at iniTh" flags
    {
        int nW = 0;
#pragma unroll
        for (int w = 0; w < DET_THREADS / 32; ++w) nW += sm.wcount[w];
        if (nW == 0) return;
        const int myW = sm.wcount[tid >> 5], wb = (tid >> 5) * DET_WLIST_PER_WARP;
        for (int i = tid & 31; i < myW; i += 32) {
            const int idx = sm.wlist[wb + i];
            const int r = idx >> 6, qq = idx & 63;
            unsigned w = E[r][qq];
            while (w) {  // at most two survivors per word (never 8-adjacent)
                const int k = (__ffs(w) - 1) >> 3;
                const unsigned u = (w >> (8 * k)) & 0xffu;
                w &= ~(0xffu << (8 * k));
                const int cx = 4 * qq + k - ph;
                const unsigned cell = sm.cellOf[cx + 4];
                const int slot = atomicAdd(&sm.nSurv, 1);
                if (slot < DET_MAX_SURV) surv[slot] = u | ((unsigned)r << 8) | ((unsigned)cx << 16) | (cell << 24);
                if ((int)u >= iniU) sm.cellHasIni[cell] = 1;
            }
        }
    }
    __syncthreads();

    // ---- pass C: per-cell threshold choice, reserve, emit
    const int nS = min(sm.nSurv, DET_MAX_SURV);
    if (nS == 0) return;
    int myCount = 0;
    for (int i = tid; i < nS; i += DET_THREADS) {
        const unsigned e = surv[i];
        const int u = e & 0xff;
        myCount += (u >= iniU || (!sm.cellHasIni[e >> 24] && u >= minU)) ? 1 : 0;
    }
    if (myCount) atomicAdd(&sm.nEmit, myCount);
    __syncthreads();
    if (sm.nEmit == 0) return;
    if (tid == 0) sm.emitBase = atomicAdd(&plan.candCount[f * ORB_MAX_LEVELS + l], sm.nEmit);
    __syncthreads();
    if (myCount) {
        int slot = atomicAdd(&sm.emitFill, myCount) + sm.emitBase;
        uint2* out = L.cand + (size_t)f * L.candCap;
        for (int i = tid; i < nS; i += DET_THREADS) {
            const unsigned e = surv[i];
            const int u = e & 0xff;
            if (u >= iniU || (!sm.cellHasIni[e >> 24] && u >= minU)) {
                // coordinates relative to (minBorderX, minBorderY) as in vToDistributeKeys
                const unsigned xrel = (unsigned)(X0 + 3 - ORB_MINB) + ((e >> 16) & 0xffu);
                const unsigned yrel = (unsigned)(Y0 + 3 - ORB_MINB) + ((e >> 8) & 0xffu);
                if ((unsigned)slot < L.candCap) out[slot] = make_uint2(xrel | (yrel << 16), (unsigned)(u + plan.lowTh - 1));
                ++slot;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_octree: ORBextractor::DistributeOctTree in closed form (DESIGN.md "octree").
//
// One CTA per (source level, frame).  Every candidate gets a 32-bit path code
// [root:6][c_1:2]...[c_13:2] by replaying DivideNode's floor-halving on its own box.
// After sorting by code, `div_j` = depth at which neighbours j-1, j part; the node count
// after pass t is 1 + #{div_j <= t}; p* = first pass with count >= N or all singletons;
// final nodes = runs between div_j <= p*; per node the first max-response key; output
// order = (birth pass desc, alternating-direction path) -- the std::list push_front order.
// ------------------------------------------------------------------------------------------
#define OCT_SMEM_A 8192
#define OCT_SMEM_B 4096
#define OCT_D ORB_OCT_DEPTH

template <typename T>
__device__ __forceinline__ void bitonic_sort(T* a, unsigned npad) {
    for (unsigned k = 2; k <= npad; k <<= 1) {
        for (unsigned j = k >> 1; j > 0; j >>= 1) {
            for (unsigned i = threadIdx.x; i < (npad >> 1); i += blockDim.x) {
                const unsigned lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                const unsigned hi = lo | j;
                const bool up = (lo & k) == 0;
                const T x = a[lo], y = a[hi];
                if ((x > y) == up) {
                    a[lo] = y;
                    a[hi] = x;
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int div_depth(unsigned ca, unsigned cb) {
    const unsigned x = ca ^ cb;
    if (x == 0) return OCT_D + 1;
    const int hb = 31 - __clz(x);
    if (hb >= 2 * OCT_D) return 0;  // different roots
    return OCT_D - (hb >> 1);
}

struct OctShared {
    int hist[OCT_D + 2];
    int pstar, K, nvalid, bad;
    int warpSums[32];  // one per warp of the 1024-thread CTA
};

// Path code of one candidate: root = int(x / hX) (ORBextractor.cc:248), then OCT_D floor-halving
// splits of the root box exactly as DivideNode does (:179-218).  Returns false when the root
// index is out of range (the reference drops such keys, :249).
__device__ __forceinline__ bool path_code(const OrbLevel& L, int x, int y, unsigned& code) {
    const int root = (int)__fdiv_rn((float)x, L.hX);
    if (!(root >= 0 && root < L.nIni)) return false;
    int ulx = (int)__fmul_rn(L.hX, (float)root);
    int urx = (int)__fmul_rn(L.hX, (float)(root + 1));
    int uly = 0, bly = L.H;
    code = (unsigned)root;
#pragma unroll
    for (int d = 0; d < OCT_D; ++d) {
        const int midx = ulx + ((urx - ulx) >> 1);  // extents are >= 0 here
        const int midy = uly + ((bly - uly) >> 1);
        const unsigned cx = x >= midx, cy = y >= midy;
        if (cx) ulx = midx; else urx = midx;
        if (cy) uly = midy; else bly = midy;
        code = (code << 2) | (cy << 1) | cx;
    }
    return true;
}

// The same replay cut at `depth` <= OCTF_MAX_DEPTH halvings (k_octree_fast never looks deeper than its tables):
// code = [root][c_1:2]...[c_depth:2].
__device__ __forceinline__ bool path_code_to(const OrbLevel& L, int x, int y, int depth, unsigned& code) {
    const int root = (int)__fdiv_rn((float)x, L.hX);
    if (!(root >= 0 && root < L.nIni)) return false;
    int ulx = (int)__fmul_rn(L.hX, (float)root);
    int urx = (int)__fmul_rn(L.hX, (float)(root + 1));
    int uly = 0, bly = L.H;
    code = (unsigned)root;
#pragma unroll
    for (int d = 0; d < 7; ++d) {
        if (d < depth) {
            const int midx = ulx + ((urx - ulx) >> 1);
            const int midy = uly + ((bly - uly) >> 1);
            const unsigned cx = x >= midx, cy = y >= midy;
            if (cx) ulx = midx; else urx = midx;
            if (cy) uly = midy; else bly = midy;
            code = (code << 2) | (cy << 1) | cx;
        }
    }
    return true;
}

// Candidate order (cells row-major, then row-major inside the cell) as one integer that also
// carries the coordinates: cell index << 26 | y << 13 | x.
__device__ __forceinline__ unsigned long long cand_order(const OrbLevel& L, unsigned x, unsigned y) {
    const unsigned cj = (x - 3) / (unsigned)L.wCell, ci = (y - 3) / (unsigned)L.hCell;
    return ((unsigned long long)(ci * (unsigned)L.nCols + cj) << 26) | (y << 13) | x;
}
// Same with the two divisions as multiply-high by magic = 0xffffffff / d + 1 (exact for operands below 2^16).
__device__ __forceinline__ unsigned long long cand_order_magic(const OrbLevel& L, unsigned x, unsigned y, unsigned wMagic, unsigned hMagic) {
    const unsigned cj = __umulhi(x - 3, wMagic), ci = __umulhi(y - 3, hMagic);
    return ((unsigned long long)(ci * (unsigned)L.nCols + cj) << 26) | (y << 13) | x;
}

#define OCTF_THREADS 1024

// ------------------------------------------------------------------------------------------
// octree_generic: sort-based closed form, any depth up to OCT_D, for the (level, frame) problems whose stopping depth is
// beyond k_octree_fast's tables (sparse / clustered keys).  Runs inside k_octree_fast's CTA, on its shared memory
// (OCT_SMEM_A + OCT_SMEM_B keys = the 96 KB of the tables; larger problems sort in the level's global scratch), so the
// common case pays no second launch.
// ------------------------------------------------------------------------------------------
__device__ __noinline__ void octree_generic(const OrbPlan& plan, int lt, int f, unsigned char* smem_raw, OctShared& sh) {
    unsigned long long* smA = reinterpret_cast<unsigned long long*>(smem_raw);
    unsigned long long* smB = smA + OCT_SMEM_A;
    const int l = plan.lv[lt].src;
    const OrbLevel& L = plan.lv[l];
    const int tid = threadIdx.x;
    int n = plan.candCount[f * ORB_MAX_LEVELS + l];
    if (n > (int)L.candCap) n = (int)L.candCap;
    const uint2* cand = L.cand + (size_t)f * L.candCap;
    unsigned long long* scratch = plan.lv[lt].sortScratch + (size_t)f * 2 * plan.lv[lt].sortCap;

    unsigned npad = 2;
    while (npad < (unsigned)n) npad <<= 1;
    unsigned long long* A = npad <= OCT_SMEM_A ? smA : scratch;

    // ---- path codes
    if (tid < OCT_D + 2) sh.hist[tid] = 0;
    if (tid == 0) { sh.nvalid = 0; sh.bad = 0; }
    __syncthreads();
    int myValid = 0;
    for (unsigned i = tid; i < npad; i += OCTF_THREADS) {
        unsigned long long key = ~0ull;
        if (i < (unsigned)n) {
            const uint2 c = cand[i];
            unsigned code;
            if (path_code(L, (int)(c.x & 0xffff), (int)(c.x >> 16), code)) {
                key = ((unsigned long long)code << 32) | i;
                ++myValid;
            }
        }
        A[i] = key;
    }
    if (myValid) atomicAdd(&sh.nvalid, myValid);
    __syncthreads();
    n = sh.nvalid;  // keys with an out-of-range root are dropped (ORBextractor.cc:249)

    if (n > 0) bitonic_sort(A, npad);

    // ---- histogram of parting depths
    {
        int local[OCT_D + 2];
#pragma unroll
        for (int d = 0; d < OCT_D + 2; ++d) local[d] = 0;
        for (int j = tid + 1; j < n; j += OCTF_THREADS) {
            const int dd = div_depth((unsigned)(A[j - 1] >> 32), (unsigned)(A[j] >> 32));
#pragma unroll
            for (int d = 0; d < OCT_D + 2; ++d) local[d] += (dd == d);
        }
#pragma unroll
        for (int d = 0; d < OCT_D + 2; ++d) {
            int v = local[d];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((tid & 31) == 0 && v) atomicAdd(&sh.hist[d], v);
        }
    }
    __syncthreads();

    {
        const OrbLevel& T = plan.lv[lt];
        uint2* kept = T.kept + (size_t)f * T.kmax;
        if (n == 0) {
            if (tid == 0) plan.keptCount[f * ORB_MAX_LEVELS + lt] = 0;
            return;
        }
        if (tid == 0) {
            int cnt = 1, p = -1;
            cnt += sh.hist[0];
            for (int t = 1; t <= OCT_D; ++t) {
                cnt += sh.hist[t];
                if (cnt >= T.nFeat || cnt == n) { p = t; break; }
            }
            if (p < 0) {  // unseparable keys: the reference never terminates
                p = OCT_D;
                sh.bad = 1;
            }
            int K = 1;
            for (int t = 0; t <= p; ++t) K += sh.hist[t];
            sh.pstar = p;
            sh.K = K;
        }
        __syncthreads();
        const int pstar = sh.pstar, K = sh.K;
        unsigned kpad = 2;
        while (kpad < (unsigned)K) kpad <<= 1;
        unsigned long long* B = kpad <= OCT_SMEM_B ? smB : scratch + plan.lv[lt].sortCap;

        // segment ids: block-wide exclusive scan of head flags over contiguous chunks
        const int chunk = (n + OCTF_THREADS - 1) / OCTF_THREADS;
        const int beg = min(tid * chunk, n), end = min(beg + chunk, n);
        int heads = 0;
        for (int j = beg; j < end; ++j) {
            const bool head = j == 0 || div_depth((unsigned)(A[j - 1] >> 32), (unsigned)(A[j] >> 32)) <= pstar;
            heads += head;
        }
        int incl = heads;
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if ((tid & 31) >= o) incl += v;
        }
        if ((tid & 31) == 31) sh.warpSums[tid >> 5] = incl;
        __syncthreads();
        if (tid < 32) {
            int v = tid < OCTF_THREADS / 32 ? sh.warpSums[tid] : 0;
            for (int o = 1; o < 32; o <<= 1) {
                const int t2 = __shfl_up_sync(0xffffffffu, v, o);
                if (tid >= o) v += t2;
            }
            if (tid < OCTF_THREADS / 32) sh.warpSums[tid] = v;
        }
        __syncthreads();
        int seg = incl - heads + ((tid >> 5) ? sh.warpSums[(tid >> 5) - 1] : 0);
        for (unsigned i = tid + K; i < kpad; i += OCTF_THREADS) B[i] = ~0ull;

        for (int j = beg; j < end; ++j) {
            const unsigned cj = (unsigned)(A[j] >> 32);
            const int dj = j == 0 ? 0 : div_depth((unsigned)(A[j - 1] >> 32), cj);
            if (j != 0 && dj > pstar) continue;  // not a head
            // walk the run, keep the first max-response key in candidate order
            int e = j + 1, dnext = 0;
            unsigned bestIdx = (unsigned)A[j];
            uint2 bc = cand[bestIdx];
            unsigned bestScore = bc.y;
            unsigned long long bestOrd = 0;
            bool haveOrd = false;
            for (; e < n; ++e) {
                dnext = div_depth((unsigned)(A[e - 1] >> 32), (unsigned)(A[e] >> 32));
                if (dnext <= pstar) break;
                const unsigned idx = (unsigned)A[e];
                const uint2 c = cand[idx];
                if (c.y > bestScore) {
                    bestScore = c.y;
                    bestIdx = idx;
                    bc = c;
                    haveOrd = false;
                } else if (c.y == bestScore) {
                    // candidate order: cells row-major, then (y, x) inside the cell
                    auto ordOf = [&](uint2 v) -> unsigned long long { return cand_order(L, v.x & 0xffff, v.x >> 16); };
                    if (!haveOrd) { bestOrd = ordOf(bc); haveOrd = true; }
                    const unsigned long long o2 = ordOf(c);
                    if (o2 < bestOrd) { bestOrd = o2; bestIdx = idx; bc = c; }
                }
            }
            if (e >= n) dnext = 0;
            const int len = e - j;
            int birth = len == 1 ? max(dj, dnext) : pstar;
            if (birth > pstar) birth = pstar;  // (cannot happen for len==1; keeps the key well-formed)
            // order key: (D - birth) | root' | c'_1..c'_birth ; direction alternates backwards from c_birth
            const unsigned root = cj >> (2 * OCT_D);
            unsigned long long ok = (unsigned long long)(OCT_D - birth);
            bool rootDesc = birth >= 1 && (((birth - 1) & 1) == 0);
            ok = (ok << 6) | (rootDesc ? 63u - root : root);
#pragma unroll
            for (int i = 1; i <= OCT_D; ++i) {
                unsigned c = (cj >> (2 * (OCT_D - i))) & 3u;
                if (i <= birth) {
                    if (((birth - i) & 1) == 0) c = 3u - c;
                } else {
                    c = 0;
                }
                ok = (ok << 2) | c;
            }
            B[seg] = (ok << 24) | bestIdx;
            ++seg;
        }
        __syncthreads();
        bitonic_sort(B, kpad);
        for (int r = tid; r < K; r += OCTF_THREADS) {
            const unsigned idx = (unsigned)(B[r] & 0xffffffu);
            const uint2 c = cand[idx];
            const unsigned x = (c.x & 0xffff) + ORB_MINB, y = (c.x >> 16) + ORB_MINB;
            if (r < T.kmax) kept[r] = make_uint2(x | (y << 16), c.y);
        }
        if (tid == 0) {
            plan.keptCount[f * ORB_MAX_LEVELS + lt] = K;
            if (sh.bad) plan.status[f] = 1;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// k_octree_fast: sort-free path for the common case where the stopping depth p* is shallow
// (nIni * 4^p* <= OCTF_MAX_NODES).  The quadtree nodes of depth t form a table indexed by
// the t-level path prefix; one pass of shared-memory atomics fills the per-depth occupancy
// counts of every candidate's ancestors.  Then: Count_t = non-empty entries of table t; p* as
// in the reference's stop rule; per node of depth p* a 64-bit atomicMax picks the first
// max-response key; each non-empty depth-p* entry maps to one final node (itself, or the
// ancestor where it became a singleton = its birth pass); output rank = position of the final
// node in (birth desc, alternating-direction path) order, obtained by scattering into
// per-birth tables in transformed-index order and one block-wide prefix sum.
// Problems it cannot take (deep p*) are flagged for the generic sort-based kernel below.
// ------------------------------------------------------------------------------------------
#define OCTF_MAX_NODES 6144
#define OCTF_MAX_TOTAL 8192
#define OCTF_MAX_DEPTH 7

struct OctFastSmem {
    unsigned cnt[OCTF_MAX_TOTAL];              // occupancy counts, tables of depth 0..tmax back to back
    unsigned long long best[OCTF_MAX_NODES];   // per depth-p* node: score << 48 | (ORDMAX - order)
    unsigned short ordTab[OCTF_MAX_TOTAL];     // final nodes in output order: depth-p* entry + 1
    int nonEmpty[OCTF_MAX_DEPTH + 1];
    int warpSums[OCTF_THREADS / 32];
    int pstar, K, nvalid;
};

__global__ void __launch_bounds__(OCTF_THREADS, 2) k_octree_fast(const __grid_constant__ OrbPlan plan) {
    pdl_enter();
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    OctFastSmem& sm = *reinterpret_cast<OctFastSmem*>(smem_raw);
    const int lt = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
    const OrbLevel& T = plan.lv[lt];
    const OrbLevel& L = plan.lv[T.src];
    int n = plan.candCount[f * ORB_MAX_LEVELS + T.src];
    if (n > (int)L.candCap) n = (int)L.candCap;
    int* needGeneric = plan.needGeneric + f * ORB_MAX_LEVELS + lt;
    if (n == 0 || L.nIni <= 0) {
        if (tid == 0) {
            plan.keptCount[f * ORB_MAX_LEVELS + lt] = 0;
            *needGeneric = 0;
        }
        return;
    }
    const uint2* cand = L.cand + (size_t)f * L.candCap;
    unsigned* codes = reinterpret_cast<unsigned*>(T.sortScratch + (size_t)f * 2 * T.sortCap);
    const int nIni = L.nIni;
    // deepest table depth that fits, and table offsets off(t) = nIni * (4^t - 1) / 3
    int tmax = 0;
    while (tmax < OCTF_MAX_DEPTH && nIni * (1 << (2 * (tmax + 1))) <= OCTF_MAX_NODES &&
           nIni * (((1 << (2 * (tmax + 2))) - 1) / 3) <= OCTF_MAX_TOTAL)
        ++tmax;
    const int total = nIni * (((1 << (2 * (tmax + 1))) - 1) / 3);
    for (int i = tid; i < total; i += OCTF_THREADS) sm.cnt[i] = 0;
    if (tid <= OCTF_MAX_DEPTH) sm.nonEmpty[tid] = 0;
    __syncthreads();

    // ---- pass 1: codes + ancestor occupancy counts for depths 0..tmax
    for (int i = tid; i < n; i += OCTF_THREADS) {
        const uint2 c = cand[i];
        unsigned code = 0xffffffffu;
        if (path_code_to(L, (int)(c.x & 0xffff), (int)(c.x >> 16), tmax, code)) {
            int off = 0;
            for (int t = 0; t <= tmax; ++t) {
                atomicAdd(&sm.cnt[off + (code >> (2 * (tmax - t)))], 1u);
                off += nIni << (2 * t);
            }
        } else {
            code = 0xffffffffu;
        }
        codes[i] = code;
    }
    __syncthreads();
    // ---- Count_t
    {
        int local[OCTF_MAX_DEPTH + 1];
#pragma unroll
        for (int t = 0; t <= OCTF_MAX_DEPTH; ++t) local[t] = 0;
        int off = 0;
        for (int t = 0; t <= tmax; ++t) {
            const int sz = nIni << (2 * t);
            int c = 0;
            for (int i = tid; i < sz; i += OCTF_THREADS) c += sm.cnt[off + i] != 0;
            local[t] = c;
            off += sz;
        }
        int nv = 0;
        for (int i = tid; i < nIni; i += OCTF_THREADS) nv += (int)sm.cnt[i];
#pragma unroll
        for (int t = 0; t <= OCTF_MAX_DEPTH; ++t) {
            int v = local[t];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((tid & 31) == 0 && v) atomicAdd(&sm.nonEmpty[t], v);
        }
        for (int o = 16; o > 0; o >>= 1) nv += __shfl_xor_sync(0xffffffffu, nv, o);
        if (tid == 0) sm.nvalid = 0;
        __syncthreads();
        if ((tid & 31) == 0 && nv) atomicAdd(&sm.nvalid, nv);
    }
    __syncthreads();
    if (tid == 0) {
        int p = -1;
        for (int t = 1; t <= tmax; ++t)
            if (sm.nonEmpty[t] >= T.nFeat || sm.nonEmpty[t] == sm.nvalid) { p = t; break; }
        sm.pstar = p;
        sm.K = p >= 0 ? sm.nonEmpty[p] : 0;
    }
    __syncthreads();
    const int pstar = sm.pstar;
    if (sm.nvalid == 0) {
        if (tid == 0) {
            plan.keptCount[f * ORB_MAX_LEVELS + lt] = 0;
            *needGeneric = 0;
        }
        return;
    }
    if (pstar < 0) {  // deeper than the tables: the sort-based form, in place (block-uniform branch)
        __syncthreads();  // everybody has read pstar / nvalid: the tables' memory becomes the sort buffers
        if (tid == 0) *needGeneric = 0;
        octree_generic(plan, lt, f, smem_raw, *reinterpret_cast<OctShared*>(smem_raw + sizeof(OctFastSmem)));
        return;
    }
    const int nodes = nIni << (2 * pstar);
    const int offP = nIni * (((1 << (2 * pstar)) - 1) / 3);
    const int totalP = offP + nodes;  // tables of depth 0..p*
    for (int i = tid; i < nodes; i += OCTF_THREADS) sm.best[i] = 0ull;
    for (int i = tid; i < totalP; i += OCTF_THREADS) sm.ordTab[i] = 0;
    __syncthreads();
    // ---- pass 2: first max-response key per depth-p* node
    const unsigned long long ORDMAX = (1ull << 44) - 1;
    const unsigned wMagic = 0xffffffffu / (unsigned)L.wCell + 1u, hMagic = 0xffffffffu / (unsigned)L.hCell + 1u;
    for (int i = tid; i < n; i += OCTF_THREADS) {
        const unsigned code = codes[i];
        if (code == 0xffffffffu) continue;
        const uint2 c = cand[i];
        const unsigned long long key = ((unsigned long long)c.y << 48) | (ORDMAX - cand_order_magic(L, c.x & 0xffff, c.x >> 16, wMagic, hMagic));
        atomicMax(&sm.best[code >> (2 * (tmax - pstar))], key);
    }
    __syncthreads();
    // ---- final node of every non-empty depth-p* entry -> slot in the ordered tables
    for (int e = tid; e < nodes; e += OCTF_THREADS) {
        if (sm.cnt[offP + e] == 0) continue;
        int b = pstar, off = 0;
        for (int d = 0; d < pstar; ++d) {  // birth = first depth at which the key is alone
            if (sm.cnt[off + (e >> (2 * (pstar - d)))] == 1) { b = d; break; }
            off += nIni << (2 * d);
        }
        const unsigned pb = (unsigned)e >> (2 * (pstar - b));
        const unsigned cb = pb & ((1u << (2 * b)) - 1u);
        unsigned root = pb >> (2 * b);
        if (b >= 1 && (((b - 1) & 1) == 0)) root = (unsigned)nIni - 1u - root;  // root follows c_1's direction
        const unsigned tb = (root << (2 * b)) | (cb ^ (0x33333333u & ((1u << (2 * b)) - 1u)));
        // groups in output order: birth p*, p*-1, ..., 0
        int goff = 0;
        for (int d = pstar; d > b; --d) goff += nIni << (2 * d);
        sm.ordTab[goff + tb] = (unsigned short)(e + 1);
    }
    __syncthreads();
    // ---- rank = exclusive prefix count of occupied slots
    const int chunk = (totalP + OCTF_THREADS - 1) / OCTF_THREADS;
    const int beg = min(tid * chunk, totalP), end = min(beg + chunk, totalP);
    int mine = 0;
    for (int i = beg; i < end; ++i) mine += sm.ordTab[i] != 0;
    int incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += v;
    }
    if ((tid & 31) == 31) sm.warpSums[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        int v = tid < OCTF_THREADS / 32 ? sm.warpSums[tid] : 0;
        for (int o = 1; o < 32; o <<= 1) {
            const int t2 = __shfl_up_sync(0xffffffffu, v, o);
            if (tid >= o) v += t2;
        }
        if (tid < OCTF_THREADS / 32) sm.warpSums[tid] = v;
    }
    __syncthreads();
    int rank = incl - mine + ((tid >> 5) ? sm.warpSums[(tid >> 5) - 1] : 0);
    uint2* kept = T.kept + (size_t)f * T.kmax;
    for (int i = beg; i < end; ++i) {
        const int e1 = sm.ordTab[i];
        if (!e1) continue;
        const unsigned long long key = sm.best[e1 - 1];
        const unsigned long long ord = ORDMAX - (key & ORDMAX);
        const unsigned x = (unsigned)(ord & 0x1fff) + ORB_MINB, y = (unsigned)((ord >> 13) & 0x1fff) + ORB_MINB;
        if (rank < T.kmax) kept[rank] = make_uint2(x | (y << 16), (unsigned)(key >> 48));
        ++rank;
    }
    if (tid == 0) {
        plan.keptCount[f * ORB_MAX_LEVELS + lt] = sm.K;
        *needGeneric = 0;
    }
}

// ------------------------------------------------------------------------------------------
// k_blur: cv::GaussianBlur 7x7 sigma 2, BORDER_REFLECT_101, 8UC1:
// out = (sum_j sum_i w_j w_i p + 32768) >> 16, w = [18,34,48,56,48,34,18].
// Tile of 256 x 64 outputs staged in shared memory (with the 3-px reflected halo).  Each
// thread owns a 4-pixel column group and walks 16 output rows: the horizontal pass is two
// IDP.4A per pixel on byte windows cut from three aligned words, the vertical pass runs on a
// 7-row register window.  Exact integers throughout, one 32-bit store per 4 pixels.
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}

// The tile with its 3-pixel halo is one TMA box load (box byte (r, c) = level pixel (y0 - 3 + r, x0 - 16 + c), zero outside
// the image); tiles on the image border then rebuild BORDER_REFLECT_101 in shared memory: whole rows above / below the
// image first, then the three columns left / right of it.
__global__ void __launch_bounds__(BLUR_THREADS) k_blur(const __grid_constant__ OrbPlan plan, const CUtensorMap* __restrict__ maps) {
    pdl_enter();
    __shared__ __align__(128) unsigned tile[BLUR_BOX_H][BLUR_SW];
    __shared__ unsigned long long barMem;
    const int f = blockIdx.y;
    if ((int)blockIdx.x >= plan.blurTiles) return;
    const unsigned te = __ldg(plan.blurTileTab + blockIdx.x);  // level | tile row << 4 | tile column << 18
    const int l = (int)(te & 15u), ty = (int)((te >> 4) & 0x3fffu), tx = (int)(te >> 18);
    const OrbLevel& L = plan.lv[l];
    const int x0 = tx * BLUR_TW, y0 = ty * BLUR_TH;
    const int tid = threadIdx.x;
    const int rowsHere = min(BLUR_TH, L.rows - y0) + 6;
    const unsigned bar = smem_u32(&barMem);
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, (unsigned)sizeof(tile));
        tma_load_3d(smem_u32(&tile[0][0]), maps + l, x0 - 16, y0 - 3, f + plan.frameBase, bar);
    }
    mbar_wait(bar, 0);
    const bool rowFix = y0 == 0 || y0 - 3 + rowsHere > L.rows;
    const bool colFix = x0 == 0 || x0 + BLUR_TW + 3 > L.cols;
    if (rowFix) {  // rows above / below the image <- their mirror rows (inside the tile: at most 3 rows away from the edge)
        // only the rows outside the image are visited: box rows 0..2 of the top tiles (level rows -3..-1) and the box rows
        // from firstBot on (level rows >= L.rows) of the bottom tiles
        const int nTop = y0 == 0 ? 3 : 0;
        const int firstBot = max(L.rows - (y0 - 3), nTop);
        const int nFix = nTop + max(0, rowsHere - firstBot);
        static_assert(BLUR_SW == 64, "row index by shift");
        for (int i = tid; i < nFix * BLUR_SW; i += BLUR_THREADS) {
            const int k = i >> 6, w = i & 63;
            const int r = k < nTop ? k : firstBot + (k - nTop);
            const int rs = reflect101(y0 - 3 + r, L.rows) - (y0 - 3);
            if (rs >= 0 && rs < BLUR_BOX_H) tile[r][w] = tile[rs][w];
        }
        __syncthreads();
    }
    if (colFix) {  // the three columns left / right of the image <- their mirror columns
        unsigned char* tb = reinterpret_cast<unsigned char*>(&tile[0][0]);
        for (int i = tid; i < rowsHere * 6; i += BLUR_THREADS) {
            const int r = i / 6, k = i - r * 6;
            const int x = k < 3 ? k - 3 : L.cols + (k - 3);
            const int c = x - (x0 - 16);
            if (c >= 0 && c < 4 * BLUR_SW) {
                int xs = x < 0 ? -x : 2 * L.cols - 2 - x;  // one reflection is enough within 3 px
                xs = min(max(xs, 0), L.cols - 1);          // (degenerate tiny levels)
                const int cs = xs - (x0 - 16);
                if (cs >= 0 && cs < 4 * BLUR_SW) tb[r * (4 * BLUR_SW) + c] = tb[r * (4 * BLUR_SW) + cs];
            }
        }
    }
    if (rowFix || colFix) __syncthreads();
    const int q = tid % (BLUR_TW / 4), g = tid / (BLUR_TW / 4);
    const int xq = x0 + 4 * q;
    if (xq >= L.cols) return;
    const int rbase = g * BLUR_RPT;  // first output row of this thread inside the tile
    if (y0 + rbase >= L.rows) return;
    uint8_t* dstp = L.blur + (size_t)f * L.plane + (size_t)(y0 + rbase) * L.pitch + xq;  // output row of this thread, bumped per row
    const unsigned WLO = 18u | (34u << 8) | (48u << 16) | (56u << 24);
    const unsigned WHI = 48u | (34u << 8) | (18u << 16);
    // Vertical pass on row PAIRS: the horizontal sums are < 2^16, so two consecutive rows of one pixel share a register
    // (P[r] = H[r] | H[r+1] << 16) and the 7 taps are three IDP.2A plus one multiply-add:
    //   acc(y) = 32768 + (18, 34).P[y] + (48, 56).P[y+2] + (48, 34).P[y+4] + 18 H[y+6];   out = byte 2 of acc.
    const unsigned W01 = 18u | (34u << 8), W23 = 48u | (56u << 8), W45 = 48u | (34u << 8);
    // Tiles whose BLUR_TH rows all exist (71 % of the pixels of a KITTI pyramid) run the loop without its two per-row range
    // tests, 12 % of the kernel's instructions.  The choice is per CTA: a per-thread choice splits the warps that straddle
    // two row groups of a partial tile, and they then run both forms (measured: 104.9 k instead of 106.2 k frames/s).
    const bool allRows = y0 + BLUR_TH <= L.rows;
    auto rows_loop = [&](auto fullTag) {
        constexpr bool FULL = decltype(fullTag)::value;
        unsigned P[6][4], Hprev[4] = {0u, 0u, 0u, 0u};
        uint8_t* out = dstp;
#pragma unroll
        for (int rr = 0; rr < BLUR_RPT + 6; ++rr) {
            const int r = rbase + rr;
            unsigned h[4] = {0u, 0u, 0u, 0u};
            if (FULL || r < rowsHere) {
                // words q+3, q+4, q+5 = smem bytes 4q+12 .. 4q+23 = b0..b11; output pixel k (smem byte
                // 16 + 4q + k) reads b(1+k)..b(7+k)
                const unsigned w0 = tile[r][q + 3], w1 = tile[r][q + 4], w2 = tile[r][q + 5];
                h[0] = __dp4a(__byte_perm(w0, w1, 0x4321), WLO, __dp4a(__byte_perm(w1, w2, 0x4321), WHI, 0u));
                h[1] = __dp4a(__byte_perm(w0, w1, 0x5432), WLO, __dp4a(__byte_perm(w1, w2, 0x5432), WHI, 0u));
                h[2] = __dp4a(__byte_perm(w0, w1, 0x6543), WLO, __dp4a(__byte_perm(w1, w2, 0x6543), WHI, 0u));
                h[3] = __dp4a(w1, WLO, __dp4a(w2, WHI, 0u));
            }
            if (rr >= 1) {
#pragma unroll
                for (int k = 0; k < 4; ++k) P[(rr - 1) % 6][k] = __byte_perm(Hprev[k], h[k], 0x5410);
            }
            if (rr >= 6) {
                const int y = y0 + rbase + rr - 6;
                if (FULL || y < L.rows) {
                    unsigned acc[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        // rows rr-6 .. rr of the window, taps 18 34 48 56 48 34 18
                        unsigned a = __dp2a_lo(P[(rr - 6) % 6][k], W01, 32768u);
                        a = __dp2a_lo(P[(rr - 4) % 6][k], W23, a);
                        a = __dp2a_lo(P[(rr - 2) % 6][k], W45, a);
                        acc[k] = a + 18u * h[k];
                    }
                    // byte 2 of every accumulator (acc < 2^24)
                    const unsigned outw = __byte_perm(__byte_perm(acc[0], acc[1], 0x0062), __byte_perm(acc[2], acc[3], 0x0062), 0x5410);
                    *reinterpret_cast<unsigned*>(out) = outw;
                    out += L.pitch;
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) Hprev[k] = h[k];
        }
    };
    if (allRows)
        rows_loop(std::true_type{});
    else
        rows_loop(std::false_type{});
}

// ------------------------------------------------------------------------------------------
// k_describe: IC_Angle (ORBextractor.cc:21-48) on the raw level, rBRIEF
// (computeOrbDescriptor :57-73) on the blurred level, final keypoint record
// (:345-352, :486-491).  One warp per kept keypoint.
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    // cv::fastAtan2, evaluated step by step in binary32, no FMA
    const float scale = (float)(180.0 / 3.141592653589793238462643383279502884);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    const float eps = (float)2.2204460492503131e-16;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// cvRound (round half to even) of a float with |v| < 2^22 without the conversion unit: adding
// 1.5 * 2^23 leaves rint(v) in the low mantissa bits under round-to-nearest-even.
__device__ __forceinline__ int cv_round_small(float v) {
    return __float_as_int(__fadd_rn(v, 12582912.f)) - 0x4B400000;
}

// sin and cos of x in [0, 2*pi] in double precision (fdlibm kernels after a two-constant reduction
// by pi/2; error ~1 ulp of double), so that rounding to float gives the correctly rounded
// cosf / sinf the reference gets from libm.  Much shorter than the general sincos().
__device__ __forceinline__ void sincos_0_2pi(double x, double& s, double& c) {
    const int q = __double2int_rn(x * 0.63661977236758134308);  // x * 2/pi
    const double qd = (double)q;
    double r = fma(-qd, 1.57079632679489655800e+00, x);
    r = fma(-qd, 6.12323399573676603587e-17, r);
    const double z = r * r;
    double ps = 1.58969099521155010221e-10;
    ps = fma(ps, z, -2.50507602534068634195e-08);
    ps = fma(ps, z, 2.75573137070700676789e-06);
    ps = fma(ps, z, -1.98412698298579493134e-04);
    ps = fma(ps, z, 8.33333333332248946124e-03);
    ps = fma(ps, z, -1.66666666666666324348e-01);
    const double sr = fma(r * z, ps, r);
    double pc = -1.13596475577881948265e-11;
    pc = fma(pc, z, 2.08757232129817482790e-09);
    pc = fma(pc, z, -2.75573143513906633035e-07);
    pc = fma(pc, z, 2.48015872894767294178e-05);
    pc = fma(pc, z, -1.38888888888741095749e-03);
    pc = fma(pc, z, 4.16666666666666019037e-02);
    const double cr = fma(z * z, pc, fma(-0.5, z, 1.0));
    const double s1 = (q & 1) ? cr : sr, c1 = (q & 1) ? sr : cr;
    s = (q & 2) ? -s1 : s1;
    c = ((q + 1) & 2) ? -c1 : c1;
}

// Output index o of frame f -> (level, rank inside the level's kept list).  Levels are concatenated 0..n-1
// (ORBextractor.cc:466-494); lane j of the calling warp holds `pre` = keypoints of the levels before level j and
// `myK` = its own count.  Warp-uniform arguments and result.
__device__ __forceinline__ void locate_keypoint(int o, int pre, int myK, int nlevels, int lane, int& l, int& r) {
    // the last level whose prefix is <= o (empty levels share a prefix)
    const unsigned le = __ballot_sync(0xffffffffu, lane < nlevels && pre <= o && myK > 0);
    l = 31 - __clz(le);
    r = o - __shfl_sync(0xffffffffu, pre, l);
}

__device__ __forceinline__ int level_prefix(const int* kc, int nlevels, int lane, int cap, int& myK, int& total) {
    myK = lane < nlevels ? kc[lane] : 0;
    int pre = myK;
    for (int sft = 1; sft < 32; sft <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, pre, sft);
        if (lane >= sft) pre += v;
    }
    total = min(__shfl_sync(0xffffffffu, pre, 31), cap);
    return pre - myK;  // exclusive prefix
}

// ------------------------------------------------------------------------------------------
// k_describe_tile: orientation, angle, rBRIEF and the final record for the keypoints of one describe tile.
//
// One CTA = one tile of one source level of one frame (DSC_W x DSC_H keypoint positions).  The raw and the
// blurred level around it (18-px halo) are staged by one TMA box load each, so every level pixel crosses
// L2 -> SM about twice per frame instead of once per keypoint patch that covers it (the patches of a
// KITTI-shape frame cover the pyramid ~5x).  The CTA then
//   scan     walks the kept lists of the levels that share this source (a level and its same-size alias) and
//            collects the keypoints inside its core -- no binning pass, the lists are a few KB;
//   phase 1  warp per keypoint: IC_Angle moments (ORBextractor.cc:21-48) from the raw tile.  The patch row is
//            brought to a canonical alignment with a funnel shift, so the dp4a weights are keypoint
//            independent and live in registers (lane = row % 4, word 0..7; 8 steps cover the 31 rows);
//   phase 2  thread per keypoint: cv::fastAtan2 and the cos / sin of the descriptor rotation (:59-60);
//   phase 3  warp per keypoint: the 182 rBRIEF tests (:57-73) gathered from the blurred tile, bits packed by
//            ballot, and the final record (:345-352, :486-491).
// A tile holding more than DSC_LIST keypoints (dense adversarial input) is processed in rounds over slices of
// the kept lists.
// ------------------------------------------------------------------------------------------
struct DescTileSmem {
    unsigned xy[DSC_LIST];        // x | y << 16, level coordinates
    unsigned resp[DSC_LIST];      // FAST response
    unsigned lo[DSC_LIST];        // level << 24 | output index
    unsigned twin[DSC_LIST];      // 1 + output index of the same keypoint in the alias level's list (0: none)
    int m01[DSC_LIST], m10[DSC_LIST];
    float angle[DSC_LIST], ca[DSC_LIST], sb[DSC_LIST];
    int levels[ORB_MAX_LEVELS];   // levels whose keypoints live on this source's pixels,
    int lvCount[ORB_MAX_LEVELS];  // their kept counts
    int lvBase[ORB_MAX_LEVELS];   // and the output index of their first keypoint
    int nLevels;
    int count;
    unsigned long long bar;
};
static const size_t kDescTileBytes = (size_t)DSC_BOX_W * DSC_BOX_H;
static const size_t kDescSmem = kDescTileBytes + sizeof(DescTileSmem);

__device__ __forceinline__ unsigned lds_u8(unsigned addr) {
    unsigned v;
    // not volatile, no clobber: the address always depends on shared-memory values read after the barrier that
    // publishes the tile, so the load cannot move above it, and the compiler stays free to batch the gathers
    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__global__ void __launch_bounds__(DSC_THREADS, 4) k_describe_tile(const __grid_constant__ OrbPlan plan, const DetectMaps* __restrict__ maps,
                                                                  orb_keypoint_dev* __restrict__ kps, uint8_t* __restrict__ desc,
                                                                  int cap, int* __restrict__ counts) {
    pdl_enter();
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const unsigned char* tileT = smem_raw;  // the raw tile (phase 1), then the blurred tile (phase 3)
    DescTileSmem& sm = *reinterpret_cast<DescTileSmem*>(smem_raw + kDescTileBytes);
    const int f = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int* kc = plan.keptCount + f * ORB_MAX_LEVELS;
    if (blockIdx.x == 0 && tid == 0) {
        int tot = 0;
        for (int i = 0; i < plan.nlevels; ++i) tot += kc[i];
        counts[f] = tot;
    }
    if ((int)blockIdx.x >= plan.totalDescTiles) return;
    const unsigned te = __ldg(plan.descTileTab + blockIdx.x);  // source level | tile row << 4 | tile column << 18
    const int s = (int)(te & 15u), ty = (int)((te >> 4) & 0x3fffu), tx = (int)(te >> 18);
    const int bx0 = DSC_W * tx, by0 = 1 + DSC_H * ty;            // level pixel of box byte (0, 0)
    const unsigned xlo = ORB_EDGE + DSC_W * tx, ylo = ORB_EDGE + DSC_H * ty;  // core: [xlo, xlo + DSC_W) x [ylo, ylo + DSC_H)

    // ---- warp 0: output offsets of the levels (levels are concatenated 0..n-1, ORBextractor.cc:466-494), the levels
    // that live on this source's pixels, their keypoint counts
    const unsigned bar = smem_u32(&sm.bar);
    if (warp == 0) {
        int myK, total;
        const int pre = level_prefix(kc, plan.nlevels, lane, cap, myK, total);
        const bool mine = lane < plan.nlevels && plan.lv[lane].src == s;
        const unsigned onSrc = __ballot_sync(0xffffffffu, mine);
        if (mine) {
            const int slot = __popc(onSrc & ((1u << lane) - 1u));
            sm.levels[slot] = lane;
            sm.lvCount[slot] = min(myK, plan.lv[lane].kmax);
            sm.lvBase[slot] = pre;
        }
        if (lane == 0) {
            sm.nLevels = __popc(onSrc);
            sm.count = 0;
            mbar_init(bar, 1);
        }
    }
    __syncthreads();
    const int nLv = sm.nLevels;
    {
        int nOnSrc = 0;
        for (int j = 0; j < nLv; ++j) nOnSrc += sm.lvCount[j];
        if (nOnSrc == 0) return;  // uniform over the CTA: nothing was issued yet
    }
    unsigned parity = 0;  // phase of the mbarrier the next TMA load completes
    bool sliced = false;
    int sl = 0, sr0 = 0;  // sliced mode: level slot and first rank of the current slice
    for (;;) {
        // the raw tile travels while the lists are scanned (every earlier read of the buffer is behind a barrier)
        if (tid == 0) {
            fence_proxy_async();
            mbar_expect_tx(bar, (unsigned)kDescTileBytes);
            tma_load_3d(smem_u32(tileT), &maps->raw[s], bx0, by0, f + plan.frameBase, bar);
        }
        // ---- scan: keypoints of this tile (whole lists, or one slice of DSC_LIST ranks of one level)
        for (int j = sliced ? sl : 0; j < (sliced ? sl + 1 : nLv); ++j) {
            const int l = sm.levels[j];
            const OrbLevel& L = plan.lv[l];
            const int nK = sm.lvCount[j];
            const int base = sm.lvBase[j];
            const uint2* kept = L.kept + (size_t)f * L.kmax;
            const int r1 = sliced ? min(nK, sr0 + DSC_LIST) : nK;
            // A level and its same-size alias (levels 0 and 1 of this fork: the scale table starts {1, 1, ...}) run
            // their octrees on the same candidates; when both stop at the same depth their kept lists are equal entry by
            // entry.  Rank r of the source level then also serves rank r of the alias (its "twin"): orientation and descriptor
            // are computed once and written twice.  dup(r) is evaluated identically from both sides.
            const bool twinSide = nLv > 1 && j < 2;
            const uint2* other = nullptr;
            int nOther = 0, twinBase = 0;
            if (twinSide) {
                const OrbLevel& Lo = plan.lv[sm.levels[1 - j]];
                other = Lo.kept + (size_t)f * Lo.kmax;
                nOther = sm.lvCount[1 - j];
                twinBase = sm.lvBase[1];
            }
            for (int r = (sliced ? sr0 : 0) + tid; r < r1; r += DSC_THREADS) {
                const uint2 k = kept[r];
                const unsigned x = k.x & 0xffffu, y = k.x >> 16;
                if (x - xlo < (unsigned)DSC_W && y - ylo < (unsigned)DSC_H && base + r < cap) {
                    bool dup = false;
                    if (twinSide && r < nOther && twinBase + r < cap) {
                        const uint2 ko = other[r];
                        dup = ko.x == k.x && ko.y == k.y;
                    }
                    if (dup && j == 1) continue;  // served by the source level's entry
                    const int slot = atomicAdd(&sm.count, 1);
                    if (slot < DSC_LIST) {
                        sm.xy[slot] = k.x;
                        sm.resp[slot] = k.y;
                        sm.lo[slot] = ((unsigned)l << 24) | (unsigned)(base + r);
                        sm.twin[slot] = dup ? (unsigned)(twinBase + r + 1) : 0u;
                    }
                }
            }
        }
        __syncthreads();
        const int cnt = sm.count;
        mbar_wait(bar, parity);  // the raw tile has landed (also: never leave with a load in flight)
        parity ^= 1u;
        if (!sliced && cnt > DSC_LIST) {  // too many for one round: start over, slice by slice
            sliced = true;
            sl = 0;
            sr0 = 0;
            __syncthreads();
            if (tid == 0) sm.count = 0;
            __syncthreads();
            continue;
        }
        if (cnt > 0) {
            // ---- phase 1: moments.  lane (r4, j) = (lane / 8, lane % 8): canonical word j of patch rows it*4 + r4
            {
                int2 icw[8];  // IC_Angle weights of this lane (keypoint independent)
#pragma unroll
                for (int it = 0; it < 8; ++it) icw[it] = __ldg(plan.icTab + it * 32 + lane);
                const int r4 = lane >> 3, j = lane & 7;
                for (int i = warp; i < cnt; i += DSC_THREADS / 32) {
                    const unsigned xy = sm.xy[i];
                    const int px = (int)(xy & 0xffffu) - bx0 - 15, py = (int)(xy >> 16) - by0 - 15;  // patch origin in the box
                    const unsigned sh = (unsigned)(px & 3) * 8u;
                    const unsigned* p = reinterpret_cast<const unsigned*>(tileT) + (py + r4) * (DSC_BOX_W / 4) + (px >> 2) + j;
                    int m10 = 0, m01 = 0;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {  // row 31 (it = 7, r4 = 3) carries zero weights; it is inside the box
                        const unsigned w = __funnelshift_r(p[0], p[1], sh);
                        m10 = dp4a_us(w, icw[it].x, m10);
                        m01 = dp4a_us(w, icw[it].y, m01);
                        p += 4 * (DSC_BOX_W / 4);
                    }
#pragma unroll
                    for (int sft = 16; sft > 0; sft >>= 1) {
                        m10 += __shfl_xor_sync(0xffffffffu, m10, sft);
                        m01 += __shfl_xor_sync(0xffffffffu, m01, sft);
                    }
                    if (lane == 0) {
                        sm.m01[i] = m01;
                        sm.m10[i] = m10;
                    }
                }
            }
            __syncthreads();
            // the blurred tile replaces the raw one while the angles are computed
            if (tid == 0) {
                fence_proxy_async();
                mbar_expect_tx(bar, (unsigned)kDescTileBytes);
                tma_load_3d(smem_u32(tileT), &maps->blur[s], bx0, by0, f + plan.frameBase, bar);
            }
            // ---- phase 2: angle, cos, sin -- one thread per keypoint
            for (int i = tid; i < cnt; i += DSC_THREADS) {
                const float angle = fast_atan2_deg((float)sm.m01[i], (float)sm.m10[i]);
                const float factorPI = (float)(3.141592653589793238462643383279502884 / 180.f);
                const float rad = __fmul_rn(angle, factorPI);
                double sd, cd;
                sincos_0_2pi((double)rad, sd, cd);
                sm.angle[i] = angle;
                sm.ca[i] = (float)cd;
                sm.sb[i] = (float)sd;
            }
            __syncthreads();
            mbar_wait(bar, parity);
            parity ^= 1u;
            // ---- phase 3: rBRIEF on the blurred tile + final record
            {
                // the 182 test pairs: lane i holds pairs i, i + 32, ... (keypoint independent)
                float4 pr[6];
#pragma unroll
                for (int wq = 0; wq < 6; ++wq)
                    pr[wq] = (wq * 32 + lane < 182) ? __ldg(plan.pairTab + wq * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
                for (int i = warp; i < cnt; i += DSC_THREADS / 32) {
                    const unsigned xy = sm.xy[i];
                    const int x = (int)(xy & 0xffffu), y = (int)(xy >> 16);
                    const float a = sm.ca[i], b = sm.sb[i];
                    // cvRound by the 1.5 * 2^23 trick (cv_round_small) with the constant's integer image folded into the
                    // base address: byte (r, c) of the patch = pb[R * DSC_BOX_W + C], R = float_as_int(r + magic) etc.
                    // (32-bit shared-window addresses, wrap-around arithmetic)
                    const unsigned pb = smem_u32(tileT) + (unsigned)((y - by0) * DSC_BOX_W + (x - bx0)) - 0x4B400000u * (unsigned)(DSC_BOX_W + 1);
                    unsigned myWord = 0;  // lane i < 8 ends up holding descriptor word i (words 6, 7 are zero)
#pragma unroll
                    for (int wq = 0; wq < 6; ++wq) {
                        // pairs beyond 181 are (0,0)-(0,0): t0 == t1, bit 0 -- like the fork's zero-filled pattern tail
                        const unsigned r0 = (unsigned)__float_as_int(__fadd_rn(__fadd_rn(__fmul_rn(pr[wq].x, b), __fmul_rn(pr[wq].y, a)), 12582912.f));
                        const unsigned c0 = (unsigned)__float_as_int(__fadd_rn(__fsub_rn(__fmul_rn(pr[wq].x, a), __fmul_rn(pr[wq].y, b)), 12582912.f));
                        const unsigned r1 = (unsigned)__float_as_int(__fadd_rn(__fadd_rn(__fmul_rn(pr[wq].z, b), __fmul_rn(pr[wq].w, a)), 12582912.f));
                        const unsigned c1 = (unsigned)__float_as_int(__fadd_rn(__fsub_rn(__fmul_rn(pr[wq].z, a), __fmul_rn(pr[wq].w, b)), 12582912.f));
                        const unsigned t0 = lds_u8(pb + r0 * DSC_BOX_W + c0);
                        const unsigned t1 = lds_u8(pb + r1 * DSC_BOX_W + c1);
                        const unsigned wbits = __ballot_sync(0xffffffffu, t0 < t1);
                        if ((lane & 7) == wq) myWord = wbits;
                    }
                    const unsigned lo = sm.lo[i], twin = sm.twin[i];
                    // lanes 0-7 / 16: descriptor and record of the keypoint; lanes 8-15 / 17: of its twin in the alias level
                    const bool second = (lane & 8) || lane == 17;
                    const int o = second ? (int)twin - 1 : (int)(lo & 0xffffffu), l = second ? sm.levels[1] : (int)(lo >> 24);
                    if (lane < 16 && o >= 0) reinterpret_cast<unsigned*>(desc + ((size_t)f * cap + o) * 32)[lane & 7] = myWord;
                    if ((lane == 16 || lane == 17) && o >= 0) {
                        const OrbLevel& L = plan.lv[l];
                        const float sc = l != 0 ? L.scale : 1.0f;  // keypoint.pt *= scale for level != 0 (:486-491)
                        orb_keypoint_dev kp;
                        kp.x = __fmul_rn((float)x, sc);
                        kp.y = __fmul_rn((float)y, sc);
                        kp.size = (float)L.patchSize;
                        kp.angle = sm.angle[i];
                        kp.response = (float)sm.resp[i];
                        kp.octave = l;
                        kp.class_id = -1;
                        kps[(size_t)f * cap + o] = kp;
                    }
                }
            }
        }
        if (!sliced) break;
        // next slice
        {
            const int nK = sm.lvCount[sl];
            sr0 += DSC_LIST;
            if (sr0 >= nK) {
                sr0 = 0;
                ++sl;
            }
        }
        if (sl >= nLv) break;
        __syncthreads();  // phase 3 of every warp is done with the tile and the list
        if (tid == 0) sm.count = 0;
        __syncthreads();
    }
}

// The 182 live rBRIEF test pairs as float4 (x0, y0, x1, y1) for plan.pairTab.
void orbk_build_pair_table(float4* out) {
    static const int8_t pairs[728] = {
#include "brief_pairs_182.inc"
    };
    for (int i = 0; i < 182; ++i)
        out[i] = make_float4((float)pairs[4 * i], (float)pairs[4 * i + 1], (float)pairs[4 * i + 2], (float)pairs[4 * i + 3]);
}

// Host-side construction of the IC_Angle weight table: entry [it][lane] for patch row vr = it * 4 + lane / 8 and
// canonical word j = lane % 8 (patch columns u = 4j - 15 .. 4j - 12): .x = four signed bytes u, .y = four signed
// bytes v = vr - 15, both 0 outside the disc |u| <= umax[|v|] (ORBextractor.cc:155-169) and on the padding row 31.
// m10 += dp4a(pixels, .x); m01 += dp4a(pixels, .y).
void orbk_build_ic_table(int2* out) {
    static const int umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
    for (int it = 0; it < 8; ++it)
        for (int lane = 0; lane < 32; ++lane) {
            const int vr = it * 4 + lane / 8, j = lane % 8;
            unsigned wu = 0, wm = 0;
            if (vr < 31) {
                const int v = vr - 15, av = v < 0 ? -v : v;
                for (int b = 0; b < 4; ++b) {
                    const int u = 4 * j + b - 15;
                    const int au = u < 0 ? -u : u;
                    if (au <= umax[av]) {
                        wu |= (unsigned)(u & 0xff) << (8 * b);
                        wm |= (unsigned)(v & 0xff) << (8 * b);
                    }
                }
            }
            out[it * 32 + lane] = make_int2((int)wu, (int)wm);
        }
}

// ------------------------------------------------------------------------------------------
// k_ingest: the step before the path, fused into the level-0 load: cv::remap(INTER_LINEAR, CV_32FC1 maps,
// BORDER_CONSTANT 0) of the raw frame (reference Examples/Stereo/stereo_euroc.cc:136-137) and / or
// cv::cvtColor RGB/BGR(A) -> gray (src/Tracking.cc:118-126), written straight into the pitched level-0 buffer,
// so the rectified / gray image never makes its own round trip through HBM.
// remap: sx = cvRound(mapx * 32), (ix, fx) = (sx >> 5, sx & 31); taps blended with cvRound((1-fy)(1-fx) * 32768) ...
// (exact multiples of 32; (0,0) is {32767, 0, 0, 1} after the saturation to short), (sum + 2^14) >> 15.
// gray: variant 4 (R*9798 + G*19235 + B*3735 + 2^14) >> 15, variant 3 (R*4899 + G*9617 + B*1868 + 2^13) >> 14.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int ingest_pixel(const uint8_t* __restrict__ S, int srows, int scols, size_t sstride, int channels, int bgr,
                                            int variant, const float* __restrict__ mapx, const float* __restrict__ mapy, int dcols,
                                            int x, int y) {
    int v[3] = {0, 0, 0};
    const int nch = channels >= 3 ? 3 : 1;
    if (mapx) {
        const int sx = __float2int_rn(__fmul_rn(mapx[(size_t)y * dcols + x], 32.0f));
        const int sy = __float2int_rn(__fmul_rn(mapy[(size_t)y * dcols + x], 32.0f));
        const int fx = sx & 31, fy = sy & 31;
        const int ix = max(-32768, min(32767, sx >> 5)), iy = max(-32768, min(32767, sy >> 5));
        int w0 = (32 - fy) * (32 - fx) * 32, w3 = fy * fx * 32;
        const int w1 = (32 - fy) * fx * 32, w2 = fy * (32 - fx) * 32;
        if (w0 == 32768) {
            w0 = 32767;
            w3 = 1;
        }
        const bool x0 = ix >= 0 && ix < scols, x1 = ix + 1 >= 0 && ix + 1 < scols;
        const bool y0 = iy >= 0 && iy < srows, y1 = iy + 1 >= 0 && iy + 1 < srows;
        const uint8_t* r0 = S + (size_t)(y0 ? iy : 0) * sstride;
        const uint8_t* r1 = S + (size_t)(y1 ? iy + 1 : 0) * sstride;
        const size_t c0 = (size_t)(x0 ? ix : 0) * channels, c1 = (size_t)(x1 ? ix + 1 : 0) * channels;
        for (int ch = 0; ch < nch; ++ch) {
            const int t00 = (x0 && y0) ? r0[c0 + ch] : 0, t01 = (x1 && y0) ? r0[c1 + ch] : 0;
            const int t10 = (x0 && y1) ? r1[c0 + ch] : 0, t11 = (x1 && y1) ? r1[c1 + ch] : 0;
            v[ch] = min(255, max(0, (t00 * w0 + t01 * w1 + t10 * w2 + t11 * w3 + (1 << 14)) >> 15));
        }
    } else {
        const uint8_t* p = S + (size_t)y * sstride + (size_t)x * channels;
        for (int ch = 0; ch < nch; ++ch) v[ch] = p[ch];
    }
    int g = v[0];
    if (nch == 3) {
        const int r = bgr ? v[2] : v[0], b = bgr ? v[0] : v[2];
        g = variant == 3 ? (r * 4899 + v[1] * 9617 + b * 1868 + (1 << 13)) >> 14 : (r * 9798 + v[1] * 19235 + b * 3735 + (1 << 14)) >> 15;
    }
    return g;
}

// four output pixels per thread, one aligned 32-bit store (the level-0 pitch is a multiple of 64; columns past
// dcols inside the last word are written as 0)
__global__ void __launch_bounds__(256) k_ingest(const uint8_t* __restrict__ raw, int srows, int scols, size_t sstride, size_t sframe,
                                                int channels, int bgr, int variant, const float* __restrict__ mapx,
                                                const float* __restrict__ mapy, int drows, int dcols, uint8_t* __restrict__ dst,
                                                int dpitch, unsigned long long dplane) {
    pdl_enter();
    const int x4 = (blockIdx.x * 64 + threadIdx.x) * 4, y = blockIdx.y * 4 + threadIdx.y, f = blockIdx.z;
    if (x4 >= dcols || y >= drows) return;
    const uint8_t* S = raw + (size_t)f * sframe;
    uint32_t word = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (x4 + i < dcols)
            word |= (uint32_t)ingest_pixel(S, srows, scols, sstride, channels, bgr, variant, mapx, mapy, dcols, x4 + i, y) << (8 * i);
    *reinterpret_cast<uint32_t*>(dst + (size_t)f * dplane + (size_t)y * dpitch + x4) = word;
}

cudaError_t orbk_ingest(const uint8_t* raw, int nframes, int srows, int scols, size_t sstride, size_t sframe, int channels, int bgr,
                        int variant, const float* mapx, const float* mapy, int drows, int dcols, uint8_t* dst, int dpitch,
                        unsigned long long dplane, cudaStream_t st) {
    dim3 block(64, 4), grid((dcols + 255) / 256, (drows + 3) / 4, nframes);
    k_ingest<<<grid, block, 0, st>>>(raw, srows, scols, sstride, sframe, channels, bgr, variant, mapx, mapy, drows, dcols, dst, dpitch, dplane);
    ++g_launches;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// k_repitch: densely packed host frames (row stride == cols), copied to the device with ONE
// linear transfer per chunk (row-by-row 2D copies are several times slower over PCIe), are laid
// out with the internal 64-byte aligned pitch here.  4 bytes per thread, funnel-shifted loads.
// ------------------------------------------------------------------------------------------
// `dense` is the 4-byte aligned base of the whole landing buffer and frame0 the first frame of this chunk inside it: a
// chunk's own first byte (frame0 * rows * cols) need not be word aligned (odd-area frames), so the word index and the
// funnel shift are both taken from the offset relative to the aligned base.
__global__ void __launch_bounds__(256) k_repitch(const uint8_t* __restrict__ dense, int frame0, int rows, int cols,
                                                 uint8_t* __restrict__ dst, int pitch, unsigned long long plane, size_t nwords) {
    pdl_enter();
    const int k = blockIdx.x * 256 + threadIdx.x;  // output word in the row
    const int y = blockIdx.y, f = blockIdx.z;
    if (4 * k >= cols) return;
    const size_t off = ((size_t)(frame0 + f) * rows + y) * cols + 4 * (size_t)k;  // byte offset in the dense buffer
    const unsigned* w = reinterpret_cast<const unsigned*>(dense) + (off >> 2);
    // nwords = 32-bit words of the dense buffer: the word behind the last one is not read (a caller's buffer has no slack)
    const unsigned lo = __ldg(w), hi = (off >> 2) + 1 < nwords ? __ldg(w + 1) : 0u;
    *reinterpret_cast<unsigned*>(dst + f * plane + (size_t)y * pitch + 4 * k) = __funnelshift_r(lo, hi, (unsigned)(off & 3) * 8);
}

cudaError_t orbk_repitch(const uint8_t* dense, int frame0, int nframes, int rows, int cols, uint8_t* dst, int pitch,
                         unsigned long long plane, cudaStream_t st) {
    if ((uintptr_t)dense & 3) return cudaErrorMisalignedAddress;
    const size_t nwords = ((size_t)(frame0 + nframes) * rows * cols + 3) / 4;
    k_repitch<<<dim3((cols + 1023) / 1024, rows, nframes), 256, 0, st>>>(dense, frame0, rows, cols, dst, pitch, plane, nwords);
    orbk_count_launch(1);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
cudaError_t orbk_encode_level_map(CUtensorMap* out, const uint8_t* base, int cols, int rows, int frames, int pitch,
                                  unsigned long long plane, int boxW, int boxH) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static const EncodeFn fn = []() -> EncodeFn {  // looked up once, thread-safe
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        return (EncodeFn)p;
    }();
    if (!fn) return cudaErrorNotSupported;
    if (((uintptr_t)base & 15) || (pitch & 15) || (plane & 15)) return cudaErrorMisalignedAddress;
    cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)(frames > 0 ? frames : 1)};
    cuuint64_t gstr[2] = {(cuuint64_t)pitch, (cuuint64_t)plane};
    cuuint32_t box[3] = {(cuuint32_t)boxW, (cuuint32_t)boxH, 1u};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

unsigned long long orbk_launch_count() { return g_launches.load(); }
void orbk_count_launch(int n) { g_launches += n; }

static_assert((size_t)(OCT_SMEM_A + OCT_SMEM_B) * 8 <= sizeof(OctFastSmem), "the sort buffers of octree_generic live in k_octree_fast's tables");
static const size_t kOctFastSmem = sizeof(OctFastSmem) + sizeof(OctShared);

cudaError_t orbk_init_device() {
    cudaError_t e = cudaFuncSetAttribute(k_detect, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)detect_smem_bytes(DET_TILE_H));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_describe_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDescSmem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_octree_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kOctFastSmem);
}

cudaError_t orbk_run_extract(const OrbPlan& plan, int nframes, orb_keypoint_dev* d_kps, uint8_t* d_desc, int cap,
                             int* d_counts, const OrbStreams& ss, const DetectMaps* d_maps, cudaEvent_t* ev) {
    // with per-stage events requested everything runs on one stream, so that every stage's
    // event-timed duration is its own (no overlap); otherwise the blur overlaps detect + octree
    cudaStream_t st = ss.st, st2 = ss.st2;
    cudaError_t e;
#ifdef ORB_B200_STAGE_KNOCKOUT  // timing probe (tools/probes/resident_probe.py LATE_ENV): stages left out once the buffers are filled
    const char* ko = getenv("ORB_B200_SKIP");
    const int skip = ko ? atoi(ko) : 0;  // 1 pyramid, 2 detect, 4 octree, 8 blur, 16 describe
#else
    const int skip = 0;
#endif
    if (!(skip & 2)) {
    e = cudaMemsetAsync(plan.candCount, 0, sizeof(int) * ORB_MAX_LEVELS * nframes, st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(plan.status, 0, sizeof(int) * nframes, st);
    if (e != cudaSuccess) return e;
    }
    if (ev) cudaEventRecord(ev[0], st);
    // Level 0 is the input: its detect tiles do not need the pyramid.  Outside profiling the pyramid chain (six dependent,
    // shrinking launches that mostly wait) therefore runs on the second stream next to the level-0 detect, followed there by the
    // blur; the first stream picks the other levels up once the pyramid is complete.
    const int tiles0 = plan.lv[0].src == 0 ? plan.lv[0].nTiles : 0;
    const bool split = !ev && tiles0 > 0 && plan.lv[0].tileBase == 0;
    cudaStream_t stp = split ? st2 : st;  // where the pyramid is built
    if (split) {
        e = cudaEventRecord(ss.fork, st);
        if (e != cudaSuccess) return e;
        e = cudaStreamWaitEvent(st2, ss.fork, 0);
        if (e != cudaSuccess) return e;
        if (!(skip & 2)) {
            k_detect<<<dim3(tiles0, nframes), DET_THREADS, detect_smem_bytes(plan.detRows), st>>>(plan, d_maps->m, 0);
            ++g_launches;
        }
    }
    // pyramid: level l = resize(level l-1); same-size levels alias their source
    bool first = true;
    for (int l = 1; l < plan.nlevels; ++l) {
        const OrbLevel& D = plan.lv[l];
        if (D.src != l || (skip & 1)) continue;
        const OrbLevel& S = plan.lv[plan.lv[l - 1].src];
        if (D.xgrp && D.rszTiled) {
            dim3 grid((D.cols + RSZ_W - 1) / RSZ_W, (D.rows + RSZ_H - 1) / RSZ_H, nframes);
            if (first && split)  // first kernel behind an event wait: a plain launch
                k_resize_tile<<<grid, RSZ_THREADS, 0, stp>>>(&d_maps->rsz[l], plan.frameBase, D.img, D.pitch, D.plane, D.rows, D.cols, D.xgrp,
                                                             D.xcoef4, D.ytab, D.ycoef);
            else
                launch_pdl(k_resize_tile, grid, dim3(RSZ_THREADS), 0, stp, &d_maps->rsz[l], plan.frameBase, D.img, D.pitch, D.plane, D.rows,
                           D.cols, D.xgrp, D.xcoef4, D.ytab, D.ycoef);
        } else if (D.xgrp) {
            dim3 block(64, 4), grid((D.cols + 255) / 256, (D.rows + 15) / 16, nframes);
            k_resize4<<<grid, block, 0, stp>>>(S.img, S.pitch, S.plane, D.img, D.pitch, D.plane, D.rows, D.cols, D.xgrp,
                                               D.xcoef4, D.ytab, D.ycoef);
        } else {
            dim3 block(64, 4), grid((D.cols + 255) / 256, (D.rows + 3) / 4, nframes);
            k_resize<<<grid, block, 0, stp>>>(S.img, S.pitch, S.plane, D.img, D.pitch, D.plane, D.rows, D.cols, D.xtab,
                                              D.xcoef, D.ytab, D.ycoef);
        }
        first = false;
        ++g_launches;
    }
    if (ev) cudaEventRecord(ev[1], st);
    const int blurTiles = plan.blurTiles;
    if (!ev) {
        // the blur needs only the pyramid and overlaps detect + octree on the second stream
        if (split) {
            e = cudaEventRecord(ss.pyr, st2);
            if (e != cudaSuccess) return e;
        } else {
            e = cudaEventRecord(ss.fork, st);
            if (e != cudaSuccess) return e;
            e = cudaStreamWaitEvent(st2, ss.fork, 0);
            if (e != cudaSuccess) return e;
        }
        if (!(skip & 8)) {
            k_blur<<<dim3(blurTiles, nframes), BLUR_THREADS, 0, st2>>>(plan, d_maps->blr);
            ++g_launches;
        }
        e = cudaEventRecord(ss.join, st2);
        if (e != cudaSuccess) return e;
    }
    if (split) {
        if (plan.totalTiles > tiles0) {
            e = cudaStreamWaitEvent(st, ss.pyr, 0);
            if (e != cudaSuccess) return e;
            if (!(skip & 2)) {
                k_detect<<<dim3(plan.totalTiles - tiles0, nframes), DET_THREADS, detect_smem_bytes(plan.detRows), st>>>(plan, d_maps->m, tiles0);
                ++g_launches;
            }
        }
    } else if (plan.totalTiles > 0) {
        launch_pdl(k_detect, dim3(plan.totalTiles, nframes), dim3(DET_THREADS), detect_smem_bytes(plan.detRows), st, plan, d_maps->m, 0);
        ++g_launches;
    }
    if (ev) cudaEventRecord(ev[2], st);
    if (!(skip & 4)) {
        launch_pdl(k_octree_fast, dim3(plan.nlevels, nframes), dim3(OCTF_THREADS), kOctFastSmem, st, plan);
        ++g_launches;
    }
    if (ev) {
        // profiling: stages back to back on one stream, blur after the octree
        cudaEventRecord(ev[3], st);
        cudaEventRecord(ev[6], st);
        k_blur<<<dim3(blurTiles, nframes), BLUR_THREADS, 0, st>>>(plan, d_maps->blr);
        ++g_launches;
        cudaEventRecord(ev[7], st);
    }
    if (ev) cudaEventRecord(ev[4], st);
    // describe: one CTA per tile of keypoint positions; needs the kept lists (this stream) and the blurred levels
    if (!ev) {
        e = cudaStreamWaitEvent(st, ss.join, 0);
        if (e != cudaSuccess) return e;
    }
    if (!(skip & 16)) {
        launch_pdl(k_describe_tile, dim3(std::max(1, plan.totalDescTiles), nframes), dim3(DSC_THREADS), kDescSmem, st, plan, d_maps, d_kps, d_desc, cap,
                   d_counts);
        ++g_launches;
    }
    if (ev) cudaEventRecord(ev[5], st);
    return cudaGetLastError();
}
