// Which pipes do the integer / fp16x2 instructions of the detect, blur and describe kernels use on
// sm_100a?  Each mode runs independent chains of one instruction class (or an interleaved mix of two)
// and reports warp-instructions per clock per SM for each class.  If a mix takes max(a, b) instead of
// a + b, the two classes issue to different pipes.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>

#define ITERS 2048
#define CHAINS 8

enum { HMAX2B = 2048, PRMTI = 4096, LDS = 8192, VIMNMX2 = 16384, HADD2 = 32768, VIMNMX3 = 1, HMNMX2 = 2, HFMA2 = 4, PRMT = 8, LOP3 = 16, IMAD = 32, IDP4A = 64, HRELU = 128, POPC = 256, IADD3 = 512, SHF = 1024 };

template <int MODE>
__global__ void k(unsigned* out, unsigned seed, long long* cycles) {
    unsigned a[CHAINS], b[CHAINS];
    __shared__ unsigned sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i * seed;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
        a[i] = (threadIdx.x * 2654435761u + i * 40503u + seed) & 0x00ff00ffu | 0x64006400u;
        b[i] = (threadIdx.x * 40503u + i * 2654435761u + seed) & 0x00ff00ffu | 0x64006400u;
    }
    const unsigned c = (seed * 77u) & 0x00ff00ffu | 0x64006400u;
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (MODE & VIMNMX3) a[i] = __vimax3_u16x2(a[i], a[(i + 1) % CHAINS], c + it);
            if (MODE & HMNMX2) {
                __half2 x = *reinterpret_cast<__half2*>(&b[i]), y = *reinterpret_cast<const __half2*>(&c);
                x = __hmin2(x, y);
                b[i] = *reinterpret_cast<unsigned*>(&x);
            }
            if (MODE & HFMA2) {
                __half2 x = *reinterpret_cast<__half2*>(&b[i]), y = *reinterpret_cast<const __half2*>(&c);
                x = __hfma2(x, y, x);
                b[i] = *reinterpret_cast<unsigned*>(&x);
            }
            if (MODE & HRELU) {
                __half2 x = *reinterpret_cast<__half2*>(&b[i]), y = *reinterpret_cast<const __half2*>(&c);
                x = __hfma2_relu(x, y, x);
                b[i] = *reinterpret_cast<unsigned*>(&x);
            }
            if (MODE & HMAX2B) {  // a second, independent HMNMX2 stream on the a[] registers
                __half2 x = *reinterpret_cast<__half2*>(&a[i]), y = *reinterpret_cast<__half2*>(&a[(i + 1) % CHAINS]);
                x = __hmax2(x, y);
                a[i] = *reinterpret_cast<unsigned*>(&x);
            }
            if (MODE & PRMTI) b[i] = __byte_perm(b[i], c, 0x5432);
            if (MODE & LDS) b[i] ^= sm[(threadIdx.x + i * 32 + (b[i] & 1)) & 1023];
            if (MODE & VIMNMX2) a[i] = __vmaxu2(a[i], a[(i + 1) % CHAINS]);
            if (MODE & HADD2) {
                __half2 x = *reinterpret_cast<__half2*>(&b[i]), y = *reinterpret_cast<const __half2*>(&c);
                x = __hadd2(x, y);
                b[i] = *reinterpret_cast<unsigned*>(&x);
            }
            if (MODE & PRMT) b[i] = __byte_perm(b[i], c, b[(i + 3) % CHAINS]);
            if (MODE & LOP3) b[i] = (b[i] & (c + it)) ^ b[(i + 3) % CHAINS];
            if (MODE & IMAD) b[i] = b[i] * c + b[(i + 3) % CHAINS];
            if (MODE & IDP4A) b[i] = __dp4a(b[i], c, b[i]);
            if (MODE & POPC) b[i] = __popc(b[i]) + 0x5555;
            if (MODE & IADD3) b[i] = b[i] + (c ^ it) + b[(i + 3) % CHAINS];
            if (MODE & SHF) b[i] = __funnelshift_l(b[i], c, it);
        }
    }
    const long long t1 = clock64();
    unsigned r = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) r ^= a[i] ^ b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE>
void run(unsigned* d, long long* dc, const char* name, int nclasses) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k<MODE>, 256, 0);
    k<MODE><<<148 * 8, 256>>>(d, 1, dc);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(d, 2, dc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long cyc;
    cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost);
    // 8 CTAs x 8 warps per SM (64 warps)
    const double per_class = 64.0 * ITERS * CHAINS / (double)cyc;
    printf("%-24s occ %d  %8.1f us  %10lld cycles  %.3f warp-instr/clk/SM per class, %.3f total\n", name, nb, ms * 1e3, cyc, per_class, per_class * nclasses);
}

int main() {
    unsigned* d;
    long long* dc;
    cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaMalloc(&dc, 8);
    run<HMAX2B>(d, dc, "HMNMX2 (a-chain)", 1);
    run<VIMNMX2>(d, dc, "VIMNMX.U16x2 2-input", 1);
    run<PRMTI>(d, dc, "PRMT imm", 1);
    run<LDS>(d, dc, "LDS+LOP", 2);
    run<HADD2>(d, dc, "HADD2", 1);
    run<HMAX2B | HMNMX2>(d, dc, "HMNMX2 + HMNMX2", 2);
    run<HMAX2B | PRMTI>(d, dc, "HMNMX2 + PRMT imm", 2);
    run<HMAX2B | LOP3>(d, dc, "HMNMX2 + LOP3", 2);
    run<HMAX2B | IADD3>(d, dc, "HMNMX2 + IADD3", 2);
    run<HMAX2B | LDS>(d, dc, "HMNMX2 + LDS+LOP", 3);
    run<HMAX2B | HADD2>(d, dc, "HMNMX2 + HADD2", 2);
    run<HMAX2B | IMAD>(d, dc, "HMNMX2 + IMAD", 2);
    run<VIMNMX3 | LDS>(d, dc, "VIMNMX3 + LDS+LOP", 3);
    run<VIMNMX3>(d, dc, "VIMNMX3.U16x2", 1);
    run<HMNMX2>(d, dc, "HMNMX2", 1);
    run<HFMA2>(d, dc, "HFMA2", 1);
    run<HRELU>(d, dc, "HFMA2.RELU", 1);
    run<PRMT>(d, dc, "PRMT", 1);
    run<LOP3>(d, dc, "LOP3", 1);
    run<IMAD>(d, dc, "IMAD", 1);
    run<IDP4A>(d, dc, "IDP.4A", 1);
    run<POPC>(d, dc, "POPC+IADD", 2);
    run<IADD3>(d, dc, "IADD3", 1);
    run<SHF>(d, dc, "SHF", 1);
    run<VIMNMX3 | HMNMX2>(d, dc, "VIMNMX3 + HMNMX2", 2);
    run<VIMNMX3 | HFMA2>(d, dc, "VIMNMX3 + HFMA2", 2);
    run<VIMNMX3 | HRELU>(d, dc, "VIMNMX3 + HFMA2.RELU", 2);
    run<VIMNMX3 | PRMT>(d, dc, "VIMNMX3 + PRMT", 2);
    run<VIMNMX3 | LOP3>(d, dc, "VIMNMX3 + LOP3", 2);
    run<VIMNMX3 | IMAD>(d, dc, "VIMNMX3 + IMAD", 2);
    run<VIMNMX3 | IDP4A>(d, dc, "VIMNMX3 + IDP.4A", 2);
    run<IDP4A | IMAD>(d, dc, "IDP.4A + IMAD", 2);
    run<PRMT | IMAD>(d, dc, "PRMT + IMAD", 2);
    run<LOP3 | IADD3>(d, dc, "LOP3 + IADD3", 2);
    run<SHF | VIMNMX3>(d, dc, "SHF + VIMNMX3", 2);
    run<SHF | IMAD>(d, dc, "SHF + IMAD", 2);
    return 0;
}
