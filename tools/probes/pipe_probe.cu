// Which pipe runs fp16x2 min/max on sm_100a?  Throughput of VIMNMX3.U16x2 alone, HMNMX2 alone,
// HFMA2 alone, and interleaved mixes.  If a mix takes max(a, b) instead of a + b, the two
// instruction classes issue to different pipes.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>

#define ITERS 4096
#define CHAINS 8

template <int MODE>
__global__ void k(unsigned* out, unsigned seed) {
    unsigned a[CHAINS], b[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
        a[i] = (threadIdx.x * 2654435761u + i * 40503u + seed) & 0x00ff00ffu | 0x64006400u;
        b[i] = (threadIdx.x * 40503u + i * 2654435761u + seed) & 0x00ff00ffu | 0x64006400u;
    }
    const unsigned c = (seed * 77u) & 0x00ff00ffu | 0x64006400u;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (MODE == 0 || MODE == 3 || MODE == 4) a[i] = __vimax3_u16x2(a[i], b[i], c + it);  // VIMNMX3
            if (MODE == 1 || MODE == 3) {  // HMNMX2
                __half2 x = *reinterpret_cast<__half2*>(&b[i]), y = *reinterpret_cast<__half2*>(&a[(i + 1) % CHAINS]);
                x = __hmin2(x, y);
                b[i] = *reinterpret_cast<unsigned*>(&x) + 1;
            }
            if (MODE == 2 || MODE == 4) {  // HFMA2
                __half2 x = *reinterpret_cast<__half2*>(&b[i]), y = *reinterpret_cast<const __half2*>(&c);
                x = __hfma2(x, y, x);
                b[i] = *reinterpret_cast<unsigned*>(&x);
            }
            if (MODE == 5) a[i] = __vimax3_u16x2(a[i], b[i], c + it), b[i] = __vimin3_u16x2(b[i], a[i], c);  // 2x VIMNMX3
        }
    }
    unsigned r = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) r ^= a[i] ^ b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
float run(unsigned* d) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(d, 1);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(d, 2);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    unsigned* d;
    cudaMalloc(&d, 148 * 8 * 256 * 4);
    const char* names[] = {"VIMNMX3 only", "HMNMX2 only", "HFMA2 only", "VIMNMX3 + HMNMX2", "VIMNMX3 + HFMA2", "2x VIMNMX3"};
    float t[6] = {run<0>(d), run<1>(d), run<2>(d), run<3>(d), run<4>(d), run<5>(d)};
    const double ops = 148.0 * 8 * 256 / 32 * ITERS * CHAINS;  // warp-instructions of each class
    for (int i = 0; i < 6; ++i) printf("%-20s %8.3f ms  (%.2f warp-instr/clk/SM per class at 1.9 GHz)\n", names[i], t[i], ops / (t[i] * 1e-3) / 148 / 1.9e9);
    return 0;
}
