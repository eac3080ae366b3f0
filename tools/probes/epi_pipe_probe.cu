// Issue rates of the instructions the matcher's epilogue is made of, per SM sub-partition (SMSP), on B200:
// W warps per SMSP run ILP-8 chains of one opcode (or of two opcodes interleaved) and report warp instructions per clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o epi_pipe_probe epi_pipe_probe.cu && ./epi_pipe_probe
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

enum Op { VMIN2, VMIN3, HMIN2, IMAD, FMNMX, LOP3, PRMT, MIX_V_H, MIX_V_I, MIX_V3_I, MIX_H_I, VADDMIN, MIX_VA_I, N_OPS };
static const char* NAMES[] = {"VIMNMX.U16x2", "VIMNMX3.U16x2", "HMNMX2", "IMAD", "FMNMX", "LOP3", "PRMT",
                              "VIMNMX+HMNMX2", "VIMNMX+IMAD", "VIMNMX3+IMAD", "HMNMX2+IMAD", "VIADDMNMX", "VIADDMNMX+IMAD"};

__device__ __forceinline__ unsigned hmin2u(unsigned a, unsigned b) {
    unsigned d;
    asm volatile("min.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ unsigned hmax2u(unsigned a, unsigned b) {
    unsigned d;
    asm volatile("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ unsigned vmin2(unsigned a, unsigned b) {
    unsigned d;
    asm volatile("vmin2.u32.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(0u));
    return d;
}

template <int OP>
__global__ void k(int iters, unsigned seed, unsigned mul, unsigned* out, long long* cycles) {
    unsigned x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = seed * (threadIdx.x + 1) + i * 0x01010101u;
    unsigned y = seed ^ 0x12345678u;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            unsigned z[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const unsigned a = x[i], b = x[(i + 3) & 7];  // partner from another chain: nothing folds, ILP stays 8
                unsigned d = 0;
                if (OP == VMIN2) d = (r & 1) ? __vminu2(a, b) : __vmaxu2(a, b);
                if (OP == VMIN3) d = (r & 1) ? __vimin3_u16x2(a, b, y) : __vimax3_u16x2(a, b, y);
                if (OP == HMIN2) d = (r & 1) ? hmin2u(a, b) : hmax2u(a, b);
                if (OP == IMAD) d = a * mul + b;
                if (OP == FMNMX) d = __float_as_uint((r & 1) ? fminf(__uint_as_float(a), __uint_as_float(b)) : fmaxf(__uint_as_float(a), __uint_as_float(b)));
                if (OP == LOP3) d = (a & b) ^ mul;
                if (OP == PRMT) d = __byte_perm(a, b, mul);
                if (OP == MIX_V_H) d = (i & 1) ? ((r & 1) ? __vminu2(a, b) : __vmaxu2(a, b)) : ((r & 1) ? hmin2u(a, b) : hmax2u(a, b));
                if (OP == MIX_V_I) d = (i & 1) ? ((r & 1) ? __vminu2(a, b) : __vmaxu2(a, b)) : a * mul + b;
                if (OP == MIX_V3_I) d = (i & 1) ? ((r & 1) ? __vimin3_u16x2(a, b, y) : __vimax3_u16x2(a, b, y)) : a * mul + b;
                if (OP == MIX_H_I) d = (i & 1) ? ((r & 1) ? hmin2u(a, b) : hmax2u(a, b)) : a * mul + b;
                if (OP == VADDMIN) d = __viaddmin_u16x2(a, b, y);
                if (OP == MIX_VA_I) d = (i & 1) ? __viaddmin_u16x2(a, b, y) : a * mul + b;
                z[i] = d;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = z[i];
        }
    }
    const long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(int warpsPerSmsp, unsigned* out, long long* cyc) {
    const int iters = 2000, threads = warpsPerSmsp * 4 * 32, blocks = 148;
    k<OP><<<blocks, threads>>>(iters, 3u, 5u, out, cyc);
    k<OP><<<blocks, threads>>>(iters, 3u, 5u, out, cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += (double)h[i] / 148;
    const double inst = (double)iters * 32 * warpsPerSmsp;  // warp instructions per SMSP
    printf("%-16s warps/SMSP %d: %.3f warp-instr / clk / SMSP\n", NAMES[OP], warpsPerSmsp, inst / avg);
}

int main() {
    unsigned* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 148 * 8);
    for (int w : {1, 2, 4}) {
        run<VMIN2>(w, out, cyc);
        run<VMIN3>(w, out, cyc);
        run<HMIN2>(w, out, cyc);
        run<IMAD>(w, out, cyc);
        run<FMNMX>(w, out, cyc);
        run<LOP3>(w, out, cyc);
        run<PRMT>(w, out, cyc);
        run<MIX_V_H>(w, out, cyc);
        run<MIX_V_I>(w, out, cyc);
        run<MIX_V3_I>(w, out, cyc);
        run<MIX_H_I>(w, out, cyc);
        run<VADDMIN>(w, out, cyc);
        run<MIX_VA_I>(w, out, cyc);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
