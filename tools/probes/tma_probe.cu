// Minimal TMA probe: which descriptor / coordinate variants load correctly on this GPU?
// nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu && ./tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void k_probe(const CUtensorMap* map, int x, int y, int z, int boxW, int boxH, uint8_t* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned b = smem_u32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(boxW * boxH) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smem_u32(smem)), "l"((unsigned long long)map), "r"(x), "r"(y), "r"(z), "r"(b) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(smem)), "l"((unsigned long long)map), "r"(x), "r"(y), "r"(b) : "memory");
    }
    asm volatile(
        "{\n.reg .pred p;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra LAB_DONE;\nbra LAB_WAIT;\nLAB_DONE:\n}\n" ::"r"(b), "r"(0) : "memory");
    __syncthreads();
    for (int i = threadIdx.x; i < boxW * boxH; i += blockDim.x) out[i] = smem[i];
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int run(EncodeFn enc, int rank, int cols, int rows, int frames, int pitch, int boxW, int boxH, int x, int y, int z, const uint8_t* d_img,
        const std::vector<uint8_t>& h_img) {
    CUtensorMap m;
    cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)(rank == 3 ? rows : rows * frames), (cuuint64_t)frames};
    cuuint64_t gstr[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * rows};
    cuuint32_t box[3] = {(cuuint32_t)boxW, (cuuint32_t)boxH, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, (void*)d_img, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("rank %d box %dx%d at (%d,%d,%d): encode=%d ", rank, boxW, boxH, x, y, z, (int)r);
    if (r != CUDA_SUCCESS) { printf("\n"); return 1; }
    CUtensorMap* d_m;
    cudaMalloc(&d_m, sizeof m);
    cudaMemcpy(d_m, &m, sizeof m, cudaMemcpyHostToDevice);
    uint8_t* d_out;
    cudaMalloc(&d_out, boxW * boxH);
    cudaMemset(d_out, 0xEE, boxW * boxH);
    if (rank == 3) k_probe<3><<<1, 128, boxW * boxH>>>(d_m, x, y, z, boxW, boxH, d_out);
    else k_probe<2><<<1, 128, boxW * boxH>>>(d_m, x, rank == 2 ? z * rows + y : y, 0, boxW, boxH, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("-> %s\n", cudaGetErrorString(e)); return 2; }
    std::vector<uint8_t> out(boxW * boxH);
    cudaMemcpy(out.data(), d_out, out.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r2 = 0; r2 < boxH; ++r2)
        for (int c = 0; c < boxW; ++c) {
            int xx = x + c, yy = y + r2;
            uint8_t want = (xx < cols && yy < rows && xx >= 0 && yy >= 0) ? h_img[((size_t)z * rows + yy) * pitch + xx] : 0;
            if (rank == 2 && yy >= rows && z * rows + yy < rows * frames && xx < cols) want = h_img[((size_t)z * rows + yy) * pitch + xx];
            bad += out[r2 * boxW + c] != want;
        }
    printf("-> ok, %d mismatching bytes\n", bad);
    cudaFree(d_m);
    cudaFree(d_out);
    return 0;
}

int main() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no entry point\n"); return 1; }
    EncodeFn enc = (EncodeFn)p;
    const int cols = 640, rows = 480, frames = 3, pitch = 640;
    std::vector<uint8_t> h((size_t)pitch * rows * frames);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t* d;
    cudaMalloc(&d, h.size());
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    int variants[][6] = {
        // rank, boxW, boxH, x, y, z
        {2, 128, 16, 0, 0, 0},   {2, 128, 16, 16, 5, 1},  {2, 128, 16, 17, 5, 1},  {2, 256, 38, 16, 16, 0}, {2, 256, 38, 47, 16, 2},
        {3, 128, 16, 0, 0, 0},   {3, 128, 16, 16, 5, 1},  {3, 128, 16, 17, 5, 1},  {3, 256, 38, 16, 16, 0}, {3, 256, 38, 47, 16, 2},
        {3, 256, 65, 499, 440, 1}, {2, 256, 65, 499, 440, 1},
    };
    for (auto& v : variants) {
        int rc = run(enc, v[0], cols, rows, frames, pitch, v[1], v[2], v[3], v[4], v[5], d, h);
        if (rc == 2) {  // sticky error: the context is gone
            printf("context lost, stopping\n");
            break;
        }
    }
    return 0;
}
