// microbench.cu — FP32 FMA-pipe peak, measured on the box the bench runs on (roofline denominator
// for the FP32-bound noise stage; MEASURED_PEAKS.json only carries HBM and bf16 tensor peaks).
#include "nz_common.cuh"

namespace nz {
namespace {

constexpr int CHAINS = 16;

__global__ void __launch_bounds__(256) fma_peak_kernel(float* __restrict__ sink, int iters, float a, float b) {
    float v[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) v[c] = (float)(threadIdx.x + c) * 1e-3f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) v[c] = fmaf(v[c], a, b);
    }
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s += v[c];
    if (s == 123.456f) sink[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true; keeps the chains live
}

}  // namespace

int32_t launch_fma_peak(float* d_sink, int grid, int iters, double* flops, cudaStream_t s) {
    NZ_REQUIRE(d_sink && grid > 0 && iters > 0, "fma_peak: bad arguments");
    fma_peak_kernel<<<grid, 256, 0, s>>>(d_sink, iters, 0.999f, 0.001f);
    NZ_LAUNCHED();
    if (flops) *flops = 2.0 * CHAINS * (double)iters * 256.0 * (double)grid;
    return NZ_OK;
}

}  // namespace nz
