// ubench2.cu — second probe (run on the B200): packed FP32 (f32x2) issue/pipe rates and shared-memory table
// lookups (conflict-free bank-private LDS.32 vs random LDS.64 / LDS.128), the two levers considered for the
// fBm kernel.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench2 tools/ubench2.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
// ---- packed fp32 ----
__global__ void __launch_bounds__(256) k_ffma2(float* sink, int iters, float a, float b) {
    unsigned long long v[CHAINS], A, B;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(A) : "f"(a));
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(B) : "f"(b));
    for (int c = 0; c < CHAINS; c++) {
        float x = (float)(threadIdx.x * 8 + c) * 0.37f + a;
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(v[c]) : "f"(x));
    }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[c]) : "l"(A), "l"(B));
    }
    float s = 0;
    for (int c = 0; c < CHAINS; c++) { float lo, hi; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[c])); s += lo + hi; }
    if (s == 123.456f) sink[0] = s;
}
__global__ void __launch_bounds__(256) k_fadd2(float* sink, int iters, float a, float b) {
    unsigned long long v[CHAINS], B;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(B) : "f"(b));
    for (int c = 0; c < CHAINS; c++) {
        float x = (float)(threadIdx.x * 8 + c) * 0.37f + a;
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(v[c]) : "f"(x));
    }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v[c]) : "l"(B));
    }
    float s = 0;
    for (int c = 0; c < CHAINS; c++) { float lo, hi; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[c])); s += lo + hi; }
    if (s == 123.456f) sink[0] = s;
}
// ffma2 interleaved with N alu-pipe ops (fmax) on separate chains: does the alu pipe ride in the free issue slots?
template <int NALU>
__global__ void __launch_bounds__(256) k_ffma2_alu(float* sink, int iters, float a, float b) {
    unsigned long long v[CHAINS], A, B;
    int w[CHAINS];
    const int bi = __float_as_int(b);
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(A) : "f"(a));
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(B) : "f"(b));
    for (int c = 0; c < CHAINS; c++) {
        float x = (float)(threadIdx.x * 8 + c) * 0.37f + a;
        w[c] = __float_as_int(x);
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(v[c]) : "f"(x));
    }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[c]) : "l"(A), "l"(B));
#pragma unroll
            for (int k = 0; k < NALU; k++) asm volatile("xor.b32 %0, %0, %1;\n\tadd.s32 %0, %0, %1;" : "+r"(w[c]) : "r"(bi));
        }
    }
    float s = 0;
    for (int c = 0; c < CHAINS; c++) { float lo, hi; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[c])); s += lo + hi + (float)w[c]; }
    if (s == 123.456f) sink[0] = s;
}
// scalar ffma + N fmax for comparison
template <int NALU>
__global__ void __launch_bounds__(256) k_ffma_alu(float* sink, int iters, float a, float b) {
    float v[CHAINS];
    int w[CHAINS];
    const int bi = __float_as_int(b);
    for (int c = 0; c < CHAINS; c++) { v[c] = (float)(threadIdx.x * 8 + c) * 0.37f + a; w[c] = __float_as_int(v[c]); }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[c]) : "f"(a), "f"(b));
#pragma unroll
            for (int k = 0; k < NALU; k++) asm volatile("xor.b32 %0, %0, %1;\n\tadd.s32 %0, %0, %1;" : "+r"(w[c]) : "r"(bi));
        }
    }
    float s = 0;
    for (int c = 0; c < CHAINS; c++) s += v[c] + (float)w[c];
    if (s == 123.456f) sink[0] = s;
}

// ---- shared-memory table lookups: chains of dependent lookups, index = f(previous value) ----
// MODE 0: bank-private LDS.32   tab[entry*32 + lane]            (conflict-free by construction)
// MODE 1: random LDS.32          tab[entry]
// MODE 2: random LDS.64          tab2[entry]
// MODE 3: random LDS.128         tab4[entry]
// MODE 4: same-address (broadcast) LDS.64
template <int MODE>
__global__ void __launch_bounds__(1024) k_lds(float* sink, int iters, float a, float b) {
    extern __shared__ unsigned smem[];
    const int n = (MODE == 0) ? 289 * 32 : (MODE == 1 ? 289 : (MODE == 2 || MODE == 4 ? 289 * 2 : 289 * 4));
    for (int i = threadIdx.x; i < n; i += blockDim.x) smem[i] = (i * 2654435761u) >> 8;
    __syncthreads();
    unsigned v[CHAINS];
    const unsigned lane = threadIdx.x & 31;
    for (int c = 0; c < CHAINS; c++) v[c] = (MODE == 4) ? c * 31 : (threadIdx.x * 7 + c * 13);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
            unsigned e = v[c] % 289u;     // index arithmetic is part of the loop but identical across modes
            if (MODE == 0) v[c] = smem[e * 32 + lane];
            if (MODE == 1) v[c] = smem[e];
            if (MODE == 2 || MODE == 4) { uint2 t = reinterpret_cast<const uint2*>(smem)[e]; v[c] = MODE == 4 ? (t.x ^ t.y) * 0 + c * 31 + (t.x & 0) : (t.x ^ t.y); }
            if (MODE == 3) { uint4 t = reinterpret_cast<const uint4*>(smem)[e]; v[c] = t.x ^ t.y ^ t.z ^ t.w; }
        }
    }
    unsigned s = 0;
    for (int c = 0; c < CHAINS; c++) s += v[c];
    if (s == 123456u) sink[0] = (float)s;
}

template <typename K>
void run(const char* name, K kern, float* sink, double ops_per_iter, size_t smem, int threads = 256, int ctas_per_sm = 8) {
    int dev = 0, sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int iters = 2048, grid = sms * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<grid, threads, smem>>>(sink, 64, 0.999f, 0.001f);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        kern<<<grid, threads, smem>>>(sink, iters, 0.999f, 0.001f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("%-22s FAILED: %s\n", name, cudaGetErrorString(e)); return; }
    double winst = (double)grid * (threads / 32) * iters * CHAINS * ops_per_iter;
    double per_s = winst / (best * 1e-3);
    printf("%-22s %8.3f ms  %8.1f Gwarp-op/s  %6.3f warp-op/clk/SM @max %d MHz\n", name, best, per_s * 1e-9,
           per_s / sms / (khz * 1e3), khz / 1000);
}

int main() {
    float* sink; cudaMalloc(&sink, 1024);
    if (cudaGetLastError() != cudaSuccess) { printf("no device\n"); return 1; }
    printf("# packed fp32: warp-op = one instruction (an f32x2 instruction carries 2 lane-ops per lane)\n");
    run("ffma2", k_ffma2, sink, 1, 0);
    run("fadd2", k_fadd2, sink, 1, 0);
    run("ffma + 2 alu", k_ffma_alu<1>, sink, 3, 0);
    run("ffma2 + 2 alu", k_ffma2_alu<1>, sink, 3, 0);
    run("ffma2 + 4 alu", k_ffma2_alu<2>, sink, 5, 0);
    run("ffma2 + 6 alu", k_ffma2_alu<3>, sink, 7, 0);
    printf("# shared-memory lookups (each op = mod-289 index arithmetic + one LDS)\n");
    run("lds32 bank-private", k_lds<0>, sink, 1, 289 * 32 * 4, 1024, 1);
    run("lds32 random", k_lds<1>, sink, 1, 289 * 32 * 4, 1024, 1);
    run("lds64 random", k_lds<2>, sink, 1, 289 * 32 * 4, 1024, 1);
    run("lds128 random", k_lds<3>, sink, 1, 289 * 32 * 4, 1024, 1);
    run("lds64 broadcast", k_lds<4>, sink, 1, 289 * 32 * 4, 1024, 1);
    return 0;
}
