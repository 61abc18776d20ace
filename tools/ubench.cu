// ubench.cu — instruction-throughput probe for the ops the noise kernel is made of (run on the B200).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu && tools/ubench
// Each kernel runs 8 independent dependent-chains of ONE instruction per thread; the table printed is
// warp-instructions / clock / SM (4.0 = one per SMSP per clock).
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
#define DEF_KERNEL(NAME, ASM)                                                         \
    __global__ void __launch_bounds__(256) k_##NAME(float* sink, int iters, float a, float b) { \
        float v[CHAINS];                                                              \
        for (int c = 0; c < CHAINS; c++) v[c] = (float)(threadIdx.x * 8 + c) * 0.37f + a; \
        for (int i = 0; i < iters; i++) {                                             \
            _Pragma("unroll") for (int c = 0; c < CHAINS; c++) { ASM; }               \
        }                                                                             \
        float s = 0;                                                                  \
        for (int c = 0; c < CHAINS; c++) s += v[c];                                   \
        if (s == 123.456f) sink[0] = s;                                               \
    }

DEF_KERNEL(ffma, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[c]) : "f"(a), "f"(b)))
DEF_KERNEL(ffma_imm, asm volatile("fma.rn.f32 %0, %0, 0f3F7FBE77, 0f3A83126F;" : "+f"(v[c])))
DEF_KERNEL(fmul, asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(v[c]) : "f"(a)))
DEF_KERNEL(fadd, asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(v[c]) : "f"(b)))
DEF_KERNEL(floor, asm volatile("cvt.rmi.f32.f32 %0, %0;" : "+f"(v[c])))
DEF_KERNEL(fmax, asm volatile("max.f32 %0, %0, %1;" : "+f"(v[c]) : "f"(b)))
DEF_KERNEL(fabs_sub, asm volatile("abs.f32 %0, %0;\n\tsub.rn.f32 %0, %0, %1;" : "+f"(v[c]) : "f"(b)))
DEF_KERNEL(selp, asm volatile("{.reg .pred p; setp.gt.f32 p, %0, %1; selp.f32 %0, %2, %0, p;}" : "+f"(v[c]) : "f"(a), "f"(b)))
DEF_KERNEL(f2i_i2f, asm volatile("{.reg .s32 t; cvt.rmi.s32.f32 t, %0; cvt.rn.f32.s32 %0, t;}" : "+f"(v[c])))
DEF_KERNEL(rcp, asm volatile("rcp.approx.f32 %0, %0;" : "+f"(v[c])))
DEF_KERNEL(div, v[c] = a / v[c])
DEF_KERNEL(sinf_, v[c] = sinf(v[c]))
DEF_KERNEL(fmodf_, v[c] = fmodf(v[c] + a, 1010.0f))

__global__ void __launch_bounds__(256) k_imad(int* sink, int iters, int a, int b) {
    int v[CHAINS];
    for (int c = 0; c < CHAINS; c++) v[c] = threadIdx.x * 8 + c + a;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(v[c]) : "r"(a), "r"(b));
    }
    int s = 0;
    for (int c = 0; c < CHAINS; c++) s += v[c];
    if (s == 123456) sink[0] = s;
}
__global__ void __launch_bounds__(256) k_mulhi(int* sink, int iters, int a, int b) {
    unsigned v[CHAINS];
    for (int c = 0; c < CHAINS; c++) v[c] = threadIdx.x * 8 + c + a;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(v[c]) : "r"(a));
    }
    unsigned s = 0;
    for (int c = 0; c < CHAINS; c++) s += v[c];
    if (s == 123456) sink[0] = s;
}
// mixed: alternate FFMA (fma pipe) with FMNMX (alu pipe) to see dual issue
DEF_KERNEL(ffma_fmax, asm volatile("fma.rn.f32 %0, %0, %1, %2;\n\tmax.f32 %0, %0, %2;" : "+f"(v[c]) : "f"(a), "f"(b)))
DEF_KERNEL(ffma_floor, asm volatile("fma.rn.f32 %0, %0, %1, %2;\n\tcvt.rmi.f32.f32 %0, %0;" : "+f"(v[c]) : "f"(a), "f"(b)))

template <typename K, typename T>
void run(const char* name, K kern, T* sink, int ops_per_iter, float a, float b) {
    int dev = 0, sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const int iters = 4096, grid = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<grid, 256>>>(sink, 64, a, b);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        kern<<<grid, 256>>>(sink, iters, a, b);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double winst = (double)grid * 8 /*warps*/ * iters * CHAINS * ops_per_iter;
    double per_s = winst / (best * 1e-3);
    printf("%-12s %8.3f ms  %8.1f Gwarp-inst/s  %6.3f warp-inst/clk/SM @max %d MHz  (%.2f T lane-ops/s)\n", name, best,
           per_s * 1e-9, per_s / sms / (khz * 1e3), khz / 1000, per_s * 32e-12);
}

int main() {
    float* sink; cudaMalloc(&sink, 1024);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("no device: %s\n", cudaGetErrorString(e)); return 1; }
    run("ffma", k_ffma, sink, 1, 0.999f, 0.001f);
    run("ffma_imm", k_ffma_imm, sink, 1, 0.999f, 0.001f);
    run("fmul", k_fmul, sink, 1, 0.9999f, 0.001f);
    run("fadd", k_fadd, sink, 1, 0.999f, 0.001f);
    run("floor", k_floor, sink, 1, 0.999f, 0.001f);
    run("fmax", k_fmax, sink, 1, 0.999f, 0.001f);
    run("fabs_sub", k_fabs_sub, sink, 1, 0.999f, 0.5f);
    run("setp+selp", k_selp, sink, 2, 0.999f, 0.001f);
    run("f2i+i2f", k_f2i_i2f, sink, 2, 0.999f, 0.001f);
    run("rcp", k_rcp, sink, 1, 0.999f, 0.001f);
    run("div(ieee)", k_div, sink, 1, 1.7f, 0.001f);
    run("sinf", k_sinf_, sink, 1, 0.999f, 0.001f);
    run("fmodf", k_fmodf_, sink, 1, 1.5f, 0.001f);
    run("imad", k_imad, (int*)sink, 1, 3, 7);
    run("mul.hi", k_mulhi, (int*)sink, 1, 3, 7);
    run("ffma+fmax", k_ffma_fmax, sink, 2, 0.999f, 0.001f);
    run("ffma+floor", k_ffma_floor, sink, 2, 0.999f, 0.001f);
    return 0;
}
