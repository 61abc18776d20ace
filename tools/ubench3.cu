// ubench3.cu — FFMA2/FMUL2/FADD2 rate as a function of operand form (full register pairs vs scalar broadcast vs
// immediates), and whether ALU instructions issue underneath them.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench3 tools/ubench3.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CHAINS 8
template <int MODE>
__global__ void __launch_bounds__(256) k(float* sink, int iters, float a, float b) {
    float2 v[CHAINS], w[CHAINS], u[CHAINS];
    int z[CHAINS];
    const int bi = __float_as_int(b);
    for (int c = 0; c < CHAINS; c++) {
        v[c] = make_float2(threadIdx.x * 0.37f + c, threadIdx.x * 0.11f - c);
        w[c] = make_float2(a + c * 1e-6f, a - c * 1e-6f + threadIdx.x * 1e-9f);
        u[c] = make_float2(b + c * 1e-6f, b - c * 1e-6f + threadIdx.x * 1e-9f);
        z[c] = threadIdx.x + c;
    }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
            if (MODE == 0) v[c] = __ffma2_rn(v[c], w[0], u[0]);
            if (MODE == 1) v[c] = __ffma2_rn(v[c], w[c], u[c]);
            if (MODE == 2) v[c] = __ffma2_rn(v[c], make_float2(a, a), u[c]);
            if (MODE == 3) v[c] = __fmul2_rn(v[c], w[c]);
            if (MODE == 4) v[c] = __fadd2_rn(v[c], u[c]);
            if (MODE == 5) v[c] = __ffma2_rn(v[c], v[c], u[c]);
            if (MODE == 6) v[c] = __fmul2_rn(v[c], v[c]);
            if (MODE == 7) v[c] = __fmul2_rn(v[c], make_float2(0.999f, 0.999f));
            if (MODE == 8) v[c] = __fmul2_rn(v[c], make_float2(a, a));
            if (MODE == 9) v[c] = __ffma2_rn(v[c], w[c], make_float2(0.001f, 0.001f));
            if (MODE == 10) { v[c].x = fmaf(v[c].x, w[c].x, u[c].x); v[c].y = fmaf(v[c].y, w[c].y, u[c].y); }   // 2 scalar FFMA
            if (MODE == 11 || MODE == 12 || MODE == 13) {
                if (MODE == 11) v[c] = __ffma2_rn(v[c], w[c], u[c]);
                if (MODE == 12) v[c] = __ffma2_rn(v[c], make_float2(a, a), u[c]);
                if (MODE == 13) v[c] = __fmul2_rn(v[c], w[c]);
                asm volatile("xor.b32 %0, %0, %1;\n\tadd.s32 %0, %0, %1;" : "+r"(z[c]) : "r"(bi));
            }
            if (MODE == 14) { v[c] = __ffma2_rn(v[c], make_float2(a, a), u[c]); w[c].x = fmaf(w[c].x, a, b); }  // ffma2 + scalar ffma
        }
    }
    float s = 0;
    for (int c = 0; c < CHAINS; c++) s += v[c].x + v[c].y + w[c].x + u[c].y + (float)z[c];
    if (s == 123.456f) sink[0] = s;
}
template <typename K>
void run(const char* name, K kern, float* sink, int insts) {
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int iters = 4096, grid = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<grid, 256>>>(sink, 64, 0.999f, 0.001f);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        kern<<<grid, 256>>>(sink, iters, 0.999f, 0.001f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double groups = (double)grid * 8 * iters * CHAINS;     // warp-level groups of `insts` instructions
    double clk_per_group = (best * 1e-3) * (khz * 1e3) * sms * 4 / groups;
    printf("%-40s %8.3f ms  %5.2f clk per group of %d inst per SMSP\n", name, best, clk_per_group, insts);
}
int main() {
    float* sink; cudaMalloc(&sink, 1024);
    run("ffma2 v,W,U (shared pairs, reuse)", k<0>, sink, 1);
    run("ffma2 v,w_c,u_c (3 pair regs)", k<1>, sink, 1);
    run("ffma2 v,s.F32,u_c", k<2>, sink, 1);
    run("fmul2 v,w_c", k<3>, sink, 1);
    run("fadd2 v,u_c", k<4>, sink, 1);
    run("ffma2 v,v,u_c", k<5>, sink, 1);
    run("fmul2 v,v", k<6>, sink, 1);
    run("fmul2 v,imm", k<7>, sink, 1);
    run("fmul2 v,s.F32", k<8>, sink, 1);
    run("ffma2 v,w_c,imm", k<9>, sink, 1);
    run("2x scalar ffma (3 regs)", k<10>, sink, 2);
    run("ffma2 3-pair + 2 alu", k<11>, sink, 3);
    run("ffma2 v,s,u + 2 alu", k<12>, sink, 3);
    run("fmul2 v,w + 2 alu", k<13>, sink, 3);
    run("ffma2 v,s,u + scalar ffma", k<14>, sink, 2);
    return 0;
}
