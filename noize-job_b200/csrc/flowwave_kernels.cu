// flowwave_kernels.cu — the whole flow map (fill, I x (outflow, water), velocity, normalise) in ONE launch.
//
// Same arithmetic, cell for cell, as flow_kernels.cu / the flow map of oracle/noize_oracle.cpp (FlowMapComponents.cs:20-165,
// FlowMapStage.cs:124-195).  What changes is where the state lives.
//
// Per-iteration kernels stream water + 4 flow fields through HBM every iteration (~44 B/cell/iteration,
// 61 GB at 16384^2 x 5).  But water and flows are DERIVED from the height field, so with all iterations
// fused only 4 B/cell are read and 4 B/cell written.  A square smem tile cannot hold it (24 B/cell of
// state, 2I+.. halo on four sides), so the kernel streams instead (wavefront / time-skewed blocking):
//
//   a CTA owns a strip of FLW = 256 columns (2I halo columns each side) and walks DOWN a chunk of rows.
//   At step s it works on one row per pipeline stage, each stage lagging the previous one by 2 rows so that
//   everything a stage reads was produced in an EARLIER step (one __syncthreads per step):
//       L    : load height row s                                   (global -> smem ring of 4I rows)
//       A_t  : outflow step of level t on row s-(4t-2)   reads H_{t-1} = w_{t-1}+h (3 rows), w_{t-1}, f_{t-1}
//       B_t  : water   step of level t on row s-4t       reads f_t (3 rows), w_{t-1};  writes w_t and H_t
//       V    : velocity + normalise      on row s-4I     reads f_I (3 rows);           writes the result row
//   Each level keeps a ring of 5 rows per field (a row is last read 4 steps after it was written).
//   State per CTA: (20 I + 10 (I-1) + 4 I) rows x 1 KB = 160 KB for I = 5: one 512-thread CTA per SM.
//
// Every warp takes (stage, 64-column chunk) work items round-robin; a lane owns 2 adjacent columns, so
// rows are read with LDS.64 and only the two outer neighbours are scalar loads.
//
// Clamp-to-edge (TileData.cs:72-77) is applied where the reference applies it: neighbour column/row
// indices are clamped to the GRID, so a border cell reads its own current-level value.  At strip/chunk
// edges that are not grid edges the clamp yields garbage that advances one cell per stage and stays inside
// the 2I-wide halo.
#include "nz_common.cuh"

namespace nz {
namespace {

constexpr int FLW = 256;          // strip width in floats, halo included
constexpr int FL_THREADS = 512;
constexpr int FL_WARPS = FL_THREADS / 32;
constexpr int RING = 5;
constexpr int XCH = FLW / 64;     // 64-column chunks per row
constexpr int FLOW_WAVE_MAX_I = 5;
constexpr float TIMESTEP = 0.2f;
constexpr float WATER0 = 0.0001f;  // FillArrayJob value, FlowMapStage.cs:129

struct WaveParams {
    const float* h;
    float* out;
    int W, H, I;
    int zc;        // rows per chunk
    int swi;       // interior columns per strip = FLW - 2*HX
    int hx;        // halo columns each side (2I rounded up to even)
    float nmin, nrange;
};

__device__ __forceinline__ float2 ld2(const float* row, int c) { return *reinterpret_cast<const float2*>(row + c); }
__device__ __forceinline__ void st2(float* row, int c, float a, float b) { *reinterpret_cast<float2*>(row + c) = make_float2(a, b); }

// ComputeFlowStep.CalculateCell for one cell (W,E,S,N order; S = z-1, N = z+1)
__device__ __forceinline__ void flow_cell(float H0, float HW, float HE, float HS, float HN, float w0, float fW, float fE,
                                          float fS, float fN, float& oW, float& oE, float& oS, float& oN) {
    const float flW = fmaxf(0.0f, fW + (H0 - HW));
    const float flE = fmaxf(0.0f, fE + (H0 - HE));
    const float flS = fmaxf(0.0f, fS + (H0 - HS));
    const float flN = fmaxf(0.0f, fN + (H0 - HN));
    const float sum_ = (flW + flE) + (flS + flN);
    float K = 0.0f;
    if (sum_ > 0.0f) {
        K = w0 / (sum_ * TIMESTEP);
        K = fminf(fmaxf(K, 0.0f), 1.0f);
    }
    const bool pos = sum_ > 0.0f;
    oW = pos ? flW * K : 0.0f;
    oE = pos ? flE * K : 0.0f;
    oS = pos ? flS * K : 0.0f;
    oN = pos ? flN * K : 0.0f;
}

template <int I>
__global__ void __launch_bounds__(FL_THREADS, 1) flow_wave_kernel(WaveParams p) {
    extern __shared__ __align__(16) float sm[];
    const int W = p.W, H = p.H;
    // smem layout (rows of FLW floats)
    float* fbase = sm;                                   // [I][4][RING]   level t=1..I, field W,E,S,N
    float* wbase = fbase + I * 4 * RING * FLW;           // [I-1][RING]    w_t, t=1..I-1
    float* Hbase = wbase + (I - 1) * RING * FLW;         // [I-1][RING]    H_t = w_t + h
    float* hbase = Hbase + (I - 1) * RING * FLW;         // [4I]           height rows
    constexpr int RH = 4 * I;
#define F_ROW(t, k, z) (fbase + ((((t) - 1) * 4 + (k)) * RING + (z) % RING) * FLW)
#define W_ROW(t, z) (wbase + (((t) - 1) * RING + (z) % RING) * FLW)
#define H_ROW(t, z) (Hbase + (((t) - 1) * RING + (z) % RING) * FLW)
#define HT_ROW(z) (hbase + ((z) % RH) * FLW)

    const int xs0 = blockIdx.x * p.swi - p.hx;           // grid x of strip column 0 (even)
    const int zc0 = blockIdx.y * p.zc, zc1 = min(zc0 + p.zc, H);
    const int cmin = max(0, -xs0), cmax = min(FLW - 1, W - 1 - xs0);   // strip columns inside the grid
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nitems = (2 * I + 1) * XCH;
    const int hlo = max(0, zc0 - 2 * I), hhi = min(H, zc1 + 2 * I);

    for (int s = zc0 - 2 * I; s < zc1 + 4 * I; s++) {
        for (int item = warp; item < nitems; item += FL_WARPS) {
            const int stage = item / XCH;                // 0 = L, 2t-1 = A_t, 2t = B_t (t < I), 2I = V
            const int c = (item % XCH) * 64 + lane * 2;  // this lane's two columns: c, c+1
            const int cl = max(c - 1, cmin), cr = min(c + 2, cmax);
            if (stage == 0) {
                // ---- L: height row s ---------------------------------------------------------------
                if (s >= hlo && s < hhi) {
                    const float* g = p.h + (size_t)s * W;
                    const int gx = xs0 + c;
                    float a, b;
                    if (gx >= 0 && gx + 1 < W) {
                        const float2 v = __ldg(reinterpret_cast<const float2*>(g + gx));
                        a = v.x; b = v.y;
                    } else {
                        a = __ldg(g + min(max(gx, 0), W - 1));
                        b = __ldg(g + min(max(gx + 1, 0), W - 1));
                    }
                    st2(HT_ROW(s), c, a, b);
                }
            } else if (stage == 2 * I) {
                // ---- V: velocity magnitude + normalise, row s - 4I -----------------------------------
                const int r = s - 4 * I;
                if (r >= zc0 && r < zc1) {
                    const int rs = max(r - 1, 0), rn = min(r + 1, H - 1);
                    const float* fWr = F_ROW(I, 0, r); const float* fEr = F_ROW(I, 1, r);
                    const float2 fW = ld2(fWr, c), fE = ld2(fEr, c);
                    const float fE_l = fEr[cl], fW_r = fWr[cr];
                    const float2 fS = ld2(F_ROW(I, 2, r), c), fN = ld2(F_ROW(I, 3, r), c);
                    const float2 fS_n = ld2(F_ROW(I, 2, rn), c), fN_s = ld2(F_ROW(I, 3, rs), c);
                    float res[2];
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        const float dl = (q ? fE.x : fE_l) - (q ? fW.y : fW.x);
                        const float dr = (q ? fE.y : fE.x) - (q ? fW_r : fW.y);
                        const float dt = (q ? fS_n.y : fS_n.x) - (q ? fN.y : fN.x);
                        const float db = (q ? fS.y : fS.x) - (q ? fN_s.y : fN_s.x);
                        const float vx = (dl + dr) * 0.5f, vy = (dt + db) * 0.5f;
                        float v = sqrtf(fmaf(vy, vy, vx * vx));
                        if (p.nrange < 1e-12f) v = 0.0f;
                        res[q] = (v - p.nmin) / p.nrange;
                    }
                    const int gx = xs0 + c;
                    if (c >= p.hx && c < FLW - p.hx && gx + 1 < W)
                        *reinterpret_cast<float2*>(p.out + (size_t)r * W + gx) = make_float2(res[0], res[1]);
                }
            } else if (stage & 1) {
                // ---- A_t: outflow step, row s - (4t-2) -----------------------------------------------
                const int t = (stage + 1) >> 1;
                const int r = s - (4 * t - 2);
                const int lo = max(0, zc0 - (2 * I - 2 * t + 1)), hi = min(H, zc1 + (2 * I - 2 * t + 1));
                if (r >= lo && r < hi) {
                    const int rs = max(r - 1, 0), rn = min(r + 1, H - 1);
                    float2 H0, HS, HN, w0, fW, fE, fS, fN;
                    float HWl, HEr;
                    if (t == 1) {
                        // level 0: water == 1e-4 everywhere, flows == 0: H_0 = 1e-4 + h computed on the fly
                        const float* hr = HT_ROW(r);
                        const float2 a = ld2(hr, c), b = ld2(HT_ROW(rs), c), d = ld2(HT_ROW(rn), c);
                        H0 = make_float2(WATER0 + a.x, WATER0 + a.y);
                        HS = make_float2(WATER0 + b.x, WATER0 + b.y);
                        HN = make_float2(WATER0 + d.x, WATER0 + d.y);
                        HWl = WATER0 + hr[cl];
                        HEr = WATER0 + hr[cr];
                        w0 = make_float2(WATER0, WATER0);
                        fW = fE = fS = fN = make_float2(0.0f, 0.0f);
                    } else {
                        const float* Hr = H_ROW(t - 1, r);
                        H0 = ld2(Hr, c); HS = ld2(H_ROW(t - 1, rs), c); HN = ld2(H_ROW(t - 1, rn), c);
                        HWl = Hr[cl]; HEr = Hr[cr];
                        w0 = ld2(W_ROW(t - 1, r), c);
                        fW = ld2(F_ROW(t - 1, 0, r), c); fE = ld2(F_ROW(t - 1, 1, r), c);
                        fS = ld2(F_ROW(t - 1, 2, r), c); fN = ld2(F_ROW(t - 1, 3, r), c);
                    }
                    float oW0, oE0, oS0, oN0, oW1, oE1, oS1, oN1;
                    flow_cell(H0.x, HWl, H0.y, HS.x, HN.x, w0.x, fW.x, fE.x, fS.x, fN.x, oW0, oE0, oS0, oN0);
                    flow_cell(H0.y, H0.x, HEr, HS.y, HN.y, w0.y, fW.y, fE.y, fS.y, fN.y, oW1, oE1, oS1, oN1);
                    st2(F_ROW(t, 0, r), c, oW0, oW1);
                    st2(F_ROW(t, 1, r), c, oE0, oE1);
                    st2(F_ROW(t, 2, r), c, oS0, oS1);
                    st2(F_ROW(t, 3, r), c, oN0, oN1);
                }
            } else {
                // ---- B_t: water step, row s - 4t (t < I) ---------------------------------------------
                const int t = stage >> 1;
                const int r = s - 4 * t;
                const int lo = max(0, zc0 - (2 * I - 2 * t)), hi = min(H, zc1 + (2 * I - 2 * t));
                if (r >= lo && r < hi) {
                    const int rs = max(r - 1, 0), rn = min(r + 1, H - 1);
                    const float* fWr = F_ROW(t, 0, r); const float* fEr = F_ROW(t, 1, r);
                    const float2 fW = ld2(fWr, c), fE = ld2(fEr, c);
                    const float2 fS = ld2(F_ROW(t, 2, r), c), fN = ld2(F_ROW(t, 3, r), c);
                    const float fE_l = fEr[cl], fW_r = fWr[cr];
                    const float2 fN_s = ld2(F_ROW(t, 3, rs), c), fS_n = ld2(F_ROW(t, 2, rn), c);
                    const float2 w = (t == 1) ? make_float2(WATER0, WATER0) : ld2(W_ROW(t - 1, r), c);
                    const float2 hh = ld2(HT_ROW(r), c);
                    const float out0 = ((fW.x + fE.x) + fS.x) + fN.x;
                    const float out1 = ((fW.y + fE.y) + fS.y) + fN.y;
                    const float in0 = ((fE_l + fW.y) + fN_s.x) + fS_n.x;
                    const float in1 = ((fE.x + fW_r) + fN_s.y) + fS_n.y;
                    const float nw0 = fmaxf(0.0f, fmaf(in0 - out0, TIMESTEP, w.x));
                    const float nw1 = fmaxf(0.0f, fmaf(in1 - out1, TIMESTEP, w.y));
                    st2(W_ROW(t, r), c, nw0, nw1);
                    st2(H_ROW(t, r), c, nw0 + hh.x, nw1 + hh.y);
                }
            }
        }
        __syncthreads();
    }
#undef F_ROW
#undef W_ROW
#undef H_ROW
#undef HT_ROW
}

size_t wave_smem_bytes(int I) { return (size_t)(20 * I + 10 * (I - 1) + 4 * I) * FLW * sizeof(float); }

}  // namespace

bool flow_wave_supported(int width, int rows, int iterations, const void* a, const void* b) {
    return iterations >= 1 && iterations <= FLOW_WAVE_MAX_I && (width & 1) == 0 && rows >= 1 &&
           (((uintptr_t)a | (uintptr_t)b) & 7) == 0;
}

// d_out must not alias d_height
int32_t launch_flow_wave(const float* d_height, float* d_out, int width, int rows, int iterations, float norm_min,
                         float norm_max, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        NZ_CUDA(cudaFuncSetAttribute(flow_wave_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wave_smem_bytes(1)));
        NZ_CUDA(cudaFuncSetAttribute(flow_wave_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wave_smem_bytes(2)));
        NZ_CUDA(cudaFuncSetAttribute(flow_wave_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wave_smem_bytes(3)));
        NZ_CUDA(cudaFuncSetAttribute(flow_wave_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wave_smem_bytes(4)));
        NZ_CUDA(cudaFuncSetAttribute(flow_wave_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wave_smem_bytes(5)));
        attr_set = true;
    }
    const int I = iterations;
    WaveParams p;
    p.h = d_height; p.out = d_out; p.W = width; p.H = rows; p.I = I;
    p.hx = (2 * I + 1) & ~1;
    p.swi = FLW - 2 * p.hx;
    p.nmin = norm_min;
    p.nrange = norm_max - norm_min;
    const int strips = cdiv(width, p.swi);
    // rows per chunk: long enough to amortise the 6I-row pipeline fill, and a CTA count that fills whole waves
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int best_nz = 1;
    double best_cost = 1e300;
    for (int nz = 1; nz <= rows && nz <= 4096; nz++) {
        const int zc = cdiv(rows, nz);
        if (zc < 64 && nz > 1) break;
        const long long ctas = (long long)strips * cdiv(rows, zc);
        const long long waves = (ctas + sms - 1) / sms;
        const double cost = (double)waves * (zc + 6 * I);   // steps executed by the busiest SM
        if (cost < best_cost) { best_cost = cost; best_nz = nz; }
    }
    p.zc = cdiv(rows, best_nz);
    dim3 grid(strips, cdiv(rows, p.zc));
    switch (I) {
        case 1: flow_wave_kernel<1><<<grid, FL_THREADS, wave_smem_bytes(1), s>>>(p); break;
        case 2: flow_wave_kernel<2><<<grid, FL_THREADS, wave_smem_bytes(2), s>>>(p); break;
        case 3: flow_wave_kernel<3><<<grid, FL_THREADS, wave_smem_bytes(3), s>>>(p); break;
        case 4: flow_wave_kernel<4><<<grid, FL_THREADS, wave_smem_bytes(4), s>>>(p); break;
        default: flow_wave_kernel<5><<<grid, FL_THREADS, wave_smem_bytes(5), s>>>(p); break;
    }
    NZ_LAUNCHED();
    return NZ_OK;
}

}  // namespace nz
