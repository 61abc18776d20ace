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

// ---- shared-memory rings -------------------------------------------------------------------------------
// Every ring holds RING = 5 rows of FLW floats; ring q starts at float offset q*RING*FLW.
//   F(t,k)  t=1..I, k=W,E,S,N   outflows of level t
//   Wt(t)   t=1..I-1            water of level t
//   Ht(t)   t=1..I-1            water + height of level t (what the outflow step of level t+1 reads)
//   HC(t)   t=0..max(I-2,0)     height rows travelling with the pipeline: HC(0) is written by the loader,
//                               HC(t) by the water stage of level t (which needs h to form Ht(t))
template <int I> struct Rings {
    static constexpr int NHC = I - 1 > 1 ? I - 1 : 1;
    __host__ __device__ static constexpr int F(int t, int k) { return (((t - 1) * 4 + k) * RING) * FLW; }
    __host__ __device__ static constexpr int Wt(int t) { return ((4 * I + (t - 1)) * RING) * FLW; }
    __host__ __device__ static constexpr int Ht(int t) { return ((4 * I + (I - 1) + (t - 1)) * RING) * FLW; }
    __host__ __device__ static constexpr int HC(int t) { return ((4 * I + 2 * (I - 1) + t) * RING) * FLW; }
    static constexpr int ROWS = (4 * I + 2 * (I - 1) + NHC) * RING;
};

struct Lane {
    float* sm;
    int R[RING];   // R[j] = float offset (ring-relative) of the slot of row s-j, plus this lane's first column
    int dl, dr;    // column offsets of the clamped west neighbour of cell c (-1 or 0) and east neighbour of c+1 (+2 or +1)
    int c;         // first of this lane's two strip columns
    int H;         // grid rows
};

// slot of row (s - lag + d), d in {-1,0,+1}; rows outside the grid clamp onto the border row
__device__ __forceinline__ int slot_of(const Lane& L, int lag, int d, int r) {
    const int j0 = lag % RING, jm = (lag + 1) % RING, jp = (lag + RING - 1) % RING;
    if (d == 0) return L.R[j0];
    if (d < 0) return r == 0 ? L.R[j0] : L.R[jm];
    return r == L.H - 1 ? L.R[j0] : L.R[jp];
}

template <int I, int T>
__device__ __forceinline__ void stage_outflow(const Lane& L, int s, int zc0, int zc1) {
    using RG = Rings<I>;
    constexpr int lag = 4 * T - 2;
    const int r = s - lag;
    const int lo = max(0, zc0 - (2 * I - 2 * T + 1)), hi = min(L.H, zc1 + (2 * I - 2 * T + 1));
    if (r < lo || r >= hi) return;
    const int o0 = slot_of(L, lag, 0, r), os = slot_of(L, lag, -1, r), on = slot_of(L, lag, +1, r);
    float* sm = L.sm;
    float2 H0, HS, HN, w0, fW, fE, fS, fN;
    float HWl, HEr;
    if (T == 1) {
        const float2 a = ld2(sm + RG::HC(0), o0), b = ld2(sm + RG::HC(0), os), d = ld2(sm + RG::HC(0), on);
        H0 = make_float2(WATER0 + a.x, WATER0 + a.y);
        HS = make_float2(WATER0 + b.x, WATER0 + b.y);
        HN = make_float2(WATER0 + d.x, WATER0 + d.y);
        HWl = WATER0 + sm[RG::HC(0) + o0 + L.dl];
        HEr = WATER0 + sm[RG::HC(0) + o0 + L.dr];
        w0 = make_float2(WATER0, WATER0);
        fW = fE = fS = fN = make_float2(0.0f, 0.0f);
    } else {
        constexpr int P = T > 1 ? T - 1 : 1;
        H0 = ld2(sm + RG::Ht(P), o0); HS = ld2(sm + RG::Ht(P), os); HN = ld2(sm + RG::Ht(P), on);
        HWl = sm[RG::Ht(P) + o0 + L.dl]; HEr = sm[RG::Ht(P) + o0 + L.dr];
        w0 = ld2(sm + RG::Wt(P), o0);
        fW = ld2(sm + RG::F(P, 0), o0); fE = ld2(sm + RG::F(P, 1), o0);
        fS = ld2(sm + RG::F(P, 2), o0); fN = ld2(sm + RG::F(P, 3), o0);
    }
    float oW0, oE0, oS0, oN0, oW1, oE1, oS1, oN1;
    flow_cell(H0.x, HWl, H0.y, HS.x, HN.x, w0.x, fW.x, fE.x, fS.x, fN.x, oW0, oE0, oS0, oN0);
    flow_cell(H0.y, H0.x, HEr, HS.y, HN.y, w0.y, fW.y, fE.y, fS.y, fN.y, oW1, oE1, oS1, oN1);
    st2(sm + RG::F(T, 0), o0, oW0, oW1);
    st2(sm + RG::F(T, 1), o0, oE0, oE1);
    st2(sm + RG::F(T, 2), o0, oS0, oS1);
    st2(sm + RG::F(T, 3), o0, oN0, oN1);
}

template <int I, int T>
__device__ __forceinline__ void stage_water(const Lane& L, int s, int zc0, int zc1) {
    using RG = Rings<I>;
    constexpr int lag = 4 * T;
    const int r = s - lag;
    const int lo = max(0, zc0 - (2 * I - 2 * T)), hi = min(L.H, zc1 + (2 * I - 2 * T));
    if (r < lo || r >= hi) return;
    const int o0 = slot_of(L, lag, 0, r), os = slot_of(L, lag, -1, r), on = slot_of(L, lag, +1, r);
    float* sm = L.sm;
    const float2 fW = ld2(sm + RG::F(T, 0), o0), fE = ld2(sm + RG::F(T, 1), o0);
    const float2 fS = ld2(sm + RG::F(T, 2), o0), fN = ld2(sm + RG::F(T, 3), o0);
    const float fE_l = sm[RG::F(T, 1) + o0 + L.dl], fW_r = sm[RG::F(T, 0) + o0 + L.dr];
    const float2 fN_s = ld2(sm + RG::F(T, 3), os), fS_n = ld2(sm + RG::F(T, 2), on);
    constexpr int P = T > 1 ? T - 1 : 1;
    const float2 w = (T == 1) ? make_float2(WATER0, WATER0) : ld2(sm + RG::Wt(P), o0);
    const float2 hh = ld2(sm + RG::HC(T - 1), o0);
    const float out0 = ((fW.x + fE.x) + fS.x) + fN.x;
    const float out1 = ((fW.y + fE.y) + fS.y) + fN.y;
    const float in0 = ((fE_l + fW.y) + fN_s.x) + fS_n.x;
    const float in1 = ((fE.x + fW_r) + fN_s.y) + fS_n.y;
    const float nw0 = fmaxf(0.0f, fmaf(in0 - out0, TIMESTEP, w.x));
    const float nw1 = fmaxf(0.0f, fmaf(in1 - out1, TIMESTEP, w.y));
    st2(sm + RG::Wt(T), o0, nw0, nw1);
    st2(sm + RG::Ht(T), o0, nw0 + hh.x, nw1 + hh.y);
    if (T + 1 < I) st2(sm + RG::HC(T < I - 1 ? T : 0), o0, hh.x, hh.y);   // hand the height row to the next water stage
}

template <int I>
__device__ __forceinline__ void stage_velocity(const Lane& L, int s, int zc0, int zc1, const WaveParams& p, int xs0) {
    using RG = Rings<I>;
    constexpr int lag = 4 * I;
    const int r = s - lag;
    if (r < zc0 || r >= zc1) return;
    const int o0 = slot_of(L, lag, 0, r), os = slot_of(L, lag, -1, r), on = slot_of(L, lag, +1, r);
    float* sm = L.sm;
    const float2 fW = ld2(sm + RG::F(I, 0), o0), fE = ld2(sm + RG::F(I, 1), o0);
    const float fE_l = sm[RG::F(I, 1) + o0 + L.dl], fW_r = sm[RG::F(I, 0) + o0 + L.dr];
    const float2 fS = ld2(sm + RG::F(I, 2), o0), fN = ld2(sm + RG::F(I, 3), o0);
    const float2 fS_n = ld2(sm + RG::F(I, 2), on), fN_s = ld2(sm + RG::F(I, 3), os);
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
    const int gx = xs0 + L.c;
    if (L.c >= p.hx && L.c < FLW - p.hx && gx + 1 < p.W)
        *reinterpret_cast<float2*>(p.out + (size_t)r * p.W + gx) = make_float2(res[0], res[1]);
}

template <int I>
__device__ __forceinline__ void stage_load(const Lane& L, int s, int hlo, int hhi, const WaveParams& p, int xs0) {
    using RG = Rings<I>;
    if (s < hlo || s >= hhi) return;
    const float* g = p.h + (size_t)s * p.W;
    const int gx = xs0 + L.c;
    float a, b;
    if (gx >= 0 && gx + 1 < p.W) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(g + gx));
        a = v.x; b = v.y;
    } else {
        a = __ldg(g + min(max(gx, 0), p.W - 1));
        b = __ldg(g + min(max(gx + 1, 0), p.W - 1));
    }
    st2(L.sm + RG::HC(0), L.R[0], a, b);
}

// Static, cost-balanced roles: the 4 warps of a 64-column chunk split the 2I+1 stages of a step
//   role 0: outflow 1,2   role 1: outflow 3,4   role 2: outflow 5 + velocity   role 3: water 1..I-1 + loader
template <int I>
__global__ void __launch_bounds__(FL_THREADS, 1) flow_wave_kernel(WaveParams p) {
    extern __shared__ __align__(16) float sm[];
    const int H = p.H;
    const int xs0 = blockIdx.x * p.swi - p.hx;           // grid x of strip column 0 (even)
    const int zc0 = blockIdx.y * p.zc, zc1 = min(zc0 + p.zc, H);
    const int cmin = max(0, -xs0), cmax = min(FLW - 1, p.W - 1 - xs0);   // strip columns inside the grid
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int role = warp & 3;
    Lane L;
    L.sm = sm;
    L.c = (warp >> 2) * 64 + lane * 2;
    L.dl = max(L.c - 1, cmin) - L.c;
    L.dr = min(L.c + 2, cmax) - L.c;
    L.H = H;
    const int hlo = max(0, zc0 - 2 * I), hhi = min(H, zc1 + 2 * I);
    // first step, rounded down to a multiple of RING so that slot(s) = s mod RING starts at 0
    const int s_first = zc0 - 2 * I;
    const int s_begin = s_first - (((s_first % RING) + RING) % RING);
#pragma unroll
    for (int j = 0; j < RING; j++) L.R[j] = ((RING - j) % RING) * FLW + L.c;   // rows s_begin - j

    for (int s = s_begin; s < zc1 + 4 * I; s++) {
        if (role == 0) {
            stage_outflow<I, 1>(L, s, zc0, zc1);
            if (I >= 2) stage_outflow<I, (I >= 2 ? 2 : 1)>(L, s, zc0, zc1);
        } else if (role == 1) {
            if (I >= 3) stage_outflow<I, (I >= 3 ? 3 : 1)>(L, s, zc0, zc1);
            if (I >= 4) stage_outflow<I, (I >= 4 ? 4 : 1)>(L, s, zc0, zc1);
        } else if (role == 2) {
            if (I >= 5) stage_outflow<I, (I >= 5 ? 5 : 1)>(L, s, zc0, zc1);
            stage_velocity<I>(L, s, zc0, zc1, p, xs0);
        } else {
            stage_load<I>(L, s, hlo, hhi, p, xs0);
            if (I >= 2) stage_water<I, 1>(L, s, zc0, zc1);
            if (I >= 3) stage_water<I, (I >= 3 ? 2 : 1)>(L, s, zc0, zc1);
            if (I >= 4) stage_water<I, (I >= 4 ? 3 : 1)>(L, s, zc0, zc1);
            if (I >= 5) stage_water<I, (I >= 5 ? 4 : 1)>(L, s, zc0, zc1);
        }
        __syncthreads();
        // advance the slot registers: row s+1 takes the slot row s-4 vacates
        const int r4 = L.R[RING - 1];
#pragma unroll
        for (int j = RING - 1; j > 0; j--) L.R[j] = L.R[j - 1];
        L.R[0] = r4;
    }
}

size_t wave_smem_bytes(int I) {
    const int nhc = I - 1 > 1 ? I - 1 : 1;
    return (size_t)(4 * I + 2 * (I - 1) + nhc) * RING * FLW * sizeof(float);
}

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
