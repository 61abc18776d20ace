// flowwave_kernels.cu — the whole flow map (fill, I x (outflow, water), velocity, normalise) in ONE launch.
//
// Same arithmetic, cell for cell, as flow_kernels.cu / the flow map of oracle/noize_oracle.cpp
// (FlowMapComponents.cs:20-165, FlowMapStage.cs:124-195).  What changes is where the state lives.
//
// Per-iteration kernels stream water + 4 flow fields through HBM every iteration (~44 B/cell/iteration,
// 61 GB at 16384^2 x 5).  But water and flows are DERIVED from the height field, so with all iterations
// fused only 4 B/cell are read and 4 B/cell written.  A square smem tile cannot hold it (24 B/cell of
// state, 2I halo on four sides), so the kernel streams instead (wavefront / time-skewed blocking):
//
//   a CTA owns a strip of FLW = 128 columns (2I halo columns each side) and walks DOWN a chunk of rows.
//   At step s it works on one row per pipeline stage, each stage lagging the previous one by 2 rows so that
//   everything a stage reads was produced in an EARLIER step (one __syncthreads per step):
//       L    : load height row s                                   (global -> smem)
//       A_t  : outflow step of level t on row s-(4t-2)   reads H_{t-1} = w_{t-1}+h (3 rows), w_{t-1}, f_{t-1}
//       B_t  : water   step of level t on row s-4t       reads f_t (3 rows), w_{t-1}, h;  writes w_t, H_t
//       V    : velocity + normalise      on row s-4I     reads f_I (3 rows);              writes the result row
//   Every field of every level is a ring of 5 rows (a row is last read 4 steps after it was written).
//   State per CTA: 5 x (4I + 3(I-1)) rows x 512 B = 80 KB for I = 5: two 512-thread CTAs per SM, so one CTA computes
//   while the other waits at its barrier.  A step costs about (instructions of the busiest warp) x ~5 cycles, so the
//   stages are split over as many warps as the register file allows: two 512-thread CTAs per SM at 56 registers.
//
// Warps are specialised: two warps per role (64 columns each); a lane owns 2 adjacent columns (LDS.64 / STS.64, only the
// two outer neighbours are scalar loads).  Roles are a static, cost-balanced split of the 2I+1 stages.  The
// byte offsets of the 5 ring slots rotate through 5 registers, so no modulo arithmetic is executed.
//
// Clamp-to-edge (TileData.cs:72-77) is applied where the reference applies it: neighbour column/row
// indices are clamped to the GRID, so a border cell reads its own current-level value.  At strip/chunk
// edges that are not grid edges the clamp yields garbage that advances one cell per stage and stays inside
// the halo.
#include "nz_common.cuh"

namespace nz {
namespace {

constexpr int FLW = 128;          // strip width in floats, halo included
constexpr int FL_THREADS = 512;
constexpr int RING = 5;
constexpr int VW = 2;             // columns per lane
constexpr int FLOW_WAVE_MAX_I = 5;
constexpr float TIMESTEP = 0.2f;
constexpr float WATER0 = 0.0001f;  // FillArrayJob value, FlowMapStage.cs:129

struct WaveParams {
    const float* h;
    float* out;
    int W, H;
    int zc;        // rows per chunk
    int swi;       // interior columns per strip = FLW - 2*hx
    int hx;        // halo columns each side: 2I (even, so strips start on a float2 boundary)
    float nmin, nrange;
    float nsign;   // copysign(1, nrange)
    int zero_ok;   // nrange is finite and non-zero: 0/nrange == 0*nsign
    const unsigned* gate;   // when set, the launch only runs if *gate == epoch (rerun requested by flow_walk_kernel)
    unsigned epoch;
    unsigned* reruns;       // with gate: counts the launches that did run
};

struct V4 {
    float v[VW];
};
// Shared-memory accessors: `slot` is an ABSOLUTE shared-window byte address (ring slot row + this lane's first
// column), FIELD a compile-time float offset of the ring (an LDS/STS immediate), `d` a per-lane float offset
// (outer neighbours).  Nothing but the access itself is executed per load: no base materialisation, no index scaling.
template <int FIELD>
__device__ __forceinline__ V4 ld4(uint32_t slot) {
    V4 r;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(r.v[0]), "=f"(r.v[1]) : "r"(slot), "n"(FIELD * 4));
    return r;
}
template <int FIELD>
__device__ __forceinline__ float ld1(uint32_t slot, int dbytes) {
    float r;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(r) : "r"(slot + dbytes), "n"(FIELD * 4));
    return r;
}
template <int FIELD>
__device__ __forceinline__ void st4(uint32_t slot, const V4& a) {
    asm volatile("st.shared.v2.f32 [%0+%1], {%2, %3};" ::"r"(slot), "n"(FIELD * 4), "f"(a.v[0]), "f"(a.v[1]) : "memory");
}
__device__ __forceinline__ V4 splat(float x) {
    V4 r;
#pragma unroll
    for (int q = 0; q < VW; q++) r.v[q] = x;
    return r;
}

// ComputeFlowStep.CalculateCell for one cell (W,E,S,N order; S = z-1, N = z+1)
__device__ __forceinline__ void flow_cell(float H0, float HW, float HE, float HS, float HN, float w0, float fW, float fE,
                                          float fS, float fN, float& oW, float& oE, float& oS, float& oN) {
    const float flW = fmaxf(0.0f, fW + (H0 - HW));
    const float flE = fmaxf(0.0f, fE + (H0 - HE));
    const float flS = fmaxf(0.0f, fS + (H0 - HS));
    const float flN = fmaxf(0.0f, fN + (H0 - HN));
    const float sum_ = (flW + flE) + (flS + flN);
    // K = clamp(w0 / (sum*dt), 0, 1).  The IEEE division's slow path (zero / denormal operands) costs ~60
    // instructions, and drained cells (w0 == 0) are common, so the two clamped outcomes are decided without
    // dividing: w0 >= d  =>  fl(w0/d) >= 1  => 1 ;  w0 <= 0  =>  quotient <= 0  => 0.  Same bits as the plain form.
    const float d = sum_ * TIMESTEP;
    // one divergent region (the quotient) instead of three nested branches; same values
    const bool pos = sum_ > 0.0f;
    const bool full = w0 >= d;
    float K = 0.0f;
    if (pos && !full && w0 > 0.0f) K = fminf(w0 / d, 1.0f);
    K = (pos && full) ? 1.0f : K;
    oW = pos ? flW * K : 0.0f;
    oE = pos ? flE * K : 0.0f;
    oS = pos ? flS * K : 0.0f;
    oN = pos ? flN * K : 0.0f;
}

// The two columns of a lane as one f32x2 pair (sm_100 FADD2 / FMUL2: one issue slot, two IEEE operations — the same
// bits as flow_cell() on each column).  Only max, the comparisons and the rare quotient stay scalar.  The pair
// (west neighbours) = (HWl, H0.x) and (east neighbours) = (H0.y, HEr) are the only operands that need assembling.
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 max0(float2 a) { return make_float2(fmaxf(0.0f, a.x), fmaxf(0.0f, a.y)); }
__device__ __forceinline__ float quot_k(float sum_, float d, float w0) {
    const bool pos = sum_ > 0.0f;
    const bool full = w0 >= d;
    float K = 0.0f;
    if (pos && !full && w0 > 0.0f) K = fminf(w0 / d, 1.0f);
    return (pos && full) ? 1.0f : K;
}
__device__ __forceinline__ void flow_cell2(float2 H0, float HWl, float HEr, float2 HS, float2 HN, float2 w0, float2 fW, float2 fE,
                                           float2 fS, float2 fN, float2& oW, float2& oE, float2& oS, float2& oN) {
    const float2 flW = max0(__fadd2_rn(fW, sub2(H0, f2(HWl, H0.x))));
    const float2 flE = max0(__fadd2_rn(fE, sub2(H0, f2(H0.y, HEr))));
    const float2 flS = max0(__fadd2_rn(fS, sub2(H0, HS)));
    const float2 flN = max0(__fadd2_rn(fN, sub2(H0, HN)));
    const float2 sum_ = __fadd2_rn(__fadd2_rn(flW, flE), __fadd2_rn(flS, flN));
    const float2 d = __fmul2_rn(sum_, f2(TIMESTEP, TIMESTEP));
    const float2 K = f2(quot_k(sum_.x, d.x, w0.x), quot_k(sum_.y, d.y, w0.y));
    const bool p0 = sum_.x > 0.0f, p1 = sum_.y > 0.0f;
    const float2 a = __fmul2_rn(flW, K), b = __fmul2_rn(flE, K), c = __fmul2_rn(flS, K), e = __fmul2_rn(flN, K);
    oW = f2(p0 ? a.x : 0.0f, p1 ? a.y : 0.0f);
    oE = f2(p0 ? b.x : 0.0f, p1 ? b.y : 0.0f);
    oS = f2(p0 ? c.x : 0.0f, p1 ? c.y : 0.0f);
    oN = f2(p0 ? e.x : 0.0f, p1 ? e.y : 0.0f);
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
    uint32_t R[RING];   // R[j] = shared-window byte address of the slot of row s-j, at this lane's first column
    int dl, dr;         // BYTE offsets (from this lane's first column) of the clamped west / east outer neighbours
    int c;              // first of this lane's VW strip columns
    int H;              // grid rows
};

// slot of row (s - lag + d), d in {-1,0,+1}; rows outside the grid clamp onto the border row
__device__ __forceinline__ uint32_t slot_of(const Lane& L, int lag, int d, int r) {
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
    const uint32_t o0 = slot_of(L, lag, 0, r), os = slot_of(L, lag, -1, r), on = slot_of(L, lag, +1, r);
    V4 H0, HS, HN, w0, fW, fE, fS, fN;
    float HWl, HEr;
    if (T == 1) {
        // level 0: water == 1e-4 everywhere, flows == 0: H_0 = 1e-4 + h computed on the fly
        const V4 a = ld4<RG::HC(0)>(o0), b = ld4<RG::HC(0)>(os), d = ld4<RG::HC(0)>(on);
#pragma unroll
        for (int q = 0; q < VW; q++) {
            H0.v[q] = WATER0 + a.v[q];
            HS.v[q] = WATER0 + b.v[q];
            HN.v[q] = WATER0 + d.v[q];
        }
        HWl = WATER0 + ld1<RG::HC(0)>(o0, L.dl);
        HEr = WATER0 + ld1<RG::HC(0)>(o0, L.dr);
        w0 = splat(WATER0);
        fW = fE = fS = fN = splat(0.0f);
    } else {
        constexpr int P = T > 1 ? T - 1 : 1;
        H0 = ld4<RG::Ht(P)>(o0); HS = ld4<RG::Ht(P)>(os); HN = ld4<RG::Ht(P)>(on);
        HWl = ld1<RG::Ht(P)>(o0, L.dl); HEr = ld1<RG::Ht(P)>(o0, L.dr);
        w0 = ld4<RG::Wt(P)>(o0);
        fW = ld4<RG::F(P, 0)>(o0); fE = ld4<RG::F(P, 1)>(o0);
        fS = ld4<RG::F(P, 2)>(o0); fN = ld4<RG::F(P, 3)>(o0);
    }
    V4 oW, oE, oS, oN;
    {
        static_assert(VW == 2, "flow_cell2 pairs the two columns of a lane");
        float2 pW, pE, pS, pN;
        flow_cell2(f2(H0.v[0], H0.v[1]), HWl, HEr, f2(HS.v[0], HS.v[1]), f2(HN.v[0], HN.v[1]), f2(w0.v[0], w0.v[1]),
                   f2(fW.v[0], fW.v[1]), f2(fE.v[0], fE.v[1]), f2(fS.v[0], fS.v[1]), f2(fN.v[0], fN.v[1]), pW, pE, pS, pN);
        oW.v[0] = pW.x; oW.v[1] = pW.y; oE.v[0] = pE.x; oE.v[1] = pE.y;
        oS.v[0] = pS.x; oS.v[1] = pS.y; oN.v[0] = pN.x; oN.v[1] = pN.y;
    }
    st4<RG::F(T, 0)>(o0, oW);
    st4<RG::F(T, 1)>(o0, oE);
    st4<RG::F(T, 2)>(o0, oS);
    st4<RG::F(T, 3)>(o0, oN);
}

template <int I, int T>
__device__ __forceinline__ void stage_water(const Lane& L, int s, int zc0, int zc1) {
    using RG = Rings<I>;
    constexpr int lag = 4 * T;
    const int r = s - lag;
    const int lo = max(0, zc0 - (2 * I - 2 * T)), hi = min(L.H, zc1 + (2 * I - 2 * T));
    if (r < lo || r >= hi) return;
    const uint32_t o0 = slot_of(L, lag, 0, r), os = slot_of(L, lag, -1, r), on = slot_of(L, lag, +1, r);
    const V4 fW = ld4<RG::F(T, 0)>(o0), fE = ld4<RG::F(T, 1)>(o0);
    const V4 fS = ld4<RG::F(T, 2)>(o0), fN = ld4<RG::F(T, 3)>(o0);
    const float fE_l = ld1<RG::F(T, 1)>(o0, L.dl), fW_r = ld1<RG::F(T, 0)>(o0, L.dr);
    const V4 fN_s = ld4<RG::F(T, 3)>(os), fS_n = ld4<RG::F(T, 2)>(on);
    constexpr int P = T > 1 ? T - 1 : 1;
    const V4 w = (T == 1) ? splat(WATER0) : ld4<RG::Wt(P)>(o0);
    const V4 hh = ld4<RG::HC(T - 1)>(o0);
    V4 nw, nH;
    {
        // the lane's two columns as one f32x2 pair (same operations, same order, same bits as the scalar form)
        const float2 pW = f2(fW.v[0], fW.v[1]), pE = f2(fE.v[0], fE.v[1]);
        const float2 out = __fadd2_rn(__fadd2_rn(__fadd2_rn(pW, pE), f2(fS.v[0], fS.v[1])), f2(fN.v[0], fN.v[1]));
        const float2 in = __fadd2_rn(__fadd2_rn(__fadd2_rn(f2(fE_l, fE.v[0]), f2(fW.v[1], fW_r)), f2(fN_s.v[0], fN_s.v[1])),
                                     f2(fS_n.v[0], fS_n.v[1]));
        const float2 w2 = max0(__ffma2_rn(sub2(in, out), f2(TIMESTEP, TIMESTEP), f2(w.v[0], w.v[1])));
        const float2 H2 = __fadd2_rn(w2, f2(hh.v[0], hh.v[1]));
        nw.v[0] = w2.x; nw.v[1] = w2.y;
        nH.v[0] = H2.x; nH.v[1] = H2.y;
    }
    st4<RG::Wt(T)>(o0, nw);
    st4<RG::Ht(T)>(o0, nH);
    if (T + 1 < I) st4<RG::HC(T < I - 1 ? T : 0)>(o0, hh);   // hand the height row to the next water stage
}

template <int I>
__device__ __forceinline__ void stage_velocity(const Lane& L, int s, int zc0, int zc1, const WaveParams& p, int xs0) {
    using RG = Rings<I>;
    constexpr int lag = 4 * I;
    const int r = s - lag;
    if (r < zc0 || r >= zc1) return;
    const uint32_t o0 = slot_of(L, lag, 0, r), os = slot_of(L, lag, -1, r), on = slot_of(L, lag, +1, r);
    const V4 fW = ld4<RG::F(I, 0)>(o0), fE = ld4<RG::F(I, 1)>(o0);
    const float fE_l = ld1<RG::F(I, 1)>(o0, L.dl), fW_r = ld1<RG::F(I, 0)>(o0, L.dr);
    const V4 fS = ld4<RG::F(I, 2)>(o0), fN = ld4<RG::F(I, 3)>(o0);
    const V4 fS_n = ld4<RG::F(I, 2)>(on), fN_s = ld4<RG::F(I, 3)>(os);
    V4 res;
#pragma unroll
    for (int q = 0; q < VW; q++) {
        const float dl = (q == 0 ? fE_l : fE.v[q > 0 ? q - 1 : 0]) - fW.v[q];
        const float dr = fE.v[q] - (q == VW - 1 ? fW_r : fW.v[q < VW - 1 ? q + 1 : 0]);
        const float dt = fS_n.v[q] - fN.v[q];
        const float db = fS.v[q] - fN_s.v[q];
        const float vx = (dl + dr) * 0.5f, vy = (dt + db) * 0.5f;
        float v = sqrtf(fmaf(vy, vy, vx * vx));
        if (p.nrange < 1e-12f) v = 0.0f;
        const float t = v - p.nmin;
        // still water (t == +-0) is common and 0/x takes the division's slow path; 0/x == 0 * sign(x) for finite x != 0
        res.v[q] = (t == 0.0f && p.zero_ok) ? t * p.nsign : t / p.nrange;
    }
    const int gx = xs0 + L.c;
    if (L.c >= p.hx && L.c < FLW - p.hx && gx + VW - 1 < p.W)
        *reinterpret_cast<float2*>(p.out + (size_t)r * p.W + gx) = make_float2(res.v[0], res.v[1]);
}

// Loader: the global read of a height row has ~1 us of latency, far more than a step takes, so the row that is
// stored to shared memory at step s was requested PF steps earlier and waited in registers.
constexpr int PF = 4;
__device__ __forceinline__ V4 load_row(const Lane& L, int row, int hlo, int hhi, const WaveParams& p, int xs0) {
    V4 a = splat(0.0f);
    if (row >= hlo && row < hhi) {
        const float* g = p.h + (size_t)row * p.W;
        const int gx = xs0 + L.c;
        if (gx >= 0 && gx + VW - 1 < p.W) {
            const float2 t = __ldg(reinterpret_cast<const float2*>(g + gx));
            a.v[0] = t.x; a.v[1] = t.y;
        } else {
#pragma unroll
            for (int q = 0; q < VW; q++) a.v[q] = __ldg(g + min(max(gx + q, 0), p.W - 1));
        }
    }
    return a;
}

// Static, cost-balanced roles, two warps each (2 x 32 lanes x 2 columns = the 128-column strip).
//   role 0..4: outflow of level role+1      role 5: velocity
//   role 6: loader + water 1,2              role 7: water 3,4
template <int I>
__global__ void __launch_bounds__(FL_THREADS, 2) flow_wave_kernel(WaveParams p) {
    extern __shared__ __align__(16) float sm[];
    if (p.gate) {
        if (*p.gate != p.epoch) return;
        if (p.reruns && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) atomicAdd(p.reruns, 1u);
    }
    const int H = p.H;
    const int xs0 = blockIdx.x * p.swi - p.hx;           // grid x of strip column 0 (even)
    const int zc0 = blockIdx.y * p.zc, zc1 = min(zc0 + p.zc, H);
    const int cmin = max(0, -xs0), cmax = min(FLW - 1, p.W - 1 - xs0);   // strip columns inside the grid
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int role = warp >> 1;
    const uint32_t sm_base = (uint32_t)__cvta_generic_to_shared(sm);
    Lane L;
    L.c = (warp & 1) * 64 + lane * VW;
    L.dl = (max(L.c - 1, cmin) - L.c) * 4;
    L.dr = (min(L.c + VW, cmax) - L.c) * 4;
    L.H = H;
    const int hlo = max(0, zc0 - 2 * I), hhi = min(H, zc1 + 2 * I);
    // first step, rounded down to a multiple of RING so that slot(s) = s mod RING starts at 0
    const int s_first = zc0 - 2 * I;
    const int s_begin = s_first - (((s_first % RING) + RING) % RING);
#pragma unroll
    for (int j = 0; j < RING; j++) L.R[j] = sm_base + (((RING - j) % RING) * FLW + L.c) * 4;   // rows s_begin - j

    V4 pf[PF];
    if (role == 6) {
#pragma unroll
        for (int j = 0; j < PF; j++) pf[j] = load_row(L, s_begin + j, hlo, hhi, p, xs0);
    }

    for (int s = s_begin; s < zc1 + 4 * I; s++) {
        switch (role) {
            case 0: stage_outflow<I, 1>(L, s, zc0, zc1); break;
            case 1: if (I >= 2) stage_outflow<I, (I >= 2 ? 2 : 1)>(L, s, zc0, zc1); break;
            case 2: if (I >= 3) stage_outflow<I, (I >= 3 ? 3 : 1)>(L, s, zc0, zc1); break;
            case 3: if (I >= 4) stage_outflow<I, (I >= 4 ? 4 : 1)>(L, s, zc0, zc1); break;
            case 4: if (I >= 5) stage_outflow<I, (I >= 5 ? 5 : 1)>(L, s, zc0, zc1); break;
            case 5: stage_velocity<I>(L, s, zc0, zc1, p, xs0); break;
            case 6:
                st4<Rings<I>::HC(0)>(L.R[0], pf[0]);          // row s, requested PF steps ago
#pragma unroll
                for (int j = 0; j + 1 < PF; j++) pf[j] = pf[j + 1];
                pf[PF - 1] = load_row(L, s + PF, hlo, hhi, p, xs0);
                if (I >= 2) stage_water<I, 1>(L, s, zc0, zc1);
                if (I >= 3) stage_water<I, (I >= 3 ? 2 : 1)>(L, s, zc0, zc1);
                break;
            default:
                if (I >= 4) stage_water<I, (I >= 4 ? 3 : 1)>(L, s, zc0, zc1);
                if (I >= 5) stage_water<I, (I >= 5 ? 4 : 1)>(L, s, zc0, zc1);
                break;
        }
        __syncthreads();
        // advance the slot registers: row s+1 takes the slot row s-4 vacates
        const uint32_t r4 = L.R[RING - 1];
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

// a, b: the two grid pointers (may be null when only the shape is being asked about)
bool flow_wave_supported(int width, int rows, int iterations, const void* a, const void* b) {
    return iterations >= 1 && iterations <= FLOW_WAVE_MAX_I && (width & 3) == 0 && rows >= 1 &&
           (((uintptr_t)a | (uintptr_t)b) & 15) == 0;
}

// d_out must not alias d_height
int32_t launch_flow_wave(const float* d_height, float* d_out, int width, int rows, int iterations, float norm_min,
                         float norm_max, cudaStream_t s, const unsigned* gate, unsigned epoch, unsigned* reruns) {
    static DeviceOnce attr_set;
    if (attr_set.need()) {
        NZ_CUDA(cudaFuncSetAttribute(flow_wave_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wave_smem_bytes(1)));
        NZ_CUDA(cudaFuncSetAttribute(flow_wave_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wave_smem_bytes(2)));
        NZ_CUDA(cudaFuncSetAttribute(flow_wave_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wave_smem_bytes(3)));
        NZ_CUDA(cudaFuncSetAttribute(flow_wave_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wave_smem_bytes(4)));
        NZ_CUDA(cudaFuncSetAttribute(flow_wave_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wave_smem_bytes(5)));
        attr_set.mark();
    }
    const int I = iterations;
    WaveParams p;
    p.h = d_height; p.out = d_out; p.W = width; p.H = rows;
    p.gate = gate; p.epoch = epoch; p.reruns = reruns;
    p.hx = 2 * I;
    p.swi = FLW - 2 * p.hx;
    p.nmin = norm_min;
    p.nrange = norm_max - norm_min;
    p.nsign = copysignf(1.0f, p.nrange);
    p.zero_ok = (p.nrange != 0.0f) && isfinite(p.nrange);
    const int strips = cdiv(width, p.swi);
    // rows per chunk: long enough to amortise the 6I-row pipeline fill, and a CTA count that fills whole waves
    const int sms = sm_count();
    int best_nz = 1;
    double best_cost = 1e300;
    for (int nz = 1; nz <= rows && nz <= 4096; nz++) {
        const int zc = cdiv(rows, nz);
        if (zc < 64 && nz > 1) break;
        const long long ctas = (long long)strips * cdiv(rows, zc);
        const long long slots = 2LL * sms;                // two CTAs are resident per SM and hide each other's barrier waits
        const long long waves = (ctas + slots - 1) / slots;
        // steps executed by the busiest SM slot; a lone CTA on an SM runs its steps ~1.7x slower than one of a pair
        const double cost = (double)waves * (zc + 6 * I + RING) * (ctas < slots ? 1.7 : 1.0);
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
