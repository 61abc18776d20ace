// flowtile_kernels.cu — the whole flow map on a TILE that lives in one SM (second fused formulation).
//
// Same arithmetic, cell for cell, as flow_kernels.cu / flowwave_kernels.cu / the oracle (FlowMapComponents.cs:20-165,
// FlowMapStage.cs:124-195); tests compare the three paths bitwise.
//
// flowwave_kernels.cu streams rows through warp-specialised stages and pays one CTA barrier per ROW (176 instructions
// per cell-iteration, 27-38 % barrier stalls).  Here a persistent CTA takes a 128 x 88 tile with a 2I-cell halo and runs
// all I iterations on it in place — overlapped (trapezoid) temporal blocking:
//   * shared memory holds the four outflow planes and H = water + height (20 B/cell: 128 x 88 cells = 220 KB — only a
//     227 KB/SM part can hold a tile large enough to keep 63 % of it useful);
//   * water and height of a cell stay in REGISTERS of the one thread that owns it for the whole tile (a warp owns 4
//     rows, a lane 4 adjacent columns of each);
//   * east/west neighbours come from warp shuffles of values the lane has loaded anyway, north/south neighbours are one
//     conflict-free LDS.128 each;
//   * every warp does the same work in every phase, so the 2I+1 barriers per TILE separate balanced phases.
// Clamp-to-edge (TileData.cs:72-77): a cell on the grid border uses its own value for the missing neighbour, exactly as
// the clamped index of the reference does; cells of a tile that lie outside the grid are never read by cells inside.
#include "nz_common.cuh"

namespace nz {
namespace {

constexpr int TW = 128;                 // tile columns: 32 lanes x 4
constexpr int FT_WARPS = 22;
constexpr int FT_G = 4;                 // rows per warp
constexpr int TH = FT_WARPS * FT_G;     // 88 tile rows
constexpr int FT_THREADS = FT_WARPS * 32;
constexpr int PLANE = TW * TH;          // floats per plane
constexpr int FT_SMEM = 5 * PLANE * 4;  // fW, fE, fS, fN, H
constexpr float TIMESTEP = 0.2f;
constexpr float WATER0 = 0.0001f;       // FillArrayJob value, FlowMapStage.cs:129

struct TileParams {
    const float* h;
    float* out;
    int W, H;
    int tiles_x, n_tiles;
    int hx, hz;        // halo columns (2I rounded up to 4) and rows (2I)
    float nmin, nrange, nsign;
    int zero_ok;
    // CYCLE form (one cycle of the subtractive-flow erosion, launch_flow_tile_cycle): the outflows come from fin (null: all
    // zero, the first cycle) instead of starting at zero, the final outflows of the tile's interior go to fout, and the
    // output is the eroded height  h - factor * normalised velocity  instead of the normalised velocity
    const float* fin[4];
    float* fout[4];
    float factor;
};

struct F4 {
    float v[4];
};
__device__ __forceinline__ F4 lds4(const float* p) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    F4 r;
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    return r;
}
__device__ __forceinline__ void sts4(float* p, const F4& a) { *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]); }

// ComputeFlowStep.CalculateCell for one cell — identical to flow_cell() of flowwave_kernels.cu
__device__ __forceinline__ void flow_cell(float H0, float HW, float HE, float HS, float HN, float w0, float fW, float fE,
                                          float fS, float fN, float& oW, float& oE, float& oS, float& oN) {
    const float flW = fmaxf(0.0f, fW + (H0 - HW));
    const float flE = fmaxf(0.0f, fE + (H0 - HE));
    const float flS = fmaxf(0.0f, fS + (H0 - HS));
    const float flN = fmaxf(0.0f, fN + (H0 - HN));
    const float sum_ = (flW + flE) + (flS + flN);
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

// One tile.  BORDER = false is the body for tiles that contain no grid border (the great majority): no clamp selects,
// no load guards.
template <int I, bool BORDER, bool CYCLE = false>
__device__ __forceinline__ void flow_tile_body(const TileParams& p, float* const pW, float* const pE, float* const pS,
                                               float* const pN, float* const pH, int warp, int col, int x0, int z0) {
        const int gx = x0 + col;
        const bool edgeL = BORDER && gx == 0, edgeR = BORDER && gx + 3 == p.W - 1;   // my cell 0 / cell 3 lies on the grid's west / east border
        float h[FT_G][4], w[FT_G][4];
        __syncthreads();                                               // the previous tile's velocity phase has read the planes
#pragma unroll
        for (int g = 0; g < FT_G; g++) {
            const int r = warp + FT_WARPS * g, gz = z0 + r;
            float4 t = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (!BORDER || (gz >= 0 && gz < p.H && gx >= 0 && gx + 3 < p.W)) t = __ldg(reinterpret_cast<const float4*>(p.h + (size_t)gz * p.W + gx));
            h[g][0] = t.x; h[g][1] = t.y; h[g][2] = t.z; h[g][3] = t.w;
            F4 H0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                w[g][q] = WATER0;
                H0.v[q] = WATER0 + h[g][q];
            }
            sts4(pH + r * TW + col, H0);
        }
        __syncthreads();

#pragma unroll 1
        for (int t = 0; t < I; t++) {
            // ---- outflow step of level t+1: reads H (neighbours), own water, own previous outflows; writes the outflows in place
#pragma unroll
            for (int g = 0; g < FT_G; g++) {
                const int r = warp + FT_WARPS * g, gz = z0 + r;
                const int o = r * TW + col;
                F4 H0, HS, HN, fW, fE, fS, fN;
#pragma unroll
                for (int q = 0; q < 4; q++) H0.v[q] = w[g][q] + h[g][q];
                HS = lds4(pH + max(r - 1, 0) * TW + col);
                HN = lds4(pH + min(r + 1, TH - 1) * TW + col);
                if (BORDER && gz == 0) HS = H0;
                if (BORDER && gz == p.H - 1) HN = H0;
                if (t == 0 && CYCLE && p.fin[0]) {
                    // the flow fields persist across the cycles of the subtractive-flow erosion: this cycle starts from them
                    const bool in = !BORDER || (gz >= 0 && gz < p.H && gx >= 0 && gx + 3 < p.W);
                    float4 a = make_float4(0.0f, 0.0f, 0.0f, 0.0f), b = a, c = a, d = a;
                    if (in) {
                        const size_t gi = (size_t)gz * p.W + gx;
                        a = __ldg(reinterpret_cast<const float4*>(p.fin[0] + gi));
                        b = __ldg(reinterpret_cast<const float4*>(p.fin[1] + gi));
                        c = __ldg(reinterpret_cast<const float4*>(p.fin[2] + gi));
                        d = __ldg(reinterpret_cast<const float4*>(p.fin[3] + gi));
                    }
                    fW.v[0] = a.x; fW.v[1] = a.y; fW.v[2] = a.z; fW.v[3] = a.w;
                    fE.v[0] = b.x; fE.v[1] = b.y; fE.v[2] = b.z; fE.v[3] = b.w;
                    fS.v[0] = c.x; fS.v[1] = c.y; fS.v[2] = c.z; fS.v[3] = c.w;
                    fN.v[0] = d.x; fN.v[1] = d.y; fN.v[2] = d.z; fN.v[3] = d.w;
                } else if (t == 0) {
#pragma unroll
                    for (int q = 0; q < 4; q++) fW.v[q] = fE.v[q] = fS.v[q] = fN.v[q] = 0.0f;
                } else {
                    fW = lds4(pW + o); fE = lds4(pE + o); fS = lds4(pS + o); fN = lds4(pN + o);
                }
                float HWl = __shfl_up_sync(0xffffffffu, H0.v[3], 1), HEr = __shfl_down_sync(0xffffffffu, H0.v[0], 1);
                if (edgeL) HWl = H0.v[0];
                if (edgeR) HEr = H0.v[3];
                F4 oW, oE, oS, oN;
#pragma unroll
                for (int q = 0; q < 4; q++)
                    flow_cell(H0.v[q], q == 0 ? HWl : H0.v[q > 0 ? q - 1 : 0], q == 3 ? HEr : H0.v[q < 3 ? q + 1 : 0], HS.v[q], HN.v[q],
                              w[g][q], fW.v[q], fE.v[q], fS.v[q], fN.v[q], oW.v[q], oE.v[q], oS.v[q], oN.v[q]);
                sts4(pW + o, oW); sts4(pE + o, oE); sts4(pS + o, oS); sts4(pN + o, oN);
            }
            __syncthreads();
            // ---- water step: reads the outflows (own + neighbours); writes own water (registers) and H
#pragma unroll
            for (int g = 0; g < FT_G; g++) {
                const int r = warp + FT_WARPS * g, gz = z0 + r;
                const int o = r * TW + col;
                const F4 fW = lds4(pW + o), fE = lds4(pE + o), fS = lds4(pS + o), fN = lds4(pN + o);
                F4 fN_s = lds4(pN + max(r - 1, 0) * TW + col), fS_n = lds4(pS + min(r + 1, TH - 1) * TW + col);
                if (BORDER && gz == 0) fN_s = fN;
                if (BORDER && gz == p.H - 1) fS_n = fS;
                float fE_l = __shfl_up_sync(0xffffffffu, fE.v[3], 1), fW_r = __shfl_down_sync(0xffffffffu, fW.v[0], 1);
                if (edgeL) fE_l = fE.v[0];
                if (edgeR) fW_r = fW.v[3];
                F4 nH;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const float out = ((fW.v[q] + fE.v[q]) + fS.v[q]) + fN.v[q];
                    const float in = (((q == 0 ? fE_l : fE.v[q > 0 ? q - 1 : 0]) + (q == 3 ? fW_r : fW.v[q < 3 ? q + 1 : 0])) + fN_s.v[q]) + fS_n.v[q];
                    w[g][q] = fmaxf(0.0f, fmaf(in - out, TIMESTEP, w[g][q]));
                    nH.v[q] = w[g][q] + h[g][q];
                }
                sts4(pH + o, nH);
            }
            __syncthreads();
        }

        // ---- velocity magnitude + normalisation of the tile's interior (CreateVelocityField + NormalizeMap)
#pragma unroll
        for (int g = 0; g < FT_G; g++) {
            const int r = warp + FT_WARPS * g, gz = z0 + r;
            if (r < p.hz || r >= TH - p.hz || (BORDER && gz >= p.H)) continue;          // warp-uniform
            const int o = r * TW + col;
            const F4 fW = lds4(pW + o), fE = lds4(pE + o), fS = lds4(pS + o), fN = lds4(pN + o);
            F4 fN_s = lds4(pN + (r - 1) * TW + col), fS_n = lds4(pS + (r + 1) * TW + col);
            if (BORDER && gz == 0) fN_s = fN;
            if (BORDER && gz == p.H - 1) fS_n = fS;
            float fE_l = __shfl_up_sync(0xffffffffu, fE.v[3], 1), fW_r = __shfl_down_sync(0xffffffffu, fW.v[0], 1);
            if (edgeL) fE_l = fE.v[0];
            if (edgeR) fW_r = fW.v[3];
            F4 res;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float dl = (q == 0 ? fE_l : fE.v[q > 0 ? q - 1 : 0]) - fW.v[q];
                const float dr = fE.v[q] - (q == 3 ? fW_r : fW.v[q < 3 ? q + 1 : 0]);
                const float dt = fS_n.v[q] - fN.v[q];
                const float db = fS.v[q] - fN_s.v[q];
                const float vx = (dl + dr) * 0.5f, vy = (dt + db) * 0.5f;
                float v = sqrtf(fmaf(vy, vy, vx * vx));
                if (p.nrange < 1e-12f) v = 0.0f;
                const float tt = v - p.nmin;
                res.v[q] = (tt == 0.0f && p.zero_ok) ? tt * p.nsign : tt / p.nrange;
                // ConstantMultiply then SubtractTiles (ErosionStageSubtractiveFlow.cs:196-222): one rounding each
                if (CYCLE) res.v[q] = h[g][q] - res.v[q] * p.factor;
            }
            if (col >= p.hx && col < TW - p.hx && (!BORDER || (gx >= 0 && gx + 3 < p.W))) {
                const size_t gi = (size_t)gz * p.W + gx;
                *reinterpret_cast<float4*>(p.out + gi) = make_float4(res.v[0], res.v[1], res.v[2], res.v[3]);
                if (CYCLE) {
                    *reinterpret_cast<float4*>(p.fout[0] + gi) = make_float4(fW.v[0], fW.v[1], fW.v[2], fW.v[3]);
                    *reinterpret_cast<float4*>(p.fout[1] + gi) = make_float4(fE.v[0], fE.v[1], fE.v[2], fE.v[3]);
                    *reinterpret_cast<float4*>(p.fout[2] + gi) = make_float4(fS.v[0], fS.v[1], fS.v[2], fS.v[3]);
                    *reinterpret_cast<float4*>(p.fout[3] + gi) = make_float4(fN.v[0], fN.v[1], fN.v[2], fN.v[3]);
                }
            }
        }
}

template <int I, bool CYCLE = false>
__global__ void __launch_bounds__(FT_THREADS, 1) flow_tile_kernel(TileParams p) {
    extern __shared__ __align__(16) float sm[];
    float* const pW = sm;
    float* const pE = sm + PLANE;
    float* const pS = sm + 2 * PLANE;
    float* const pN = sm + 3 * PLANE;
    float* const pH = sm + 4 * PLANE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col = 4 * lane;
    const int int_w = TW - 2 * p.hx, int_h = TH - 2 * p.hz;

    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int bx = tile % p.tiles_x, bz = tile / p.tiles_x;
        const int x0 = bx * int_w - p.hx, z0 = bz * int_h - p.hz;      // grid coordinates of tile cell (0,0); x0 % 4 == 0
        // the tile's cells lie strictly inside the grid: no cell has a missing neighbour
        const bool inside = x0 > 0 && z0 > 0 && x0 + TW < p.W && z0 + TH < p.H;
        if (inside)
            flow_tile_body<I, false, CYCLE>(p, pW, pE, pS, pN, pH, warp, col, x0, z0);
        else
            flow_tile_body<I, true, CYCLE>(p, pW, pE, pS, pN, pH, warp, col, x0, z0);
    }
}

}  // namespace

bool flow_tile_supported(int width, int rows, int iterations, const void* a, const void* b) {
    return iterations >= 1 && iterations <= 5 && (width & 3) == 0 && rows >= 1 && (((uintptr_t)a | (uintptr_t)b) & 15) == 0;
}

static TileParams tile_params(const float* d_height, float* d_out, int width, int rows, int I, float norm_min, float norm_max) {
    TileParams p;
    p.h = d_height; p.out = d_out; p.W = width; p.H = rows;
    p.hz = 2 * I;
    p.hx = (2 * I + 3) & ~3;
    p.tiles_x = cdiv(width, TW - 2 * p.hx);
    p.n_tiles = p.tiles_x * cdiv(rows, TH - 2 * p.hz);
    p.nmin = norm_min;
    p.nrange = norm_max - norm_min;
    p.nsign = copysignf(1.0f, p.nrange);
    p.zero_ok = (p.nrange != 0.0f) && isfinite(p.nrange);
    for (int k = 0; k < 4; k++) { p.fin[k] = nullptr; p.fout[k] = nullptr; }
    p.factor = 0.0f;
    return p;
}

// One cycle of the subtractive-flow erosion, fused (ErosionStageSubtractiveFlow.ScheduleCycle, :138-222): water := 1e-4,
// `iterations` x (outflow, water) starting from the outflow fields d_fin (4 planes; null: all zero), the final outflows to
// d_fout (4 planes), and d_out = d_height - factor * normalised |velocity|.  Nothing may alias: neighbouring tiles read the
// halo of the inputs while others write their interiors.  1 <= iterations <= 5.
int32_t launch_flow_tile_cycle(const float* d_height, float* d_out, const float* d_fin, float* d_fout, size_t plane, int width, int rows,
                               int iterations, float norm_min, float norm_max, float factor, cudaStream_t s) {
    static DeviceOnce attr_set;
    if (attr_set.need()) {
        NZ_CUDA(cudaFuncSetAttribute(flow_tile_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM));
        NZ_CUDA(cudaFuncSetAttribute(flow_tile_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM));
        NZ_CUDA(cudaFuncSetAttribute(flow_tile_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM));
        NZ_CUDA(cudaFuncSetAttribute(flow_tile_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM));
        NZ_CUDA(cudaFuncSetAttribute(flow_tile_kernel<5, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM));
        attr_set.mark();
    }
    TileParams p = tile_params(d_height, d_out, width, rows, iterations, norm_min, norm_max);
    for (int k = 0; k < 4; k++) {
        p.fin[k] = d_fin ? d_fin + k * plane : nullptr;
        p.fout[k] = d_fout + k * plane;
    }
    p.factor = factor;
    const int sms = sm_count();
    const int grid = p.n_tiles < sms ? p.n_tiles : sms;
    switch (iterations) {
        case 1: flow_tile_kernel<1, true><<<grid, FT_THREADS, FT_SMEM, s>>>(p); break;
        case 2: flow_tile_kernel<2, true><<<grid, FT_THREADS, FT_SMEM, s>>>(p); break;
        case 3: flow_tile_kernel<3, true><<<grid, FT_THREADS, FT_SMEM, s>>>(p); break;
        case 4: flow_tile_kernel<4, true><<<grid, FT_THREADS, FT_SMEM, s>>>(p); break;
        default: flow_tile_kernel<5, true><<<grid, FT_THREADS, FT_SMEM, s>>>(p); break;
    }
    NZ_LAUNCHED();
    return NZ_OK;
}

// d_out must not alias d_height
int32_t launch_flow_tile(const float* d_height, float* d_out, int width, int rows, int iterations, float norm_min, float norm_max,
                         cudaStream_t s) {
    static DeviceOnce attr_set;
    if (attr_set.need()) {
        NZ_CUDA(cudaFuncSetAttribute(flow_tile_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM));
        NZ_CUDA(cudaFuncSetAttribute(flow_tile_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM));
        NZ_CUDA(cudaFuncSetAttribute(flow_tile_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM));
        NZ_CUDA(cudaFuncSetAttribute(flow_tile_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM));
        NZ_CUDA(cudaFuncSetAttribute(flow_tile_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM));
        attr_set.mark();
    }
    const int I = iterations;
    TileParams p = tile_params(d_height, d_out, width, rows, I, norm_min, norm_max);
    const int sms = sm_count();
    const int grid = p.n_tiles < sms ? p.n_tiles : sms;
    switch (I) {
        case 1: flow_tile_kernel<1><<<grid, FT_THREADS, FT_SMEM, s>>>(p); break;
        case 2: flow_tile_kernel<2><<<grid, FT_THREADS, FT_SMEM, s>>>(p); break;
        case 3: flow_tile_kernel<3><<<grid, FT_THREADS, FT_SMEM, s>>>(p); break;
        case 4: flow_tile_kernel<4><<<grid, FT_THREADS, FT_SMEM, s>>>(p); break;
        default: flow_tile_kernel<5><<<grid, FT_THREADS, FT_SMEM, s>>>(p); break;
    }
    NZ_LAUNCHED();
    return NZ_OK;
}

}  // namespace nz
