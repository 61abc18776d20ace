// pointwise_kernels.cu — the element-wise stage set: HBM-bound maps.
//
//   constant   ConstantJob<ConstantMultiply|ConstantBinarize>   Filter/ConstantJob.cs:16-47, Operators/SimpleMutation.cs:16-54
//   reduce     ReductionJob<Subtract|Multiply|RootSumSquares|Max|Min>  Filter/ReductionJob.cs:16-53, SimpleMutation.cs:56-171
//   curve      CurveJob<CurveOperator>                          Filter/Curve/CurveJob.cs:56-89
//   crop       CropJob                                          Filter/Sample/CropJob.cs:18-61
//   map range  GetMapRangeJob                                   Filter/NormalizeJob.cs:18-53
//   normalize  MapNormalizeValues<NormalizeMap>                 Filter/NormalizeJob.cs:58-92, FlowMapComponents.cs:150-166
//
// Every map reads each input once and writes each output once (4 B in / 4 B out per cell and operand): the bound is
// HBM bandwidth.  One thread moves one float4 (16 B per lane, 512 B per warp instruction); grids cover the array
// once, sized in whole CTAs of 256 threads.  The reference writes into `tmp` and copies back (SWAP_RWTILE); these
// maps are in place, which moves the same information.  Arithmetic is the oracle's, op for op.
#include <float.h>
#include "nz_common.cuh"

namespace nz {
namespace {

constexpr int PW_THREADS = 256;

template <typename F>
__global__ void __launch_bounds__(PW_THREADS) map1_kernel(float* __restrict__ a, size_t n, F f) {
    const size_t i4 = ((size_t)blockIdx.x * PW_THREADS + threadIdx.x) * 4;
    if (i4 + 3 < n && (((uintptr_t)a) & 15) == 0) {
        float4 v = *reinterpret_cast<const float4*>(a + i4);
        v.x = f(v.x); v.y = f(v.y); v.z = f(v.z); v.w = f(v.w);
        *reinterpret_cast<float4*>(a + i4) = v;
    } else {
        for (size_t i = i4; i < n && i < i4 + 4; i++) a[i] = f(a[i]);
    }
}

template <typename F>
__global__ void __launch_bounds__(PW_THREADS) map2_kernel(float* __restrict__ a, const float* __restrict__ b, size_t n, F f) {
    const size_t i4 = ((size_t)blockIdx.x * PW_THREADS + threadIdx.x) * 4;
    if (i4 + 3 < n && ((((uintptr_t)a) | ((uintptr_t)b)) & 15) == 0) {
        float4 v = *reinterpret_cast<const float4*>(a + i4);
        const float4 w = __ldg(reinterpret_cast<const float4*>(b + i4));
        v.x = f(v.x, w.x); v.y = f(v.y, w.y); v.z = f(v.z, w.z); v.w = f(v.w, w.w);
        *reinterpret_cast<float4*>(a + i4) = v;
    } else {
        for (size_t i = i4; i < n && i < i4 + 4; i++) a[i] = f(a[i], b[i]);
    }
}

struct MulC { float c; __device__ float operator()(float v) const { return v * c; } };
struct BinC { float c; __device__ float operator()(float v) const { return v >= c ? 1.0f : 0.0f; } };
struct Sub { __device__ float operator()(float a, float b) const { return a - b; } };
struct Mul { __device__ float operator()(float a, float b) const { return a * b; } };
struct Rss { __device__ float operator()(float a, float b) const { return sqrtf(fmaf(b, b, a * a)); } };
struct Max { __device__ float operator()(float a, float b) const { return fmaxf(a, b); } };
struct Min { __device__ float operator()(float a, float b) const { return fminf(a, b); } };
// NormalizeMap.CalculateCell, FlowMapComponents.cs:157-165: args = {min, max, range}
struct Norm {
    float a0, a2;
    __device__ float operator()(float v) const {
        if (a2 < 1e-12f) v = 0.0f;
        return (v - a0) / a2;
    }
};
// CurveOperator.Apply, Curve/CurveJob.cs:69-80.  The curve (<= a few hundred floats) is read through the
// read-only cache; every lane of a warp usually hits the same few lines.
struct Curve {
    const float* curve;
    float size;        // (float)CurveSize
    __device__ float operator()(float v) const {
        const float rect = fminf(fmaxf(v, 0.0f), 1.0f) * size;
        const float lower = fminf(floorf(rect), size - 2.0f);
        const int li = (int)lower;
        const float left = __ldg(curve + li), right = __ldg(curve + li + 1);
        float value = fmaf(rect - lower, right - left, left);   // math.lerp
        value = fmaxf(0.0f, value);
        return fminf(1.0f, value);
    }
};

__global__ void __launch_bounds__(PW_THREADS) crop_kernel(const float* __restrict__ in, int in_res, float* __restrict__ out, int out_res,
                                                          int offset) {
    const int x = blockIdx.x * PW_THREADS + threadIdx.x, z = blockIdx.y;
    if (x >= out_res) return;
    // ReadTileData.GetData clamps to the edge (Pipeline/Tiles/TileData.cs:100-109)
    const int sx = min(max(x + offset, 0), in_res - 1), sz = min(max(z + offset, 0), in_res - 1);
    out[(size_t)z * out_res + x] = __ldg(in + (size_t)sz * in_res + sx);
}

// GetMapRangeJob: min_ = min(min_, map[i]), max_ likewise, starting from +inf / -inf (or the caller's limits).
// min/max are exact and order-independent (math.min/max return the non-NaN operand like fminf/fmaxf), so a tree
// reduction gives the serial loop's result.  Two launches: per-CTA partials, then one CTA folds them.
constexpr int RANGE_CTAS = 148 * 8;
__global__ void __launch_bounds__(PW_THREADS) range_partial_kernel(const float* __restrict__ a, size_t n, float lim_min, float lim_max,
                                                                   float2* __restrict__ part) {
    float lo = lim_min, hi = lim_max;
    const size_t stride = (size_t)gridDim.x * PW_THREADS * 4;
    for (size_t i4 = ((size_t)blockIdx.x * PW_THREADS + threadIdx.x) * 4; i4 < n; i4 += stride) {
        if (i4 + 3 < n && (((uintptr_t)a) & 15) == 0) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(a + i4));
            lo = fminf(fminf(lo, v.x), fminf(v.y, fminf(v.z, v.w)));
            hi = fmaxf(fmaxf(hi, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
        } else {
            for (size_t i = i4; i < n && i < i4 + 4; i++) {
                lo = fminf(lo, a[i]);
                hi = fmaxf(hi, a[i]);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    __shared__ float2 w[PW_THREADS / 32];
    if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = make_float2(lo, hi);
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < PW_THREADS / 32; i++) {
            lo = fminf(lo, w[i].x);
            hi = fmaxf(hi, w[i].y);
        }
        part[blockIdx.x] = make_float2(lo, hi);
    }
}
__global__ void __launch_bounds__(PW_THREADS) range_final_kernel(const float2* __restrict__ part, int nparts, float* __restrict__ res) {
    float lo = INFINITY, hi = -INFINITY;
    for (int i = threadIdx.x; i < nparts; i += PW_THREADS) {
        lo = fminf(lo, part[i].x);
        hi = fmaxf(hi, part[i].y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    __shared__ float2 w[PW_THREADS / 32];
    if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = make_float2(lo, hi);
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < PW_THREADS / 32; i++) {
            lo = fminf(lo, w[i].x);
            hi = fmaxf(hi, w[i].y);
        }
        res[0] = lo;          // GetMapRangeJob.MIN
        res[1] = hi;          // MAX
        res[2] = hi - lo;     // RANGE
    }
}

// A run of element-wise stages on the same grid as ONE pass (second half): the host layer defers
// constant / normalize inside a residency scope and applies the run together with the next curve, or when another stage
// (or the download) needs the values.  Each step is the functor the single-stage kernel uses, applied in call order to the
// value in a register, so the result is the one the separate passes give bit for bit (--fmad=false: nothing contracts
// across steps), for 8 B of HBM traffic per cell instead of 8 B per stage.
struct ChainF {
    PointwiseOp op[PW_CHAIN_MAX];
    int count;
    __device__ float operator()(float v) const {
        for (int i = 0; i < count; i++) {
            const PointwiseOp o = op[i];
            switch (o.kind) {
                case PW_MUL: v = MulC{o.a}(v); break;
                case PW_BINARIZE: v = BinC{o.a}(v); break;
                case PW_NORMALIZE: v = Norm{o.a, o.b}(v); break;
                default: v = Curve{o.lut, o.a}(v); break;
            }
        }
        return v;
    }
};

inline int grid_for(size_t n) { return cdiv((long long)((n + 3) / 4), PW_THREADS); }

}  // namespace

int32_t launch_constant(float* d, size_t n, int op, float value, cudaStream_t s) {
    if (n == 0) return NZ_OK;
    if (op == NZ_CONSTANT_MULTIPLY) map1_kernel<<<grid_for(n), PW_THREADS, 0, s>>>(d, n, MulC{value});
    else if (op == NZ_CONSTANT_BINARIZE) map1_kernel<<<grid_for(n), PW_THREADS, 0, s>>>(d, n, BinC{value});
    else {
        set_error("nz_constant: operation %d out of range", op);
        return NZ_E_INVALID;
    }
    NZ_LAUNCHED();
    return NZ_OK;
}

int32_t launch_reduce(float* d_left, const float* d_right, size_t n, int op, cudaStream_t s) {
    if (n == 0) return NZ_OK;
    const int g = grid_for(n);
    switch (op) {
        case NZ_REDUCE_SUBTRACT: map2_kernel<<<g, PW_THREADS, 0, s>>>(d_left, d_right, n, Sub{}); break;
        case NZ_REDUCE_MULTIPLY: map2_kernel<<<g, PW_THREADS, 0, s>>>(d_left, d_right, n, Mul{}); break;
        case NZ_REDUCE_ROOTSUMSQUARES: map2_kernel<<<g, PW_THREADS, 0, s>>>(d_left, d_right, n, Rss{}); break;
        case NZ_REDUCE_MAX: map2_kernel<<<g, PW_THREADS, 0, s>>>(d_left, d_right, n, Max{}); break;
        case NZ_REDUCE_MIN: map2_kernel<<<g, PW_THREADS, 0, s>>>(d_left, d_right, n, Min{}); break;
        default:
            set_error("nz_reduce: operation %d out of range", op);
            return NZ_E_INVALID;
    }
    NZ_LAUNCHED();
    return NZ_OK;
}

int32_t launch_curve(float* d, size_t n, const float* d_curve, int curve_size, cudaStream_t s) {
    if (n == 0) return NZ_OK;
    map1_kernel<<<grid_for(n), PW_THREADS, 0, s>>>(d, n, Curve{d_curve, (float)curve_size});
    NZ_LAUNCHED();
    return NZ_OK;
}

int32_t launch_normalize(float* d, size_t n, float vmin, float range, cudaStream_t s) {
    if (n == 0) return NZ_OK;
    map1_kernel<<<grid_for(n), PW_THREADS, 0, s>>>(d, n, Norm{vmin, range});
    NZ_LAUNCHED();
    return NZ_OK;
}

int32_t launch_pointwise_chain(float* d, size_t n, const PointwiseOp* ops, int count, cudaStream_t s) {
    if (n == 0 || count == 0) return NZ_OK;
    if (count < 0 || count > PW_CHAIN_MAX) {
        set_error("pointwise chain: %d steps (at most %d)", count, PW_CHAIN_MAX);
        return NZ_E_INVALID;
    }
    ChainF f;
    for (int i = 0; i < count; i++) f.op[i] = ops[i];
    for (int i = count; i < PW_CHAIN_MAX; i++) f.op[i] = PointwiseOp{PW_MUL, 1.0f, 0.0f, nullptr};
    f.count = count;
    map1_kernel<<<grid_for(n), PW_THREADS, 0, s>>>(d, n, f);
    NZ_LAUNCHED();
    return NZ_OK;
}

int32_t launch_crop(const float* d_in, int in_res, float* d_out, int out_res, int offset, cudaStream_t s) {
    dim3 grid(cdiv(out_res, PW_THREADS), out_res);
    crop_kernel<<<grid, PW_THREADS, 0, s>>>(d_in, in_res, d_out, out_res, offset);
    NZ_LAUNCHED();
    return NZ_OK;
}

size_t map_range_scratch_bytes() { return RANGE_CTAS * sizeof(float2); }

int32_t launch_map_range(const float* d, size_t n, float lim_min, float lim_max, float* d_res3, void* d_scratch, cudaStream_t s) {
    int g = grid_for(n);
    if (g > RANGE_CTAS) g = RANGE_CTAS;
    if (g < 1) g = 1;
    range_partial_kernel<<<g, PW_THREADS, 0, s>>>(d, n, lim_min, lim_max, (float2*)d_scratch);
    NZ_LAUNCHED();
    range_final_kernel<<<1, PW_THREADS, 0, s>>>((const float2*)d_scratch, g, d_res3);
    NZ_LAUNCHED();
    return NZ_OK;
}

}  // namespace nz
