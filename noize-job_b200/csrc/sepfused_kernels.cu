// sepfused_kernels.cu — temporally blocked separable filter: T iterations of (X pass, Z pass) per launch.
//
// Same arithmetic as the one-pass kernels in filter_kernels.cu (and oracle generic_kernel_job):
//   X: total = 0; for k=-r..r  total = fma(src(x+k,z), K[r+k], total); out = total*factor
//   Z: total = 0; for k=r..-r  total = fma(src(x,z+k), K[r-k], total); out = total*factor
// with clamp-to-edge reads re-applied at every pass (Pipeline/Tiles/TileData.cs:72-77), i.e. the
// reference's KernelFilterStage loop (Filter/KernelFilterStage.cs:31-43) over
// SeparableKernelFilter.ScheduleSeries (Filter/Kernel/KernelJob.cs:165-185).
//
// Why: one (X,Z) iteration per launch moves 8 B/cell through HBM per iteration (136 B/cell for
// Gauss5 x17).  Here a CTA stages a 128x96 tile (8-cell halo) in shared memory, runs up to
// T = 8/r iterations on it ping-ponging between two smem buffers, and writes the 112x80 interior:
// 8 B/cell per T iterations.  The stage then is bound by shared-memory bandwidth / FFMA issue,
// not HBM ("Gauss5x17 fully fused is compute/smem-bound").
//
// Work decomposition inside a CTA (256 threads = 8 warps): 2 warps side by side cover the 128
// columns (each lane owns 2 adjacent columns -> LDS.64/STS.64), 4 warp rows split the rows of the
// iteration's window.  A warp walks DOWN its row segment: per step it computes the X pass of one
// new row from shared memory (3 x LDS.64 -> 2 outputs), pushes it into a (2r+1)-row register
// window, and emits the Z pass of the window's centre row (1 x STS.64).  The X-pass result never
// touches memory.
//
// Clamp-to-edge: the tile is loaded edge-replicated.  Interior tiles need nothing else (garbage
// creeps in r cells per iteration from the tile border and never reaches the interior).  Tiles that
// overlap the grid border re-replicate the new edge values into their out-of-grid cells after every
// iteration, which is exactly "an out-of-range neighbour reads the edge cell's current value".
#include "nz_common.cuh"

namespace nz {
namespace {

constexpr int FW = 128;       // tile width  (floats), halo included
constexpr int FH = 64;        // tile height (rows),   halo included
constexpr int HALO = 8;
constexpr int OW = FW - 2 * HALO;  // 112
constexpr int OH = FH - 2 * HALO;  // 80
constexpr int FTHREADS = 256;
constexpr int SMEM_BYTES = 2 * FW * FH * (int)sizeof(float);  // 98304 -> 2 CTAs / SM

template <int R>
struct TapsR {
    float k[2 * R + 1];
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// X pass of the two columns (c, c+1) of one tile row
template <int R, bool SCALE>
__device__ __forceinline__ float2 xrow(const float* __restrict__ in, int row, int c, const TapsR<R>& kx, float factor) {
    const float* p = in + row * FW + c;
    const float2 l = *reinterpret_cast<const float2*>(p - 2);
    const float2 m = *reinterpret_cast<const float2*>(p);
    const float2 r = *reinterpret_cast<const float2*>(p + 2);
    float v[8];  // columns c-2 .. c+3 (+2 more read for R>2)
    v[0] = l.x; v[1] = l.y; v[2] = m.x; v[3] = m.y; v[4] = r.x; v[5] = r.y;
    float o0 = 0.0f, o1 = 0.0f;
    if (R <= 2) {
#pragma unroll
        for (int k = -R; k <= R; k++) {
            o0 = fmaf(v[2 + k], kx.k[R + k], o0);
            o1 = fmaf(v[3 + k], kx.k[R + k], o1);
        }
    } else {
        // radius 3/4: two more float2 on each side
        const float2 ll = *reinterpret_cast<const float2*>(p - 4);
        const float2 rr = *reinterpret_cast<const float2*>(p + 4);
        float w[10] = {ll.x, ll.y, l.x, l.y, m.x, m.y, r.x, r.y, rr.x, rr.y};  // columns c-4 .. c+5
#pragma unroll
        for (int k = -R; k <= R; k++) {
            o0 = fmaf(w[4 + k], kx.k[R + k], o0);
            o1 = fmaf(w[5 + k], kx.k[R + k], o1);
        }
    }
    return SCALE ? make_float2(o0 * factor, o1 * factor) : make_float2(o0, o1);  // x * 1.0f == x: skipping is exact
}

template <int R, bool SCALE>
__global__ void __launch_bounds__(FTHREADS, 3)
sep_fused_kernel(const float* __restrict__ src, float* __restrict__ dst, int W, int H, int T, float factor,
                 TapsR<R> kx, TapsR<R> kz, int force_scalar) {
    extern __shared__ __align__(16) float smem[];
    float* bufA = smem;
    float* bufB = smem + FW * FH;
    constexpr int KS = 2 * R + 1;
    constexpr int CPAD = R <= 2 ? 2 : 4;  // columns a lane may not own at each tile side (reads stay in-tile)

    const int ox0 = blockIdx.x * OW, oz0 = blockIdx.y * OH;
    const int gx0 = ox0 - HALO, gz0 = oz0 - HALO;
    const bool boundary = gx0 < 0 || gz0 < 0 || gx0 + FW > W || gz0 + FH > H || force_scalar;

    // ---- load the tile, edge-replicated ----------------------------------------------------------
    if (!boundary) {
        for (int idx = threadIdx.x; idx < FW * FH / 4; idx += FTHREADS) {
            const int j = idx / (FW / 4), i = (idx % (FW / 4)) * 4;
            const float4 v = __ldg(reinterpret_cast<const float4*>(src + (size_t)(gz0 + j) * W + gx0 + i));
            *reinterpret_cast<float4*>(bufA + j * FW + i) = v;
        }
    } else {
        for (int idx = threadIdx.x; idx < FW * FH; idx += FTHREADS) {
            const int j = idx / FW, i = idx % FW;
            bufA[idx] = __ldg(src + (size_t)clampi(gz0 + j, 0, H - 1) * W + clampi(gx0 + i, 0, W - 1));
        }
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int seg = warp >> 1;
    const int c_own = ((warp & 1) * 32 + lane) * 2;
    const bool active = c_own >= CPAD && c_own + 1 < FW - CPAD;
    const int c = clampi(c_own, CPAD, FW - CPAD - 2);

    for (int t = 0; t < T; t++) {
        const float* in = (t & 1) ? bufB : bufA;
        float* out = (t & 1) ? bufA : bufB;
        const int z_lo = R * (t + 1), z_hi = FH - R * (t + 1);
        const int n = z_hi - z_lo;
        const int a = z_lo + (n * seg) / 4, b = z_lo + (n * (seg + 1)) / 4;

        // register window: slot s holds the X pass of tile row (a - R + s') with s' == s (mod KS)
        float2 xp[KS];
#pragma unroll
        for (int j = 0; j < 2 * R; j++) xp[j] = xrow<R, SCALE>(in, a - R + j, c, kx, factor);
        for (int z = a; z < b; z += KS) {
#pragma unroll
            for (int u = 0; u < KS; u++) {
                const int row = z + u;
                xp[(u + 2 * R) % KS] = xrow<R, SCALE>(in, min(row + R, FH - 1), c, kx, factor);
                float o0 = 0.0f, o1 = 0.0f;
#pragma unroll
                for (int k = R; k >= -R; k--) {
                    const float2 v = xp[(u + k + R) % KS];
                    o0 = fmaf(v.x, kz.k[R - k], o0);
                    o1 = fmaf(v.y, kz.k[R - k], o1);
                }
                if (active && row < b) *reinterpret_cast<float2*>(out + row * FW + c) = (SCALE ? make_float2(o0 * factor, o1 * factor) : make_float2(o0, o1));
            }
        }
        __syncthreads();
        if (boundary && t + 1 < T) {
            // re-replicate the new edge values into the out-of-grid cells of this tile
            for (int idx = threadIdx.x; idx < FW * FH; idx += FTHREADS) {
                const int j = idx / FW, i = idx % FW;
                const int gx = gx0 + i, gz = gz0 + j;
                const int cx = clampi(gx, 0, W - 1), cz = clampi(gz, 0, H - 1);
                if (cx != gx || cz != gz) {
                    const int si = cx - gx0, sj = cz - gz0;
                    if (si >= 0 && si < FW && sj >= 0 && sj < FH) out[idx] = out[sj * FW + si];
                }
            }
            __syncthreads();
        }
    }

    // ---- store the interior -----------------------------------------------------------------------
    const float* res = (T & 1) ? bufB : bufA;
    if (!boundary) {
        for (int idx = threadIdx.x; idx < OW * OH / 4; idx += FTHREADS) {
            const int j = idx / (OW / 4), i = (idx % (OW / 4)) * 4;
            const float4 v = *reinterpret_cast<const float4*>(res + (j + HALO) * FW + HALO + i);
            *reinterpret_cast<float4*>(dst + (size_t)(oz0 + j) * W + ox0 + i) = v;
        }
    } else {
        for (int idx = threadIdx.x; idx < OW * OH; idx += FTHREADS) {
            const int j = idx / OW, i = idx % OW;
            const int gx = ox0 + i, gz = oz0 + j;
            if (gx < W && gz < H) dst[(size_t)gz * W + gx] = res[(j + HALO) * FW + HALO + i];
        }
    }
}

template <int R>
int32_t launch_fused_r(float* d_data, float* d_tmp, int width, int rows, const float* kx, const float* kz, float factor,
                       int iterations, float** d_result, cudaStream_t s) {
    static DeviceOnce attr_set;  // benign race: the attribute is idempotent
    if (attr_set.need()) {
        NZ_CUDA(cudaFuncSetAttribute(sep_fused_kernel<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        NZ_CUDA(cudaFuncSetAttribute(sep_fused_kernel<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set.mark();
    }
    TapsR<R> tx, tz;
    for (int i = 0; i < 2 * R + 1; i++) {
        tx.k[i] = kx[i];
        tz.k[i] = kz[i];
    }
    // after T iterations the cells valid in a tile start at column CPAD + R*(T-1) and row R*T; both must be <= HALO
    const int cpad = R <= 2 ? 2 : 4;
    const int tmax = (HALO - cpad) / R + 1 < HALO / R ? (HALO - cpad) / R + 1 : HALO / R;
    // float4 global accesses need 16-byte aligned rows
    const int force_scalar = (width & 3) || ((uintptr_t)d_data & 15) || ((uintptr_t)d_tmp & 15);
    const int launches = (iterations + tmax - 1) / tmax;
    dim3 grid(cdiv(width, OW), cdiv(rows, OH));
    float *cur = d_data, *other = d_tmp;
    int left = iterations;
    for (int l = 0; l < launches; l++) {
        const int T = (left + (launches - l) - 1) / (launches - l);  // spread evenly
        if (factor == 1.0f)  // every Gauss table: the two `* factor` per cell-iteration are exact no-ops
            sep_fused_kernel<R, false><<<grid, FTHREADS, SMEM_BYTES, s>>>(cur, other, width, rows, T, factor, tx, tz, force_scalar);
        else
            sep_fused_kernel<R, true><<<grid, FTHREADS, SMEM_BYTES, s>>>(cur, other, width, rows, T, factor, tx, tz, force_scalar);
        NZ_LAUNCHED();
        left -= T;
        float* tmp = cur; cur = other; other = tmp;
    }
    if (d_result) {
        *d_result = cur;
    } else if (cur != d_data) {
        NZ_CUDA(cudaMemcpyAsync(d_data, cur, (size_t)width * rows * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    return NZ_OK;
}

}  // namespace

bool separable_fused_supported(int ksize) { return ksize >= 3 && ksize <= 9 && (ksize & 1); }

int32_t launch_separable_fused(float* d_data, float* d_tmp, int width, int rows, int ksize, const float* kx,
                               const float* kz, float factor, int iterations, float** d_result, cudaStream_t s) {
    switch (ksize) {
        case 3: return launch_fused_r<1>(d_data, d_tmp, width, rows, kx, kz, factor, iterations, d_result, s);
        case 5: return launch_fused_r<2>(d_data, d_tmp, width, rows, kx, kz, factor, iterations, d_result, s);
        case 7: return launch_fused_r<3>(d_data, d_tmp, width, rows, kx, kz, factor, iterations, d_result, s);
        case 9: return launch_fused_r<4>(d_data, d_tmp, width, rows, kx, kz, factor, iterations, d_result, s);
    }
    set_error("separable_fused: unsupported ksize %d", ksize);
    return NZ_E_UNSUPPORTED;
}

}  // namespace nz
