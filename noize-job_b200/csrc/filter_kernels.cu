// filter_kernels.cu — separable kernel filter, Sobel3_2D reduce and value (min) erosion (hot loop 2).
//
// Replaces GenericKernelJob<KernelTileMutation<KernelSample{X,Z}Operator>, RWTileData> and the
// KernelMin{X,Z}Operator variants (Filter/Kernel/KernelJob.cs:18-54,165-185,317-347,
// Filter/Kernel/KernelOperators.cs:18-117), with clamp-to-edge reads (Pipeline/Tiles/TileData.cs:72-77).
// The reference's per-pass copy-back (FlushWriteSlice, TileData.cs:16-40) is result-neutral and is
// replaced by ping-ponging between the two HBM buffers.
//
// Arithmetic order (must match oracle/noize_oracle.cpp generic_kernel_job):
//   X pass: total = 0; for k = -r..r : total = fma(src(x+k,z), K[r+k], total);  out = total*factor
//   Z pass: total = 0; for k = r..-r : total = fma(src(x,z+k), K[r-k], total);  out = total*factor
#include "nz_common.cuh"

namespace nz {
namespace {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// ---------------------------------------------------------------------------------------------
// Generic one-pass kernels (any odd ksize <= 25): one thread per cell, 4 cells per thread along x
// for the Z pass so loads/stores are float4 when the row pitch allows.  HBM/L2 bound.
// ---------------------------------------------------------------------------------------------
constexpr int TX = 128;

__global__ void __launch_bounds__(TX) sep_x_kernel(const float* __restrict__ src, float* __restrict__ dst, int width,
                                                   int rows, int ksize, float factor, Taps taps) {
    __shared__ float line[TX + NZ_MAX_KERNEL_WIDTH - 1];
    const int z = blockIdx.y;
    const int x0 = blockIdx.x * TX;
    const int r = (ksize - 1) >> 1;
    const float* row = src + (size_t)z * width;
    for (int i = threadIdx.x; i < TX + 2 * r; i += TX) line[i] = __ldg(row + clampi(x0 + i - r, 0, width - 1));
    __syncthreads();
    const int x = x0 + threadIdx.x;
    if (x >= width) return;
    float total = 0.0f;
    for (int k = 0; k < ksize; k++) total = fmaf(line[threadIdx.x + k], taps.k[k], total);
    dst[(size_t)z * width + x] = total * factor;
}

__global__ void __launch_bounds__(TX) sep_z_kernel(const float* __restrict__ src, float* __restrict__ dst, int width,
                                                   int rows, int ksize, float factor, Taps taps) {
    const int z = blockIdx.y;
    const int x = blockIdx.x * TX + threadIdx.x;
    if (x >= width) return;
    const int r = (ksize - 1) >> 1;
    float total = 0.0f;
    for (int k = r; k >= -r; k--) {
        const int zi = clampi(z + k, 0, rows - 1);
        total = fmaf(__ldg(src + (size_t)zi * width + x), taps.k[r - k], total);
    }
    dst[(size_t)z * width + x] = total * factor;
}

// ---------------------------------------------------------------------------------------------
// Sobel3_2D (SeparableKernelFilter.ScheduleReduce<RootSumSquaresTiles>, KernelJob.cs:187-215;
// RootSumSquaresTiles, Filter/Operators/SimpleMutation.cs:148-171), intended semantics:
//   A = Z{1,2,1}( X{-1,0,1}(src) ),  B = Z{1,0,-1}( X{1,2,1}(src) ),  out = sqrt(A*A + B*B)
// fused into one 3x3 pass in the exact order of the two separable branches.
// ---------------------------------------------------------------------------------------------
// the two X passes of one cell (taps HX {-1,0,1} and VX {1,2,1}, accumulated from 0 in tap order, factor 1)
__device__ __forceinline__ void sobel_x(float l, float c, float rr, float& a, float& b) {
    float ta = fmaf(l, -1.0f, 0.0f);
    ta = fmaf(c, 0.0f, ta);
    ta = fmaf(rr, 1.0f, ta);
    a = ta * 1.0f;
    float tb = fmaf(l, 1.0f, 0.0f);
    tb = fmaf(c, 2.0f, tb);
    tb = fmaf(rr, 1.0f, tb);
    b = tb * 1.0f;
}
// the two Z passes (descending k: taps pair K[0] with z+1, K[1] with z, K[2] with z-1) and the root-sum-squares
__device__ __forceinline__ float sobel_z(float a0, float a1, float a2, float b0, float b1, float b2) {
    float A = fmaf(a2, 1.0f, 0.0f);
    A = fmaf(a1, 2.0f, A);
    A = fmaf(a0, 1.0f, A);
    float B = fmaf(b2, 1.0f, 0.0f);
    B = fmaf(b1, 0.0f, B);
    B = fmaf(b0, -1.0f, B);
    return sqrtf(fmaf(B, B, A * A));
}

__global__ void __launch_bounds__(TX) sobel2d_kernel(const float* __restrict__ src, float* __restrict__ dst, int width,
                                                     int rows) {
    const int z = blockIdx.y;
    const int x = blockIdx.x * TX + threadIdx.x;
    if (x >= width) return;
    const int xl = max(x - 1, 0), xr = min(x + 1, width - 1);
    float ax[3], bx[3];  // X-pass results of rows z-1, z, z+1
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const float* row = src + (size_t)clampi(z + j - 1, 0, rows - 1) * width;
        sobel_x(__ldg(row + xl), __ldg(row + x), __ldg(row + xr), ax[j], bx[j]);
    }
    dst[(size_t)z * width + x] = sobel_z(ax[0], ax[1], ax[2], bx[0], bx[1], bx[2]);
}

// Row walk: a warp owns a 128-column strip (lane = 4 adjacent columns, one float4 per row) and streams down a chunk of
// rows keeping the X-pass results of the last three rows in registers: 4 B read + 4 B written per cell (+ 2 rows per
// chunk), west/east neighbours by shuffle (the strip's outer neighbours are one scalar load in lanes 0 and 31).
constexpr int SW_WARPS = 4, SW_ZC = 64, SW_PF = 3;
__global__ void __launch_bounds__(SW_WARPS * 32) sobel2d_walk_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                                     int width, int rows) {
    const int lane = threadIdx.x & 31;
    const int x0 = (blockIdx.x * SW_WARPS + (threadIdx.x >> 5)) * 128;
    if (x0 >= width) return;
    const int gx = x0 + 4 * lane;
    const bool in = gx < width;                           // width % 4 == 0: a lane's columns are all in or all out
    const int zc0 = blockIdx.y * SW_ZC, zc1 = min(zc0 + SW_ZC, rows);
    const int ex = lane == 0 ? max(x0 - 1, 0) : min(x0 + 128, width - 1);   // outer neighbour column (lanes 0 / 31)
    const bool has_e = lane == 0 || lane == 31;

    auto load = [&](int r, float4& v, float& e) {
        const float* row = src + (size_t)clampi(r, 0, rows - 1) * width;
        v = in ? __ldg(reinterpret_cast<const float4*>(row + gx)) : make_float4(0.f, 0.f, 0.f, 0.f);
        e = has_e ? __ldg(row + ex) : 0.0f;
    };
    float4 pv[SW_PF];
    float pe[SW_PF];
#pragma unroll
    for (int j = 0; j < SW_PF; j++) load(zc0 - 1 + j, pv[j], pe[j]);
    float ax[3][4], bx[3][4];
    // rows r = zc0-1 .. zc1 ; row r lands in window slot (r - (zc0-1)) % 3 ; after row r the row r-1 is complete
    for (int base = zc0 - 1; base <= zc1; base += 3) {
#pragma unroll
        for (int u = 0; u < 3; u++) {
            const int r = base + u;
            if (r <= zc1) {
                const float4 v = pv[u];
                const float e = pe[u];
                load(r + SW_PF, pv[u], pe[u]);           // SW_PF == 3: the slot just consumed
                float L = __shfl_up_sync(0xffffffffu, v.w, 1), R = __shfl_down_sync(0xffffffffu, v.x, 1);
                if (lane == 0) L = e;
                if (lane == 31) R = e;
                if (gx == 0) L = v.x;
                if (gx + 4 >= width) R = v.w;
                sobel_x(L, v.x, v.y, ax[u][0], bx[u][0]);
                sobel_x(v.x, v.y, v.z, ax[u][1], bx[u][1]);
                sobel_x(v.y, v.z, v.w, ax[u][2], bx[u][2]);
                sobel_x(v.z, v.w, R, ax[u][3], bx[u][3]);
                const int z = r - 1;                      // rows z-1, z, z+1 sit in slots (u+1)%3, (u+2)%3, u
                if (z >= zc0 && in) {
                    float o[4];
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        o[q] = sobel_z(ax[(u + 1) % 3][q], ax[(u + 2) % 3][q], ax[u][q], bx[(u + 1) % 3][q], bx[(u + 2) % 3][q], bx[u][q]);
                    *reinterpret_cast<float4*>(dst + (size_t)z * width + gx) = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Value erosion: N calls of (X pass, Z pass) with taps {-1,0} == min over the trailing window
// [x-N..x] x [z-N..z] with clamp at 0 (exact: min is associative/commutative).  Two passes:
// X window then Z window.
// ---------------------------------------------------------------------------------------------
// One fused pass: dst(x,z) = min over the trailing (n+1) x (n+1) window, staged through shared memory
// (X window first, then Z window): 4 B read + 4 B written per cell instead of two HBM round trips.
constexpr int MW = 128, MH = 32, MIN_FUSED_MAX_N = 32, MTHREADS = 256;

__global__ void __launch_bounds__(MTHREADS) min_window_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                              int width, int rows, int n) {
    extern __shared__ float msm[];
    const int aw = MW + n, ah = MH + n;   // staged input: n extra columns to the left, n extra rows above
    float* A = msm;                       // ah x aw
    float* B = msm + ah * aw;             // ah x MW : X-window minima
    const int x0 = blockIdx.x * MW, z0 = blockIdx.y * MH;
    for (int idx = threadIdx.x; idx < ah * aw; idx += MTHREADS) {
        const int j = idx / aw, i = idx - j * aw;
        A[idx] = __ldg(src + (size_t)clampi(z0 - n + j, 0, rows - 1) * width + clampi(x0 - n + i, 0, width - 1));
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < ah * MW; idx += MTHREADS) {
        const int j = idx / MW, i = idx - j * MW;
        const float* a = A + j * aw + i;
        float m = a[0];
        for (int k = 1; k <= n; k++) m = fminf(m, a[k]);
        B[idx] = m;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < MH * MW; idx += MTHREADS) {
        const int j = idx / MW, i = idx - j * MW;
        const int x = x0 + i, z = z0 + j;
        if (x < width && z < rows) {
            const float* b = B + j * MW + i;
            float m = b[0];
            for (int k = 1; k <= n; k++) m = fminf(m, b[k * MW]);
            dst[(size_t)z * width + x] = m;
        }
    }
}

// Register/shuffle variant for small windows (N <= 8, the "Value Erosion x5" of the README is N = 5).  A lane owns one
// column and walks down a chunk of rows keeping the last N+1 loaded values in registers (Z window), then takes the
// trailing X window across lanes by log-step doubling with shuffles.  A warp's 32 lanes overlap the previous warp's by
// N columns, so no shared memory and no block synchronisation is needed; loads and stores are row-contiguous.
constexpr int MWALK_ROWS = 128, MWALK_THREADS = 256, MWALK_MAX_N = 8;

template <int N>
__global__ void __launch_bounds__(MWALK_THREADS) min_walk_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                                 int width, int rows) {
    constexpr int USE = 32 - N;                                   // columns a warp actually produces
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wx0 = (blockIdx.x * (MWALK_THREADS / 32) + warp) * USE;  // first produced column of this warp
    const int x = wx0 - N + lane;                                  // this lane's column (may be < 0: clamp)
    if (wx0 >= width) return;
    const int xc = clampi(x, 0, width - 1);
    const int z0 = blockIdx.y * MWALK_ROWS, z1 = min(z0 + MWALK_ROWS, rows);
    float win[N + 1];   // win[j] = value of row (z - N + j), clamped at row 0
#pragma unroll
    for (int j = 0; j < N; j++) win[j + 1] = __ldg(src + (size_t)max(z0 - N + j, 0) * width + xc);
    for (int z = z0; z < z1; z++) {
#pragma unroll
        for (int j = 0; j < N; j++) win[j] = win[j + 1];
        win[N] = __ldg(src + (size_t)z * width + xc);
        float m = win[0];
#pragma unroll
        for (int j = 1; j <= N; j++) m = fminf(m, win[j]);
        // trailing window of N+1 lanes: doubling, then one shifted combine
        int span = 1;   // m currently covers `span` trailing lanes
#pragma unroll
        for (int k = 1; 2 * k <= N + 1; k *= 2) {
            m = fminf(m, __shfl_up_sync(0xffffffffu, m, k));
            span = 2 * k;
        }
        if (span < N + 1) m = fminf(m, __shfl_up_sync(0xffffffffu, m, N + 1 - span));
        if (lane >= N && x < width) dst[(size_t)z * width + x] = m;
    }
}

// fallback for n > MIN_FUSED_MAX_N: separable X window then Z window through HBM
__global__ void __launch_bounds__(TX) min_x_kernel(const float* __restrict__ src, float* __restrict__ dst, int width,
                                                   int rows, int n) {
    const int z = blockIdx.y;
    const int x = blockIdx.x * TX + threadIdx.x;
    if (x >= width) return;
    const float* row = src + (size_t)z * width;
    float m = 3.402823466e+38f;
    for (int k = min(n, x); k >= 0; k--) m = fminf(m, __ldg(row + x - k));
    dst[(size_t)z * width + x] = m;
}

__global__ void __launch_bounds__(TX) min_z_kernel(const float* __restrict__ src, float* __restrict__ dst, int width,
                                                   int rows, int n) {
    const int z = blockIdx.y;
    const int x = blockIdx.x * TX + threadIdx.x;
    if (x >= width) return;
    float m = 3.402823466e+38f;
    for (int k = min(n, z); k >= 0; k--) m = fminf(m, __ldg(src + (size_t)(z - k) * width + x));
    dst[(size_t)z * width + x] = m;
}

__global__ void gather_strided_kernel(float* __restrict__ dst, const unsigned char* __restrict__ src, int stride, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = *reinterpret_cast<const float*>(src + i * (size_t)stride);
}

}  // namespace

int32_t launch_separable(float* d_data, float* d_tmp, int width, int rows, int ksize, const float* kx,
                         const float* kz, float factor, int iterations, float** d_result, cudaStream_t s) {
    NZ_REQUIRE(d_data && d_tmp, "separable: null device buffer");
    NZ_REQUIRE(width > 0 && rows > 0 && rows <= 65535, "separable: bad grid %d x %d", width, rows);
    NZ_REQUIRE(ksize >= 1 && (ksize & 1) && ksize <= NZ_MAX_KERNEL_WIDTH, "separable: ksize %d must be odd and <= %d",
               ksize, NZ_MAX_KERNEL_WIDTH);
    NZ_REQUIRE(iterations >= 0 && kx && kz, "separable: bad iterations/taps");
    // NZ_SEP_PATH=walk|fused|generic forces one path (the tests compare all three bit for bit)
    const char* force = getenv("NZ_SEP_PATH");
    const bool want_generic = force && force[0] == 'g', want_fused = force && force[0] == 'f';
    // The register walk streams long columns per warp and needs ~10^7 cells to fill the GPU; below that the
    // shared-memory tile kernel (one CTA per 128 x 96 tile) has the shorter critical path.  Measured Gauss5 x17,
    // walk / fused ms: 1024^2 0.290 / 0.102, 2048^2 0.264 / 0.166, 4096^2 0.432 / 0.485.  All paths give the same bits.
    const bool small = (long long)width * rows < (8ll << 20) && separable_fused_supported(ksize) && !(force && force[0] == 'w');
    if (!want_generic && !want_fused && !small && iterations > 0 && separable_walk_supported(width, ksize, d_data, d_tmp))
        return launch_separable_walk(d_data, d_tmp, width, rows, ksize, kx, kz, factor, iterations, d_result, s);
    if (!want_generic && separable_fused_supported(ksize) && iterations > 0)
        return launch_separable_fused(d_data, d_tmp, width, rows, ksize, kx, kz, factor, iterations, d_result, s);
    Taps tx, tz;
    for (int i = 0; i < NZ_MAX_KERNEL_WIDTH; i++) {
        tx.k[i] = i < ksize ? kx[i] : 0.0f;
        tz.k[i] = i < ksize ? kz[i] : 0.0f;
    }
    dim3 grid(cdiv(width, TX), rows);
    for (int it = 0; it < iterations; it++) {
        sep_x_kernel<<<grid, TX, 0, s>>>(d_data, d_tmp, width, rows, ksize, factor, tx);
        NZ_LAUNCHED();
        sep_z_kernel<<<grid, TX, 0, s>>>(d_tmp, d_data, width, rows, ksize, factor, tz);
        NZ_LAUNCHED();
    }
    if (d_result) *d_result = d_data;
    return NZ_OK;
}

int32_t launch_sobel2d(float* d_data, float* d_tmp, int width, int rows, int iterations, float** d_result,
                       cudaStream_t s) {
    NZ_REQUIRE(d_data && d_tmp, "sobel2d: null device buffer");
    NZ_REQUIRE(width > 0 && rows > 0 && rows <= 65535 && iterations >= 0, "sobel2d: bad arguments");
    dim3 grid(cdiv(width, TX), rows);
    float *a = d_data, *b = d_tmp;
    // NZ_SOBEL_PATH=plain forces the one-thread-per-cell kernel (the tests compare the two bit for bit)
    const char* sp = getenv("NZ_SOBEL_PATH");
    // the walk needs ~600 warps of 64-row chunks to fill the machine: below 4M cells the per-cell kernel's parallelism wins
    // (NZ_SOBEL_PATH=walk forces the walk, for the tests)
    const bool big = (long long)width * rows >= (1LL << 22) || (sp && sp[0] == 'w');
    const bool walk = big && !(sp && sp[0] == 'p') && (width & 3) == 0 && (((uintptr_t)d_data | (uintptr_t)d_tmp) & 15) == 0;
    for (int it = 0; it < iterations; it++) {
        if (walk) {
            dim3 wgrid(cdiv(cdiv(width, 128), SW_WARPS), cdiv(rows, SW_ZC));
            sobel2d_walk_kernel<<<wgrid, SW_WARPS * 32, 0, s>>>(a, b, width, rows);
        } else
            sobel2d_kernel<<<grid, TX, 0, s>>>(a, b, width, rows);
        NZ_LAUNCHED();
        float* t = a; a = b; b = t;
    }
    if (d_result) {
        *d_result = a;
    } else if (a != d_data) {
        NZ_CUDA(cudaMemcpyAsync(d_data, a, (size_t)width * rows * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    return NZ_OK;
}

int32_t launch_min_erosion(float* d_data, float* d_tmp, int width, int rows, int iterations, float** d_result,
                           cudaStream_t s) {
    NZ_REQUIRE(d_data && d_tmp, "min_erosion: null device buffer");
    NZ_REQUIRE(width > 0 && rows > 0 && rows <= 65535 && iterations >= 0, "min_erosion: bad arguments");
    if (iterations > MIN_FUSED_MAX_N) NZ_REQUIRE(rows <= 65535, "min_erosion: too many rows");
    if (iterations > 0 && iterations <= MWALK_MAX_N) {
        dim3 grid(cdiv(width, (MWALK_THREADS / 32) * (32 - iterations)), cdiv(rows, MWALK_ROWS));
        switch (iterations) {
            case 1: min_walk_kernel<1><<<grid, MWALK_THREADS, 0, s>>>(d_data, d_tmp, width, rows); break;
            case 2: min_walk_kernel<2><<<grid, MWALK_THREADS, 0, s>>>(d_data, d_tmp, width, rows); break;
            case 3: min_walk_kernel<3><<<grid, MWALK_THREADS, 0, s>>>(d_data, d_tmp, width, rows); break;
            case 4: min_walk_kernel<4><<<grid, MWALK_THREADS, 0, s>>>(d_data, d_tmp, width, rows); break;
            case 5: min_walk_kernel<5><<<grid, MWALK_THREADS, 0, s>>>(d_data, d_tmp, width, rows); break;
            case 6: min_walk_kernel<6><<<grid, MWALK_THREADS, 0, s>>>(d_data, d_tmp, width, rows); break;
            case 7: min_walk_kernel<7><<<grid, MWALK_THREADS, 0, s>>>(d_data, d_tmp, width, rows); break;
            default: min_walk_kernel<8><<<grid, MWALK_THREADS, 0, s>>>(d_data, d_tmp, width, rows); break;
        }
        NZ_LAUNCHED();
        if (d_result) {
            *d_result = d_tmp;
        } else {
            NZ_CUDA(cudaMemcpyAsync(d_data, d_tmp, (size_t)width * rows * sizeof(float), cudaMemcpyDeviceToDevice, s));
        }
        return NZ_OK;
    }
    if (iterations > 0 && iterations <= MIN_FUSED_MAX_N) {
        const int n = iterations;
        const size_t smem = ((size_t)(MH + n) * (MW + n) + (size_t)(MH + n) * MW) * sizeof(float);
        static DeviceOnce attr_set;
        if (attr_set.need()) {
            NZ_CUDA(cudaFuncSetAttribute(min_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(((MH + MIN_FUSED_MAX_N) * (MW + MIN_FUSED_MAX_N) + (MH + MIN_FUSED_MAX_N) * MW) * sizeof(float))));
            attr_set.mark();
        }
        dim3 grid(cdiv(width, MW), cdiv(rows, MH));
        min_window_kernel<<<grid, MTHREADS, smem, s>>>(d_data, d_tmp, width, rows, n);
        NZ_LAUNCHED();
        if (d_result) {
            *d_result = d_tmp;
        } else {
            NZ_CUDA(cudaMemcpyAsync(d_data, d_tmp, (size_t)width * rows * sizeof(float), cudaMemcpyDeviceToDevice, s));
        }
        return NZ_OK;
    }
    if (iterations > 0) {
        dim3 grid(cdiv(width, TX), rows);
        min_x_kernel<<<grid, TX, 0, s>>>(d_data, d_tmp, width, rows, iterations);
        NZ_LAUNCHED();
        min_z_kernel<<<grid, TX, 0, s>>>(d_tmp, d_data, width, rows, iterations);
        NZ_LAUNCHED();
    }
    if (d_result) *d_result = d_data;
    return NZ_OK;
}

int32_t launch_gather_strided(float* d_dst, const unsigned char* d_src, int stride_bytes, size_t n, cudaStream_t s) {
    gather_strided_kernel<<<cdiv((long long)n, 256), 256, 0, s>>>(d_dst, d_src, stride_bytes, n);
    NZ_LAUNCHED();
    return NZ_OK;
}

}  // namespace nz
