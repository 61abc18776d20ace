// nz_common.cuh — shared declarations of libnoize_b200.so (sm_100a only).
//
// Numerics contract: every kernel is compiled with --fmad=false, so a fused multiply-add exists
// exactly where the source says fmaf()/__fmaf_rn — the same places oracle/noize_oracle.cpp writes
// fmaf().  Division and sqrt are the IEEE-rounded forms (nvcc defaults -prec-div/-prec-sqrt=true),
// denormals are kept (-ftz=false).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include "../../include/noize_b200.h"

namespace nz {

// ---- error plumbing (thread-local message, integer status) ---------------------------------
void set_error(const char* fmt, ...);
int32_t cuda_fail(cudaError_t e, const char* what);
extern std::atomic<long long> g_launches;

#define NZ_CUDA(call)                                                   \
    do {                                                                \
        cudaError_t _e = (call);                                        \
        if (_e != cudaSuccess) return ::nz::cuda_fail(_e, #call);       \
    } while (0)

#define NZ_REQUIRE(cond, ...)                                           \
    do {                                                                \
        if (!(cond)) {                                                  \
            ::nz::set_error(__VA_ARGS__);                               \
            return NZ_E_INVALID;                                        \
        }                                                               \
    } while (0)

// count + check a kernel launch
#define NZ_LAUNCHED()                                                   \
    do {                                                                \
        ::nz::g_launches.fetch_add(1, std::memory_order_relaxed);       \
        cudaError_t _e = cudaGetLastError();                            \
        if (_e != cudaSuccess) return ::nz::cuda_fail(_e, "kernel launch"); \
    } while (0)

struct Taps {
    float k[NZ_MAX_KERNEL_WIDTH];
};

// ---- device-layer launchers (one per .cu file) ----------------------------------------------
struct FractalParams {
    int width, rows, z_first;
    int octaves;
    float posx, posz, noise_size;
    float start_amp, stepdown, detune_rate;
    float G;     // exp2f(-hurst), computed on the host with the same libm call the oracle uses
    float norm;  // CalcFractalNormValue
    int fast_hash;  // every lattice index of this launch is below 2^21 (simplex may use the magic-number residue)
    int fast_hash3d;  // the same for the domain-rotated 3-D bases (their rotated / skewed coordinates are up to 3x larger)
    int nonneg;       // every noise coordinate of this launch is >= 0 (tile origin, row offset, noise size and all octave frequencies)
};
int32_t fractal_params(FractalParams* p, int width, int rows, int z_first, int noise_type, float hurst, float start_amp,
                       float stepdown, float detune, int octaves, int xpos, int zpos, int noise_size);   // abi.cu
int32_t launch_fractal(float* d_dst, int noise_type, const FractalParams& p, cudaStream_t s);
bool fractal_pair_possible(int noise_type, const FractalParams& p);    // the packed-pair kernel computes this launch correctly
bool fractal_pair_supported(int noise_type, const FractalParams& p);   // ... and the window is large enough to prefer it
int32_t launch_fractal_pair(float* d_dst, int noise_type, const FractalParams& p, cudaStream_t s);

int32_t launch_separable(float* d_data, float* d_tmp, int width, int rows, int ksize, const float* kx,
                         const float* kz, float factor, int iterations, float** d_result, cudaStream_t s);
bool separable_walk_supported(int width, int ksize, const void* a, const void* b);
int32_t launch_separable_walk(float* d_data, float* d_tmp, int width, int rows, int ksize, const float* kx,
                              const float* kz, float factor, int iterations, float** d_result, cudaStream_t s);
bool separable_fused_supported(int ksize);
int32_t launch_separable_fused(float* d_data, float* d_tmp, int width, int rows, int ksize, const float* kx,
                               const float* kz, float factor, int iterations, float** d_result, cudaStream_t s);
int32_t launch_sobel2d(float* d_data, float* d_tmp, int width, int rows, int iterations, float** d_result,
                       cudaStream_t s);
int32_t launch_min_erosion(float* d_data, float* d_tmp, int width, int rows, int iterations, float** d_result,
                           cudaStream_t s);
size_t flowmap_scratch_bytes(int width, int rows, int iterations);
int32_t launch_flowmap(float* d_height, float* d_tmp, void* d_scratch, int width, int rows, int iterations,
                       float norm_min, float norm_max, float** d_result, cudaStream_t s);
size_t subtractive_flow_scratch_bytes(int width, int rows);
int32_t launch_subtractive_flow_erosion(float* d_height, void* d_scratch, int width, int rows, int erosive_iterations,
                                        float erosive_factor, float norm_min, float norm_max, cudaStream_t s);
bool flow_wave_supported(int width, int rows, int iterations, const void* a, const void* b);
int32_t launch_flow_wave(const float* d_height, float* d_out, int width, int rows, int iterations, float norm_min,
                         float norm_max, cudaStream_t s, const unsigned* gate = nullptr, unsigned epoch = 0,
                         unsigned* reruns = nullptr);
bool flow_walk_supported(int width, int rows, int iterations, const void* a, const void* b);
bool flow_walk_range_ok(float norm_min, float norm_max);
int32_t flow_walk_reruns(unsigned long long* count);
int32_t launch_flow_walk(const float* d_height, float* d_out, int width, int rows, int iterations, float norm_min,
                         float norm_max, cudaStream_t s);
bool flow_tile_supported(int width, int rows, int iterations, const void* a, const void* b);
int32_t launch_flow_tile(const float* d_height, float* d_out, int width, int rows, int iterations, float norm_min,
                         float norm_max, cudaStream_t s);
int32_t launch_flow_tile_cycle(const float* d_height, float* d_out, const float* d_fin, float* d_fout, size_t plane, int width, int rows,
                               int iterations, float norm_min, float norm_max, float factor, cudaStream_t s);
int32_t launch_mesh(int mesh_type, void* d_vtx, uint32_t* d_idx, int R, int inRes, float tile_height,
                    float tile_size, const float* d_heights, int h_row_first, int h_rows, int vz_begin, int vz_end,
                    cudaStream_t s);
float thermal_max_diff(float talus_deg, float height_ratio, int resolution);
int32_t launch_thermal_erosion(float* d_data, float* d_tmp, int res, float talus_deg, float increment, float height_ratio,
                               int iterations, float** d_result, cudaStream_t s);
int32_t launch_constant(float* d, size_t n, int op, float value, cudaStream_t s);
int32_t launch_reduce(float* d_left, const float* d_right, size_t n, int op, cudaStream_t s);
int32_t launch_curve(float* d, size_t n, const float* d_curve, int curve_size, cudaStream_t s);
int32_t launch_normalize(float* d, size_t n, float vmin, float range, cudaStream_t s);
// one pass applying up to PW_CHAIN_MAX element-wise steps in order (pointwise_kernels.cu: ChainF)
enum { PW_MUL = 0, PW_BINARIZE = 1, PW_NORMALIZE = 2, PW_CURVE = 3 };
constexpr int PW_CHAIN_MAX = 8;
struct PointwiseOp {
    int kind;
    float a, b;          // MUL/BINARIZE: a = constant; NORMALIZE: a = min, b = range; CURVE: a = (float)curve size
    const float* lut;    // CURVE: device pointer of the curve samples
};
int32_t launch_pointwise_chain(float* d, size_t n, const PointwiseOp* ops, int count, cudaStream_t s);
int32_t launch_crop(const float* d_in, int in_res, float* d_out, int out_res, int offset, cudaStream_t s);
size_t map_range_scratch_bytes();
int32_t launch_map_range(const float* d, size_t n, float lim_min, float lim_max, float* d_res3, void* d_scratch, cudaStream_t s);
int32_t launch_fma_peak(float* d_sink, int grid, int iters, double* flops, cudaStream_t s);
int32_t launch_gather_strided(float* d_dst, const unsigned char* d_src, int stride_bytes, size_t n, cudaStream_t s);

// device-buffer pool (abi.cu): blocks of the CURRENT device; dev_free expects every use of the block to be complete
int32_t dev_alloc(void** p, size_t bytes);
void dev_free(void* p);

// side stream of the calling thread on the current device (aux_stream.cu)
int32_t aux_fork(cudaStream_t main, cudaStream_t* aux);
int32_t aux_join(cudaStream_t main);
// joins the side stream when a launcher leaves early (an error return between aux_fork and its aux_join)
struct AuxJoinGuard {
    cudaStream_t main;
    bool armed = false;
    explicit AuxJoinGuard(cudaStream_t m) : main(m) {}
    ~AuxJoinGuard() { if (armed) aux_join(main); }
    int32_t join() {
        if (!armed) return 0;
        armed = false;
        return aux_join(main);
    }
};

// host-side table helpers (tables.cpp part of abi.cu)
void gauss_table(double sigma, int width, float* out);
int32_t kernel_filter_table(int filter, float* kx, float* kz, int* ksize, float* factor);

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// One-time setup that is PER DEVICE (cudaFuncSetAttribute applies to the current device only): need() is true until
// mark() has been called on the current device.  The setups guarded by it are idempotent, so a race repeats one, no more.
struct DeviceOnce {
    std::atomic<unsigned long long> done{0};
    static unsigned long long bit() {
        int dev = 0;
        cudaGetDevice(&dev);
        return 1ull << (dev & 63);
    }
    bool need() const { return !(done.load(std::memory_order_acquire) & bit()); }
    void mark() { done.fetch_or(bit(), std::memory_order_release); }
};
// multiprocessor count of the CURRENT device (cached per device)
int sm_count();

// Which edges of the window a stencil launch works on are edges of the GRID (bit 0: first row, bit 1: last row).  The
// default is both: reads beyond them clamp to the edge row (Pipeline/Tiles/TileData.cs:72-77).  A row band in the middle
// of a larger grid (bands.cu) has neither: its first and last rows are ghost rows whose results nobody keeps, so the
// register-walk kernels let their clamp-free interior launch run over them instead of handing 2 x 32 (filter) or 2 x 64
// (flow map) rows of every strip to the slower border launch.  A thread-local hint, set around one stage call.
int grid_edges();
struct GridEdgesScope {
    int prev;
    explicit GridEdgesScope(int edges);
    ~GridEdgesScope();
};

}  // namespace nz
