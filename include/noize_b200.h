/*
 * noize_b200.h — C ABI of the B200-native heightmap hot path (libnoize_b200.so).
 *
 * This is the drop-in boundary for xshazwar/noize-job's per-cell pipeline.  Every entry
 * point below replaces one of the reference's *static job delegates* (the table entries a
 * PipelineStage binds), so a C# `[DllImport("noize_b200")]` stub with the same argument
 * order can be swapped into the stage's delegate table.
 *
 * Reference paths are relative to the upstream repo (xshazwar/noize-job).
 *
 * Conventions
 *   - every function returns int32_t: NZ_OK (0) or a negative NZ_E* code; nothing throws or
 *     aborts.  `nz_last_error()` returns a thread-local message for the last failure
 *     (reference convention: exceptions at schedule time, Pipeline/Stage/PipelineStage.cs:37).
 *   - grids are row-major float32, idx = z*width + x   (Pipeline/Tiles/TileData.cs:135-138)
 *   - reads outside the grid clamp to the edge cell     (Pipeline/Tiles/TileData.cs:72-77)
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     NZ_E_CUDA.
 *
 * Two layers:
 *   nz_*      host layer   — takes NativeSlice<float>-shaped HOST buffers (ptr,stride,length),
 *                            moves them to the GPU, runs the stage, moves the result back.
 *                            Between nz_pipeline_begin/nz_pipeline_end the device copy stays
 *                            resident (keyed by host pointer) and the D2H is deferred.
 *   nz_dev_*  device layer — operates on caller-owned DEVICE buffers on a caller stream, on
 *                            rectangular (width x rows) grids so one rank can own a row band
 *                            of a larger heightmap.  No synchronisation, no copies.
 */
#ifndef NOIZE_B200_H
#define NOIZE_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NZ_API __attribute__((visibility("default")))

/* ---- status codes ------------------------------------------------------------------ */
#define NZ_OK              0
#define NZ_E_INVALID      -1   /* bad argument (null slice, length != res*res, bad enum, ...) */
#define NZ_E_CUDA         -2   /* CUDA runtime error or no device */
#define NZ_E_NOMEM        -3   /* device/host allocation failed */
#define NZ_E_STATE        -4   /* call not legal in the current pipeline state */
#define NZ_E_UNSUPPORTED  -5   /* parameter combination not implemented */

/* ---- enums: numeric values ARE the reference's C# enum order ------------------------- */

/* NoiseStage.FractalNoise, Noise/NoiseStage.cs:15-24 (index into the delegate table :26-35) */
typedef enum {
    NZ_NOISE_SIN = 0,
    NZ_NOISE_PERLIN = 1,
    NZ_NOISE_PERIODIC_PERLIN = 2,
    NZ_NOISE_SIMPLEX = 3,
    NZ_NOISE_ROTATED_SIMPLEX = 4,
    NZ_NOISE_CELLULAR = 5,
    NZ_NOISE_DOMAIN_ROTATED_PERLIN = 6,
    NZ_NOISE_DOMAIN_ROTATED_SIMPLEX = 7,
    NZ_NOISE__COUNT = 8
} nz_noise_type;

/* KernelFilterType, Filter/Kernel/KernelJob.cs:79-94 */
typedef enum {
    NZ_FILTER_GAUSS9_S1 = 0,
    NZ_FILTER_GAUSS7_S1 = 1,
    NZ_FILTER_GAUSS5_S1 = 2,
    NZ_FILTER_GAUSS3_S1 = 3,
    NZ_FILTER_GAUSS9_S2 = 4,
    NZ_FILTER_GAUSS7_S2 = 5,
    NZ_FILTER_GAUSS5_S2 = 6,
    NZ_FILTER_GAUSS3_S2 = 7,
    NZ_FILTER_SMOOTH3 = 8,
    NZ_FILTER_SOBEL3_HORIZONTAL = 9,
    NZ_FILTER_SOBEL3_VERTICAL = 10,
    NZ_FILTER_SOBEL3_2D = 11,
    NZ_FILTER_PREWITT3_HORIZONTAL = 12,
    NZ_FILTER_PREWITT3_VERTICAL = 13,
    NZ_FILTER__COUNT = 14
} nz_kernel_filter_type;

/* GaussSigma, Filter/Kernel/Blur/BlurKernels.cs:8-25: sigma = 0.5 * (index + 1) */
typedef enum {
    NZ_SIGMA_0D50 = 0, NZ_SIGMA_1D00, NZ_SIGMA_1D50, NZ_SIGMA_2D00, NZ_SIGMA_2D50, NZ_SIGMA_3D00,
    NZ_SIGMA_3D50, NZ_SIGMA_4D00, NZ_SIGMA_4D50, NZ_SIGMA_5D00, NZ_SIGMA_5D50, NZ_SIGMA_6D00,
    NZ_SIGMA_6D50, NZ_SIGMA_7D00, NZ_SIGMA_7D50, NZ_SIGMA_8D00, NZ_SIGMA__COUNT
} nz_gauss_sigma;

/* MeshType, Mesh/Stage/MeshTileStage.cs:22-25 (index into the delegate table :30-33) */
typedef enum {
    NZ_MESH_SQUARE_GRID = 0,
    NZ_MESH_OVERSHOOT_SQUARE_GRID = 1,
    NZ_MESH__COUNT = 2
} nz_mesh_type;

/* ConstantStage.ConstantOperationType, Filter/ConstantStage.cs:15-18 (index into the delegate table :20-23) */
typedef enum {
    NZ_CONSTANT_MULTIPLY = 0,
    NZ_CONSTANT_BINARIZE = 1,
    NZ_CONSTANT__COUNT = 2
} nz_constant_op;

/* ReductionType, Filter/Reduce/ReduceStage.cs:12-18 (index into the delegate table :22-28) */
typedef enum {
    NZ_REDUCE_SUBTRACT = 0,
    NZ_REDUCE_MULTIPLY = 1,
    NZ_REDUCE_ROOTSUMSQUARES = 2,
    NZ_REDUCE_MAX = 3,
    NZ_REDUCE_MIN = 4,
    NZ_REDUCE__COUNT = 5
} nz_reduce_op;

#define NZ_MAX_KERNEL_WIDTH 25          /* BlurHelper.max_width, BlurKernels.cs:29 */
#define NZ_MESH_VERTEX_BYTES 48         /* PositionStream32.Stream0, Mesh/Streams/PositionStream.cs:77-82 */

/* NativeSlice<float> as Unity lays it out: base pointer, byte stride between elements, element
 * count.  stride_bytes may be != 4 (Scripts/Editor/VisualizePipeline.cs:141 passes one channel of
 * an RGBAFloat texture, stride 16). */
typedef struct {
    float*  ptr;
    int32_t stride_bytes;
    int32_t length;
} nz_slice_f32;

/* One vertex as the GPU writes it: {float3 position; float3 normal; float4 tangent; float2 uv}
 * (Mesh/Job/Vertex.cs:5-10, StructLayout.Sequential in PositionStream.cs:77-82). */
typedef struct {
    float position[3];
    float normal[3];
    float tangent[4];
    float uv[2];
} nz_mesh_vertex;

/* per-call timing of the LAST host-layer call made by the calling thread (milliseconds) */
typedef struct {
    float ms_h2d;
    float ms_kernel;
    float ms_d2h;
    int32_t kernel_launches;
} nz_timing;

/* ---- lifecycle / introspection ------------------------------------------------------ */

/* Select the CUDA device(s) this process drives; devices==NULL or n==0 -> device 0.  devices[0] is the device of the
 * host layer; with nz_set_bands(k) large grids are split over devices[0..k).  Optional: the first compute call
 * initialises device 0 lazily.  Safe to call again while no residency scope is open (NZ_E_STATE otherwise); a failed
 * call keeps the previous configuration.  The device layer (nz_dev_*) never changes the caller's current device. */
NZ_API int32_t nz_init(const int32_t* devices, int32_t n);
NZ_API int32_t nz_shutdown(void);
NZ_API const char* nz_last_error(void);
NZ_API const char* nz_version(void);
/* number of kernels this library has launched since nz_init (monotonic, all threads) */
NZ_API int64_t nz_kernel_launch_count(void);
NZ_API int32_t nz_last_timing(nz_timing* out);
/* Test hook (fault injection): after `skip` more device allocations of the library, the next `count` fail with NZ_E_NOMEM. */
NZ_API int32_t nz_test_fail_allocs(int32_t skip, int32_t count);

/* ---- host-side helpers that need no GPU (host logic of the stages) ------------------ */

/* FractalJob.CalcFractalNormValue, Noise/Fractal/Fractal.cs:31-40 (startingAmplitude ignored). */
NZ_API float nz_fractal_norm_value(float hurst, int32_t octaves);
/* GaussianKernel.GetKernel, Filter/Kernel/Blur/BlurKernels.cs:42-58 (+ tables :59-316):
 * normalised sampled Gaussian of odd `width` (after BlurHelper.limitWidth :30-36); writes
 * `*width_out` floats to `out` (capacity >= NZ_MAX_KERNEL_WIDTH). */
NZ_API int32_t nz_gauss_kernel(int32_t sigma, int32_t width, float* out, int32_t* width_out);
/* BlurHelper.limitWidth, BlurKernels.cs:30-36 */
NZ_API int32_t nz_limit_width(int32_t width);
/* SeparableKernelFilter tables, Filter/Kernel/KernelJob.cs:97-136,217-292: fills kx,kz (capacity
 * >= 9), *ksize, *factor for every filter except SOBEL3_2D (returns NZ_E_UNSUPPORTED there:
 * that one is a two-branch reduce, KernelJob.cs:187-215). */
NZ_API int32_t nz_kernel_filter_table(int32_t filter_type, float* kx, float* kz, int32_t* ksize, float* factor);
/* MeshTileGenerator tile maths, Scripts/MeshTileGenerator.cs:166-177,197-206 */
NZ_API int32_t nz_tile_geometry(int32_t tile_resolution, int32_t tile_size, int32_t margin,
                                int32_t* mesh_resolution, int32_t* margin_pix, float* mesh_tile_size);

/* ---- host layer: one call per reference delegate ------------------------------------- */

/* FractalJobDelegate, Noise/Fractal/Fractal.cs:76-88, table Noise/NoiseStage.cs:26-35.
 * dst.length must be resolution*resolution. */
NZ_API int32_t nz_fractal(nz_slice_f32 dst, int32_t resolution, int32_t noise_type, float hurst,
                          float starting_amplitude, float stepdown, float detune_rate, int32_t octaves,
                          int32_t xpos, int32_t zpos, int32_t noise_size);

/* SeperableKernelFilterDelegate, Filter/Kernel/KernelJob.cs:308-314, applied `iterations` times as
 * KernelFilterStage.Schedule does (Filter/KernelFilterStage.cs:31-43).  `tmp` is accepted for
 * signature parity and never touched (the GPU path ping-pongs in HBM); it may be {NULL,0,0}. */
NZ_API int32_t nz_kernel_filter(nz_slice_f32 src, nz_slice_f32 tmp, int32_t filter_type,
                                int32_t resolution, int32_t iterations);

/* SeparableKernelFilter.ScheduleSeries, KernelJob.cs:165-185: arbitrary odd ksize <= 25. */
NZ_API int32_t nz_separable(nz_slice_f32 src, nz_slice_f32 tmp, int32_t ksize, const float* kx,
                            const float* kz, float factor, int32_t resolution, int32_t iterations);

/* GaussFilter.GaussFilterDelegate, Filter/Kernel/Blur/BlurJob.cs:23-30 (+StageGaussianBlur.cs:33-46) */
NZ_API int32_t nz_gauss_filter(nz_slice_f32 src, nz_slice_f32 tmp, int32_t width, int32_t sigma,
                               int32_t resolution, int32_t iterations);
/* SmoothFilter.SmoothFilterDelegate, BlurJob.cs:46-52 (+StageSmoothBlur.cs:33-46) */
NZ_API int32_t nz_smooth_filter(nz_slice_f32 src, nz_slice_f32 tmp, int32_t width,
                                int32_t resolution, int32_t iterations);

/* ErosionKernelJobDelegate, KernelJob.cs:350 (Schedule :333-347), applied `iterations` times
 * ("Value Erosion"): per call X pass min(src(x-1,z),src(x,z)) then Z pass likewise. */
NZ_API int32_t nz_min_erosion(nz_slice_f32 src, int32_t resolution, int32_t iterations);

/* FlowMapStage.ScheduleAll, Geologic/Stage/FlowMapStage.cs:124-195: fill water 1e-4, `iterations` x
 * (outflow step, water step), velocity magnitude written over `height`, then (v-normMin)/(normMax-normMin).
 * Flow fields start at zero (the reference leaves them uninitialised, FlowMapStage.cs:55-62). */
NZ_API int32_t nz_flowmap(nz_slice_f32 height, int32_t resolution, int32_t iterations,
                          float norm_min, float norm_max);

/* HeightMapMeshJobScheduleDelegate, Mesh/Job/HeightMapMeshJob.cs:55-65, table MeshTileStage.cs:30-33.
 * vertices: (resolution+1)^2 * 48 B (Mesh.MeshData.GetVertexData<Stream0>, PositionStream.cs:121);
 * indices : 6*resolution^2 uint32     (GetIndexData<uint>, :122).  Both HOST pointers. */
NZ_API int32_t nz_heightmap_mesh(int32_t mesh_type, void* vertices, uint32_t* indices,
                                 int32_t resolution, int32_t input_resolution, int32_t margin_pix,
                                 float tile_height, float tile_size, nz_slice_f32 heights);

/* ---- host layer: the "next" rows of SURVEY.md section 8f ------------------------------- */

/* ThermalErosionFilterDelegate, Filter/Kernel/Blur/ThermalErosionFilter.cs:138-146 (Schedule :111-133; stage
 * StageThermalErosion.cs:13-29): `iterations` x 4 phases of in-place 2x2 talus relaxation.
 * talus in degrees; maxDiff = tan(talus/90 * 3.14159/2) * mesh_height_width_ratio / resolution. */
NZ_API int32_t nz_thermal_erosion(nz_slice_f32 src, float talus, float increment_ratio,
                                  float mesh_height_width_ratio, int32_t iterations, int32_t resolution);

/* ErosionStageSubtractiveFlow.Schedule -> ScheduleAll -> ScheduleCycle, Geologic/Stage/ErosionStageSubtractiveFlow.cs:138-245
 * (the whole stage is commented-out code upstream; this entry follows its text).  Cycle n = 0..erosive_iterations-1:
 * water := 1e-4, n + 1 x (outflow step, water step) on flow fields that persist across the cycles (zero before the
 * first), then height -= erosive_factor * (|velocity| - norm_min) / (norm_max - norm_min).  The stage's
 * `flowIterations` field is never read upstream (:19-20) and has no argument here. */
NZ_API int32_t nz_subtractive_flow_erosion(nz_slice_f32 height, int32_t resolution, int32_t erosive_iterations,
                                           float erosive_factor, float norm_min, float norm_max);

/* ConstantJobScheduleDelegate, Filter/ConstantJob.cs:49-55 (operators Operators/SimpleMutation.cs:16-54):
 * MULTIPLY: v * value;  BINARIZE: v >= value ? 1 : 0.  `tmp` is accepted for signature parity, never touched. */
NZ_API int32_t nz_constant(nz_slice_f32 src, nz_slice_f32 tmp, int32_t operation, float constant_value,
                           int32_t resolution);

/* ReductionJobScheduleDelegate, Filter/ReductionJob.cs:55-61 (operators SimpleMutation.cs:56-171):
 * left = op(left, right); ROOTSUMSQUARES = sqrt(a*a + b*b). */
NZ_API int32_t nz_reduce(nz_slice_f32 left, nz_slice_f32 right, nz_slice_f32 tmp, int32_t operation,
                         int32_t resolution);

/* CurveJobScheduleDelegate, Filter/Curve/CurveJob.cs:91-97 (CurveOperator.Apply :69-80): piecewise-linear
 * LUT over [0,1]; `curve` holds the samples CurveStage.ExtractCurve takes at i/samples (CurveStage.cs:27-35). */
NZ_API int32_t nz_curve(nz_slice_f32 src, nz_slice_f32 tmp, nz_slice_f32 curve, int32_t resolution);

/* CropJobDelegate, Filter/Sample/CropJob.cs:63-69: output(x,z) = input(x+offset, z+offset), clamped reads.
 * The reference never assigns CropJob.Offset (:25,43-59), i.e. offset == 0; pass (input_resolution -
 * output_resolution)/2 for the centre crop the stage's menu name promises. */
NZ_API int32_t nz_crop(nz_slice_f32 input, int32_t input_resolution, nz_slice_f32 output,
                       int32_t output_resolution, int32_t offset);

/* GetMapRangeJob, Filter/NormalizeJob.cs:18-53: res3 = {min, max, max - min} of the map, folded from
 * lim_min / lim_max (the reference's defaults are +inf / -inf). */
NZ_API int32_t nz_map_range(nz_slice_f32 map, float* res3, float lim_min, float lim_max);

/* MapNormalizeValuesDelegate, Filter/NormalizeJob.cs:94-100 with NormalizeMap (FlowMapComponents.cs:150-166):
 * args3 = {min, max, range}; v = range < 1e-12 ? 0 : v;  out = (v - min) / range. */
NZ_API int32_t nz_normalize(nz_slice_f32 src, nz_slice_f32 tmp, const float* args3, int32_t resolution);

/* ---- host layer: residency ---------------------------------------------------------- */

/* Between begin/end (per calling thread) the device mirror of every host slice a stage touches
 * stays in HBM and D2H is deferred: chained stages pay one H2D (none if the first stage is a
 * generator) and one D2H.  nz_pipeline_end flushes every dirty mirror to its host slice. */
NZ_API int32_t nz_pipeline_begin(void);
NZ_API int32_t nz_pipeline_end(void);
/* The same residency as an explicit, process-wide scope, for hosts that run the stages of one chain on
 * DIFFERENT threads (Unity executes every IJob on an arbitrary worker): create the scope when the chain is
 * scheduled, bracket each stage's native call with enter/leave on whatever thread runs it, close it (flushes every
 * dirty mirror to host memory, then frees) when the chain completes.  Streams of different threads are ordered
 * through per-mirror events.  nz_pipeline_begin/end == create+enter / close on one thread. */
NZ_API int64_t nz_scope_create(void);            /* > 0: scope id; < 0: NZ_E* */
NZ_API int32_t nz_scope_enter(int64_t scope);
NZ_API int32_t nz_scope_leave(void);
NZ_API int32_t nz_scope_close(int64_t scope);
/* Flush one mirror early (e.g. before a Burst stage reads the slice). */
NZ_API int32_t nz_flush_to_host(const float* host_ptr);
/* cudaHostRegister / Unregister a long-lived host allocation (PipelineStateManager buffers). */
NZ_API int32_t nz_pin(void* host_ptr, size_t bytes);
NZ_API int32_t nz_unpin(void* host_ptr);

/* ---- named device-resident buffers (the GPU side of PipelineStateManager) ------------------------------------------ */
/* The reference parks tiles between pipelines in named host buffers of its PipelineStateManager
 * (Pipeline/PipelineState/PipelineStateManager.cs:39-127; name = "{xpos}_{zpos}__{resolution}__{contextAlias}",
 * PipelineState/Stage/WriteGeneratorContextStage.cs:21-23) and copies them with FlushWriteSlice jobs
 * (WriteGeneratorContextStage.cs:36-52, ReadGeneratorContextStage.cs:39-51).  Here the named buffer lives in HBM and
 * outlives residency scopes, so a tile written by one pipeline is read by the next without touching the host:
 *   nz_context_write: named := src   (device-to-device when src is resident in the caller's scope, one H2D otherwise)
 *   nz_context_read : dst := named   (device-to-device into dst's mirror; the host slice follows at the scope's close)
 * Names are process-wide.  Buffers hold `length` floats on the host layer's device. */
NZ_API int32_t nz_context_write(const char* name, nz_slice_f32 src);
NZ_API int32_t nz_context_read(const char* name, nz_slice_f32 dst);
/* *length = element count, or -1 when no buffer of that name exists (PipelineStateManager.BufferExists) */
NZ_API int32_t nz_context_exists(const char* name, int32_t* length);
/* host copies for PipelineStateManager.SaveBufferToDisk / the saved-state load in GetBuffer (:64-72, :104-120) */
NZ_API int32_t nz_context_download(const char* name, float* h_dst, int32_t length);
NZ_API int32_t nz_context_upload(const char* name, const float* h_src, int32_t length);
/* PipelineStateManager.ReleaseBuffer (:129-134); returns NZ_OK also when the name is unknown */
NZ_API int32_t nz_context_release(const char* name);

/* ---- multi-GPU: row bands of one large heightmap -------------------------------------- */
/* The reference generates one tile at a time on the CPU (Scripts/MeshTileGenerator.cs:125-138,184-192); a heightmap too
 * large or too slow for one GPU is split into ROW BANDS: band b of g owns rows [b*n/g, (b+1)*n/g).  Noise needs no
 * communication; the iterated filter, the flow map, the value erosion and the mesh consume ghost rows of their
 * neighbours (r*iterations, 2*iterations+1, iterations above only, 1), exchanged ONCE per stage.  Results are bit-identical
 * to the single-GPU chain.  Two ways to reach it, both behind this C ABI:
 *   (1) nz_init(devices, n) + nz_set_bands(n): every host-layer stage call (nz_fractal, nz_kernel_filter, ...) on a grid of
 *       at least NZ_BANDS_MIN_RESOLUTION rows transparently runs on n bands, one per device, in ONE process (what a Unity
 *       host is); ghost rows move by peer-to-peer copies over NVLink, each band uploads / downloads its own rows.
 *   (2) nz_band_chain_*: the whole BASELINE C5 chain on device-resident bands; one band per PROCESS with the ghost rows
 *       exchanged by ncclSend/ncclRecv (nz_comm_*; libnccl.so.2 is bound at run time), or all bands in one process. */
#define NZ_BANDS_MIN_RESOLUTION 4096
#define NZ_COMM_ID_BYTES 128            /* sizeof(ncclUniqueId) */
#define NZ_BANDS_EXCHANGE 0             /* the ghost rows of the noise stage come from the neighbours, once per pass */
#define NZ_BANDS_RECOMPUTE 1            /* the noise stage evaluates every ghost row the chain will consume: no communication */

/* Host layer: split large grids over the first n_bands devices given to nz_init (0 or 1: off).  Devices may repeat
 * (e.g. {0,0}: two bands on one GPU, which is how the banded path is tested on a one-GPU box). */
NZ_API int32_t nz_set_bands(int32_t n_bands);

/* The chain of BASELINE.json configs[4] (one resolution^2 heightmap): fBm noise -> separable filter -> flow map -> value
 * erosion -> mesh, with the stage parameters of the reference stages (NoiseStage.cs:37-54, KernelFilterStage.cs:17-19,
 * FlowMapStage.cs:18-23, ErosionKernelJob, MeshStageData.cs:9-20).  *_iterations == 0 skips that stage; mesh_resolution == 0
 * skips the mesh. */
typedef struct {
    int32_t resolution;
    int32_t noise_type; float hurst, starting_amplitude, stepdown, detune_rate; int32_t octaves, xpos, zpos, noise_size;
    int32_t filter_type, filter_iterations;
    int32_t flow_iterations; float norm_min, norm_max;
    int32_t erosion_iterations;
    int32_t mesh_type, mesh_resolution, mesh_margin_pix; float tile_height, tile_size;
} nz_chain_config;

typedef struct {
    int32_t rank, world, device;
    int32_t z0, z1;                 /* owned heightmap rows [z0, z1) */
    int32_t vz0, vz1;               /* owned vertex rows [vz0, vz1); owned triangle rows [max(vz0,1), vz1) */
    float* d_rows;                  /* DEVICE pointer of owned row z0 of the last run's result (contiguous rows) */
    void* d_vertices;               /* DEVICE pointer of vertex row vz0 (48-byte vertices) */
    uint32_t* d_indices;            /* DEVICE pointer of triangle row max(vz0,1) */
    int64_t halo_bytes_per_run;     /* ghost-row bytes this band received per run */
} nz_band_info;

/* Host logic of the partition, no GPU needed: rows [z0, z1) and vertex rows [vz0, vz1) band `rank` of `world` owns of a
 * resolution^2 grid meshed at mesh_resolution (0: no mesh).  Fills rank, world, z0, z1, vz0, vz1 of *out; the rest is zero. */
NZ_API int32_t nz_band_geometry(int32_t resolution, int32_t world, int32_t rank, int32_t mesh_resolution, nz_band_info* out);

/* NCCL communicator over the band processes (one rank per GPU).  Rank 0 calls nz_comm_unique_id and hands the 128 bytes
 * to the others through whatever channel the host has; every rank then calls nz_comm_create.  > 0: handle; < 0: NZ_E*. */
NZ_API int32_t nz_comm_unique_id(void* id_bytes, int32_t capacity);
NZ_API int64_t nz_comm_create(const void* id_bytes, int32_t world, int32_t rank, int32_t device);
/* polls ncclCommGetAsyncError: NZ_OK, or NZ_E_CUDA with the NCCL error in nz_last_error() */
NZ_API int32_t nz_comm_async_error(int64_t comm);
NZ_API int32_t nz_comm_destroy(int64_t comm);

/* One band of the chain in this process (rank/world/device of `comm`; comm == 0: a single band on `device`).  `stream` is
 * the cudaStream_t the band's work is enqueued on (NULL = the legacy default stream, as in the device layer;
 * NZ_STREAM_OWN = a non-blocking stream the library creates). */
#define NZ_STREAM_OWN ((void*)(intptr_t)-1)
NZ_API int64_t nz_band_chain_create(const nz_chain_config* cfg, int64_t comm, int32_t device, int32_t mode, void* stream);
/* All n_bands bands in this process, band b on devices[b] (entries may repeat); ghost rows move by peer copies. */
NZ_API int64_t nz_band_chain_create_local(const nz_chain_config* cfg, const int32_t* devices, int32_t n_bands, int32_t mode);
/* Enqueue one pass of the chain (asynchronous).  timed != 0 also records a CUDA event before every stage. */
NZ_API int32_t nz_band_chain_run(int64_t chain, int32_t timed);
NZ_API int32_t nz_band_chain_sync(int64_t chain);
/* milliseconds of {noise, filter, flow, erosion, mesh} of the last timed run (max over this process's bands; synchronises) */
NZ_API int32_t nz_band_chain_stage_ms(int64_t chain, float* ms5);
NZ_API int32_t nz_band_chain_local_bands(int64_t chain);
NZ_API int32_t nz_band_chain_info(int64_t chain, int32_t local_band, nz_band_info* out);
/* Copy this process's bands into FULL-GRID host buffers (each band writes its own slice; any pointer may be NULL):
 * h_heights resolution^2 floats, h_vertices (R+1)^2 * 48 bytes, h_indices 6*R^2 uint32.  Synchronises. */
NZ_API int32_t nz_band_chain_download(int64_t chain, float* h_heights, void* h_vertices, uint32_t* h_indices);
NZ_API int32_t nz_band_chain_destroy(int64_t chain);

/* ---- tiled worlds: several tiles in flight on one GPU ------------------------------------------------------------------- */
/* The reference generates a world tile by tile, ONE tile in flight: MeshTileGenerator hands tile (tx, tz) to the generator
 * pipeline with xpos = tileResolution * tx, zpos = tileResolution * tz (Scripts/MeshTileGenerator.cs:125-138,184-192).  A
 * 1024^2 tile cannot fill 148 SMs, so a tile world keeps `slots` tiles in flight on separate streams and runs, per tile,
 * the chain of BASELINE.json configs[3]: fBm noise -> separable filter x iterations -> [edge filter (e.g. Sobel3_2D) on a
 * COPY] -> mesh.  Tiles are independent (noise is a pure function of position, every filter clamps at the tile's own border),
 * so sharding a world over GPUs needs no communication: each process runs its own tiles. */
typedef struct {
    int32_t resolution;             /* generator resolution of one tile (e.g. 1024) */
    int32_t tile_resolution;        /* world cells between tile origins (e.g. 1000: tiles overlap by resolution - tile_resolution) */
    int32_t noise_type; float hurst, starting_amplitude, stepdown, detune_rate; int32_t octaves, noise_size;
    int32_t filter_type, filter_iterations;
    int32_t edge_filter_type, edge_filter_iterations;   /* on a copy of the filtered tile; 0 iterations: none */
    int32_t mesh_type, mesh_resolution, mesh_margin_pix; float tile_height, tile_size;   /* mesh_resolution 0: no mesh */
} nz_tile_config;

NZ_API int64_t nz_tile_world_create(const nz_tile_config* cfg, int32_t device, int32_t slots);
/* Runs the chain for n tiles, tiles_xz = {tx0, tz0, tx1, tz1, ...}.  Host outputs are per-tile arrays laid out tile after
 * tile (resolution^2 floats of heights / edges, (R+1)^2 * 48 bytes of vertices, 6 R^2 indices); any of them may be NULL, in
 * which case that output stays in the slot's device buffers and is overwritten by the slot's next tile.  Synchronises. */
NZ_API int32_t nz_tile_world_run(int64_t world, const int32_t* tiles_xz, int32_t n, float* h_heights, float* h_edges,
                                 void* h_vertices, uint32_t* h_indices);
/* DEVICE pointers of the last tile a slot ran (heights, edges, vertices, indices), for tests and device-resident consumers */
NZ_API int32_t nz_tile_world_slot(int64_t world, int32_t slot, float** d_heights, float** d_edges, void** d_vertices,
                                  uint32_t** d_indices);
NZ_API int32_t nz_tile_world_destroy(int64_t world);

/* ---- device layer ------------------------------------------------------------------- */
/* All pointers are DEVICE pointers unless named h_*. `stream` is a cudaStream_t (NULL = legacy
 * default stream).  Grids are width x rows, contiguous.  Calls only enqueue work.           */

/* Noise rows [z_first, z_first+rows) of the tile whose origin is (xpos,zpos): cell (x,r) gets
 * NoiseValue(x, z_first+r) of Fractal.cs:114-131. */
NZ_API int32_t nz_dev_fractal(float* d_dst, int32_t width, int32_t rows, int32_t z_first,
                              int32_t noise_type, float hurst, float starting_amplitude, float stepdown,
                              float detune_rate, int32_t octaves, int32_t xpos, int32_t zpos,
                              int32_t noise_size, void* stream);

/* `iterations` x (X pass, Z pass).  Input in d_data; d_tmp is scratch of the same size.  The
 * result is left in whichever of the two the last pass wrote; *d_result receives that pointer.
 * Pass d_result==NULL to force the result into d_data (may cost one extra copy). */
NZ_API int32_t nz_dev_separable(float* d_data, float* d_tmp, int32_t width, int32_t rows,
                                int32_t ksize, const float* h_kx, const float* h_kz, float factor,
                                int32_t iterations, float** d_result, void* stream);
NZ_API int32_t nz_dev_kernel_filter(float* d_data, float* d_tmp, int32_t width, int32_t rows,
                                    int32_t filter_type, int32_t iterations, float** d_result, void* stream);
NZ_API int32_t nz_dev_min_erosion(float* d_data, float* d_tmp, int32_t width, int32_t rows,
                                  int32_t iterations, float** d_result, void* stream);

/* Flow map.  d_tmp is the ping-pong partner of d_height (same size).  For 1 <= iterations <= 5 on an
 * even width the whole stage is ONE fused launch that reads d_height and writes d_tmp (state lives in
 * shared memory) and needs no scratch; otherwise water + 4 flow fields are kept in d_scratch
 * (nz_dev_flowmap_scratch_bytes, 0 when the fused path applies) and the result lands in d_height. */
NZ_API size_t  nz_dev_flowmap_scratch_bytes(int32_t width, int32_t rows, int32_t iterations);
NZ_API int32_t nz_dev_flowmap(float* d_height, float* d_tmp, void* d_scratch, int32_t width, int32_t rows,
                              int32_t iterations, float norm_min, float norm_max,
                              float** d_result, void* stream);

/* Introspection: the register-resident flow kernel divides and takes square roots on guarded fast paths; a launch in
 * which any cell left the guard (heights beyond ~1e18, quotients in the denormals) is rerun on the wavefront kernel.
 * Returns how many launches on the current device were rerun so far (synchronises the device). */
NZ_API int32_t nz_dev_flow_walk_reruns(uint64_t* count);

/* Vertex rows [vz_begin, vz_end) (0 <= vz < resolution+1) and the triangle rows they close
 * (row z>0 writes the 2*resolution triangles between vertex rows z-1 and z).
 * d_heights points at input row `h_row_first` of the input_resolution^2 height grid and holds
 * `h_rows` rows; d_vertices points at vertex row vz_begin; d_indices at triangle row
 * max(vz_begin,1).  Index VALUES are global. */
NZ_API int32_t nz_dev_heightmap_mesh(int32_t mesh_type, void* d_vertices, uint32_t* d_indices,
                                     int32_t resolution, int32_t input_resolution, int32_t margin_pix,
                                     float tile_height, float tile_size, const float* d_heights,
                                     int32_t h_row_first, int32_t h_rows,
                                     int32_t vz_begin, int32_t vz_end, void* stream);

/* Device-layer forms of the section-8f rows (n = number of cells; all in place). */
/* d_scratch: nz_dev_subtractive_flow_scratch_bytes() bytes (water + 4 flow fields, live for the whole call) */
NZ_API size_t  nz_dev_subtractive_flow_scratch_bytes(int32_t width, int32_t rows);
NZ_API int32_t nz_dev_subtractive_flow_erosion(float* d_height, void* d_scratch, int32_t width, int32_t rows,
                                               int32_t erosive_iterations, float erosive_factor,
                                               float norm_min, float norm_max, void* stream);
/* d_tmp (may be NULL): ping-pong partner of d_data; with it each iteration is one fused launch (the four phases on a
 * shared-memory tile) and the result lands in *d_result (d_data or d_tmp; copied back into d_data when d_result is
 * NULL).  Without it the phases are four in-place launches per iteration. */
NZ_API int32_t nz_dev_thermal_erosion(float* d_data, float* d_tmp, int32_t resolution, float talus, float increment_ratio,
                                      float mesh_height_width_ratio, int32_t iterations, float** d_result, void* stream);
NZ_API int32_t nz_dev_constant(float* d_data, size_t n, int32_t operation, float constant_value, void* stream);
NZ_API int32_t nz_dev_reduce(float* d_left, const float* d_right, size_t n, int32_t operation, void* stream);
NZ_API int32_t nz_dev_curve(float* d_data, size_t n, const float* d_curve, int32_t curve_size, void* stream);
NZ_API int32_t nz_dev_crop(const float* d_input, int32_t input_resolution, float* d_output,
                           int32_t output_resolution, int32_t offset, void* stream);
/* d_res3: 3 floats {min, max, range}; d_scratch: nz_dev_map_range_scratch_bytes() bytes */
NZ_API size_t  nz_dev_map_range_scratch_bytes(void);
NZ_API int32_t nz_dev_map_range(const float* d_map, size_t n, float lim_min, float lim_max, float* d_res3,
                                void* d_scratch, void* stream);
NZ_API int32_t nz_dev_normalize(float* d_data, size_t n, float vmin, float range, void* stream);

/* FP32 FMA-pipe peak micro-benchmark: launches `grid` CTAs of 256 threads each running `iters`
 * x 16 independent dependent-chain FFMAs per thread; returns FLOP count in *flops.  Used by
 * bench.py to measure the FP32 roofline denominator on the box it runs on. */
NZ_API int32_t nz_dev_fma_peak(float* d_sink, int32_t grid, int32_t iters, double* flops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NOIZE_B200_H */
