// flow_kernels.cu — pipe-model flow map (hot loop 3).
//
// Replaces FlowMapStage.ScheduleAll (Geologic/Stage/FlowMapStage.cs:124-195) and its jobs:
//   FillArrayJob                      Geologic/FlowMap/FlowMapComponents.cs:176-202   water := 1e-4
//   ComputeFlowStep.CalculateCell     FlowMapComponents.cs:20-65   (job FlowMapJob.cs:17-80)
//   UpdateWaterStep.CalculateCell     FlowMapComponents.cs:81-104  (job FlowMapJob.cs:101-153)
//   CreateVelocityField.CalculateCell FlowMapComponents.cs:120-139 (job FlowMapJob.cs:168-218)
//   NormalizeMap.CalculateCell        FlowMapComponents.cs:157-165 (job Filter/NormalizeJob.cs:58-92)
// Flow fields start at zero (the reference leaves them uninitialised, FlowMapStage.cs:55-62).
//
// State per cell: water + 4 outflows (W,E,S,N).  The outflow step reads only the cell's own flows
// and the 5-point (height+water) stencil, the water step only the cell's own water and the 5-point
// flow stencil, so both update IN PLACE (the reference's READ/WRITE pairs + copy-backs vanish).
#include "nz_common.cuh"

namespace nz {
namespace {

constexpr int TX = 128;
constexpr float TIMESTEP = 0.2f;

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

struct FlowFields {
    float* water;
    float* fW;
    float* fE;
    float* fS;
    float* fN;
};

// FIRST: water is still the fill value everywhere (nothing to read); ZERO: the flow fields are still all zero.
// FlowMapStage has both in its first iteration; a later cycle of the subtractive-flow erosion refills the water
// but keeps its flows (FIRST && !ZERO).
template <bool FIRST, bool ZERO = FIRST>
__global__ void __launch_bounds__(TX) flow_step_kernel(const float* __restrict__ height, FlowFields f, int width, int rows) {
    const int z = blockIdx.y;
    const int x = blockIdx.x * TX + threadIdx.x;
    if (x >= width) return;
    const size_t i = (size_t)z * width + x;
    const int xW = max(x - 1, 0), xE = min(x + 1, width - 1);
    const size_t rS = (size_t)max(z - 1, 0) * width, rN = (size_t)min(z + 1, rows - 1) * width, r0 = (size_t)z * width;
    const float w0 = FIRST ? 0.0001f : f.water[i];
    const float h0 = __ldg(height + i);
    const float totalHt = w0 + h0;
    const float dW = totalHt - ((FIRST ? 0.0001f : f.water[r0 + xW]) + __ldg(height + r0 + xW));
    const float dE = totalHt - ((FIRST ? 0.0001f : f.water[r0 + xE]) + __ldg(height + r0 + xE));
    const float dS = totalHt - ((FIRST ? 0.0001f : f.water[rS + x]) + __ldg(height + rS + x));
    const float dN = totalHt - ((FIRST ? 0.0001f : f.water[rN + x]) + __ldg(height + rN + x));
    const float flW = fmaxf(0.0f, (ZERO ? 0.0f : f.fW[i]) + dW);
    const float flE = fmaxf(0.0f, (ZERO ? 0.0f : f.fE[i]) + dE);
    const float flS = fmaxf(0.0f, (ZERO ? 0.0f : f.fS[i]) + dS);
    const float flN = fmaxf(0.0f, (ZERO ? 0.0f : f.fN[i]) + dN);
    const float sum_ = (flW + flE) + (flS + flN);  // math.csum(float4)
    // clamp(w0 / (sum*dt), 0, 1) without dividing when the clamp decides (see flowwave_kernels.cu flow_cell)
    const float d = sum_ * TIMESTEP;
    float K = 0.0f;
    if (sum_ > 0.0f) {
        if (w0 >= d) K = 1.0f;
        else if (w0 > 0.0f) K = fminf(w0 / d, 1.0f);
    }
    // sum_ <= 0 means every flow is 0 already (they are clamped at 0), so flow*0 == the reference's explicit 0
    f.fW[i] = sum_ > 0.0f ? flW * K : 0.0f;
    f.fE[i] = sum_ > 0.0f ? flE * K : 0.0f;
    f.fS[i] = sum_ > 0.0f ? flS * K : 0.0f;
    f.fN[i] = sum_ > 0.0f ? flN * K : 0.0f;
}

template <bool FIRST>
__global__ void __launch_bounds__(TX) water_step_kernel(FlowFields f, int width, int rows) {
    const int z = blockIdx.y;
    const int x = blockIdx.x * TX + threadIdx.x;
    if (x >= width) return;
    const size_t i = (size_t)z * width + x;
    const int xW = max(x - 1, 0), xE = min(x + 1, width - 1);
    const size_t rS = (size_t)max(z - 1, 0) * width, rN = (size_t)min(z + 1, rows - 1) * width, r0 = (size_t)z * width;
    const float flowOUT = ((f.fW[i] + f.fE[i]) + f.fS[i]) + f.fN[i];
    float flowIN = 0.0f;
    flowIN += f.fE[r0 + xW];
    flowIN += f.fW[r0 + xE];
    flowIN += f.fN[rS + x];
    flowIN += f.fS[rN + x];
    const float ht = fmaf(flowIN - flowOUT, TIMESTEP, FIRST ? 0.0001f : f.water[i]);
    f.water[i] = fmaxf(0.0f, ht);
}

__global__ void __launch_bounds__(TX) velocity_norm_kernel(float* __restrict__ out, FlowFields f, int width, int rows,
                                                           float norm_min, float norm_range) {
    const int z = blockIdx.y;
    const int x = blockIdx.x * TX + threadIdx.x;
    if (x >= width) return;
    const size_t i = (size_t)z * width + x;
    const int xW = max(x - 1, 0), xE = min(x + 1, width - 1);
    const size_t rS = (size_t)max(z - 1, 0) * width, rN = (size_t)min(z + 1, rows - 1) * width, r0 = (size_t)z * width;
    const float dl = f.fE[r0 + xW] - f.fW[i];
    const float dr = f.fE[i] - f.fW[r0 + xE];
    const float dt = f.fS[rN + x] - f.fN[i];
    const float db = f.fS[i] - f.fN[rS + x];
    const float vx = (dl + dr) * 0.5f, vy = (dt + db) * 0.5f;
    float v = sqrtf(fmaf(vy, vy, vx * vx));
    if (norm_range < 1e-12f) v = 0.0f;
    out[i] = (v - norm_min) / norm_range;
}

// One cycle's epilogue of the subtractive-flow erosion, fused: velocity magnitude -> NormalizeMap -> ConstantMultiply ->
// SubtractTiles (ErosionStageSubtractiveFlow.cs:196-222).  Each of the four is one rounding, so the fused form has the
// bits of the four separate jobs.
__global__ void __launch_bounds__(TX) velocity_erode_kernel(float* __restrict__ height, FlowFields f, int width, int rows,
                                                            float norm_min, float norm_range, float erosive_factor) {
    const int z = blockIdx.y;
    const int x = blockIdx.x * TX + threadIdx.x;
    if (x >= width) return;
    const size_t i = (size_t)z * width + x;
    const int xW = max(x - 1, 0), xE = min(x + 1, width - 1);
    const size_t rS = (size_t)max(z - 1, 0) * width, rN = (size_t)min(z + 1, rows - 1) * width, r0 = (size_t)z * width;
    const float dl = f.fE[r0 + xW] - f.fW[i];
    const float dr = f.fE[i] - f.fW[r0 + xE];
    const float dt = f.fS[rN + x] - f.fN[i];
    const float db = f.fS[i] - f.fN[rS + x];
    const float vx = (dl + dr) * 0.5f, vy = (dt + db) * 0.5f;
    float v = sqrtf(fmaf(vy, vy, vx * vx));
    if (norm_range < 1e-12f) v = 0.0f;
    v = (v - norm_min) / norm_range;
    height[i] = height[i] - v * erosive_factor;
}

// iterations == 0: velocity of an all-zero flow field is 0
__global__ void __launch_bounds__(TX) fill_norm_zero_kernel(float* __restrict__ out, size_t n, float norm_min,
                                                            float norm_range) {
    size_t i = (size_t)blockIdx.x * TX + threadIdx.x;
    if (i < n) out[i] = (0.0f - norm_min) / norm_range;
}

}  // namespace

// scratch is only needed by the per-iteration path (water + 4 flow fields in HBM); the fused wavefront
// kernel (flowwave_kernels.cu) keeps all state in shared memory and needs just an output buffer
size_t flowmap_scratch_bytes(int width, int rows, int iterations) {
    // (the same predicate launch_flowmap uses, minus the pointers: a caller with a null or misaligned tmp buffer is told so there)
    if (!getenv("NZ_FLOW_UNFUSED") && flow_wave_supported(width, rows, iterations, nullptr, nullptr)) return 0;
    return (size_t)width * rows * sizeof(float) * 5;
}

int32_t launch_flowmap(float* d_height, float* d_tmp, void* d_scratch, int width, int rows, int iterations,
                       float norm_min, float norm_max, float** d_result, cudaStream_t s) {
    NZ_REQUIRE(d_height, "flowmap: null height buffer");
    NZ_REQUIRE(width > 0 && rows > 0 && rows <= 65535 && iterations >= 0, "flowmap: bad arguments");
    // NZ_FLOW_UNFUSED=1 forces the per-iteration kernels (used by the tests to cross-check the two paths bit for bit)
    // NZ_FLOW_PATH=tile|wave|reg picks one of the fused formulations (flowtile_ / flowwave_ / flowwalk_kernels.cu)
    if (d_tmp && !getenv("NZ_FLOW_UNFUSED") && flow_wave_supported(width, rows, iterations, d_height, d_tmp)) {
        const char* fp = getenv("NZ_FLOW_PATH");
        // measured at 16384^2 (register walk / tile / wave, ms): I=1 0.78 / 1.70 / 3.89, I=2 1.42 / 2.82 / 4.15,
        // I=3 2.32 / 4.25 / 4.71, I=4 3.17 / 5.62 / 5.91, I=5 4.73 / 7.63 / 6.89
        const bool walk = fp ? fp[0] == 'r' : true;
        const bool tile = fp ? fp[0] == 't' : iterations <= 4;
        int32_t rc = walk && flow_walk_supported(width, rows, iterations, d_height, d_tmp) && flow_walk_range_ok(norm_min, norm_max)
                         ? launch_flow_walk(d_height, d_tmp, width, rows, iterations, norm_min, norm_max, s)
                     : tile && flow_tile_supported(width, rows, iterations, d_height, d_tmp)
                         ? launch_flow_tile(d_height, d_tmp, width, rows, iterations, norm_min, norm_max, s)
                         : launch_flow_wave(d_height, d_tmp, width, rows, iterations, norm_min, norm_max, s);
        if (rc != NZ_OK) return rc;
        if (d_result) {
            *d_result = d_tmp;
        } else {
            NZ_CUDA(cudaMemcpyAsync(d_height, d_tmp, (size_t)width * rows * sizeof(float), cudaMemcpyDeviceToDevice, s));
        }
        return NZ_OK;
    }
    const size_t n = (size_t)width * rows;
    const float norm_range = norm_max - norm_min;  // FlowMapStage.cs:48-51
    dim3 grid(cdiv(width, TX), rows);
    if (iterations == 0) {
        fill_norm_zero_kernel<<<cdiv((long long)n, TX), TX, 0, s>>>(d_height, n, norm_min, norm_range);
        NZ_LAUNCHED();
        if (d_result) *d_result = d_height;
        return NZ_OK;
    }
    NZ_REQUIRE(d_scratch, "flowmap: null scratch buffer (need nz_dev_flowmap_scratch_bytes)");
    float* sc = (float*)d_scratch;
    FlowFields f = {sc, sc + n, sc + 2 * n, sc + 3 * n, sc + 4 * n};
    for (int it = 0; it < iterations; it++) {
        if (it == 0) {
            flow_step_kernel<true><<<grid, TX, 0, s>>>(d_height, f, width, rows);
            NZ_LAUNCHED();
            water_step_kernel<true><<<grid, TX, 0, s>>>(f, width, rows);
            NZ_LAUNCHED();
        } else {
            flow_step_kernel<false><<<grid, TX, 0, s>>>(d_height, f, width, rows);
            NZ_LAUNCHED();
            water_step_kernel<false><<<grid, TX, 0, s>>>(f, width, rows);
            NZ_LAUNCHED();
        }
    }
    velocity_norm_kernel<<<grid, TX, 0, s>>>(d_height, f, width, rows, norm_min, norm_range);
    NZ_LAUNCHED();
    if (d_result) *d_result = d_height;
    return NZ_OK;
}

// ErosionStageSubtractiveFlow.ScheduleAll / ScheduleCycle (Geologic/Stage/ErosionStageSubtractiveFlow.cs:138-230; SURVEY 8f
// rank 3 — commented-out code upstream, built from its text): cycle n runs n + 1 flow iterations on refilled water and
// PERSISTING flow fields, then erodes the heights by erosive_factor x the normalised velocity magnitude.  The heights
// change between cycles and the flows carry over, so the cycles run on the per-iteration kernels with water + 4 flows in
// d_scratch (5 fields).  The erosion of cycle n reads only the flows, so it may update the heights in place.
// Fused form (NZ_SUBFLOW_PATH=unfused turns it off): up to 5 cycles on widths that are a multiple of 4 run ONE launch per
// cycle (launch_flow_tile_cycle, flowtile_kernels.cu): the cycle's n + 1 iterations stay in an SM-resident tile, the flow
// fields are read once and written once per cycle (two sets of 4 planes, ping-pong) and the erosion is the tile's epilogue
// into a second height buffer: 9 fields of scratch, 40 B/cell/CYCLE of HBM traffic instead of ~44 B/cell/ITERATION.
size_t subtractive_flow_scratch_bytes(int width, int rows) { return (size_t)width * rows * sizeof(float) * ((width & 3) == 0 ? 9 : 5); }

int32_t launch_subtractive_flow_erosion(float* d_height, void* d_scratch, int width, int rows, int erosive_iterations,
                                        float erosive_factor, float norm_min, float norm_max, cudaStream_t s) {
    NZ_REQUIRE(d_height, "subtractive flow erosion: null height buffer");
    NZ_REQUIRE(width > 0 && rows > 0 && rows <= 65535 && erosive_iterations >= 0, "subtractive flow erosion: bad arguments");
    if (erosive_iterations == 0) return NZ_OK;
    NZ_REQUIRE(d_scratch, "subtractive flow erosion: null scratch buffer (need nz_dev_subtractive_flow_scratch_bytes)");
    const size_t n = (size_t)width * rows;
    const float norm_range = norm_max - norm_min;
    float* sc = (float*)d_scratch;
    {
        const char* e = getenv("NZ_SUBFLOW_PATH");
        const bool fused = !(e && e[0] == 'u') && erosive_iterations <= 5 && (width & 3) == 0 &&
                           (((uintptr_t)d_height | (uintptr_t)d_scratch) & 15) == 0;
        if (fused) {
            float* fA = sc;              // 4 planes
            float* fB = sc + 4 * n;      // 4 planes
            float* hcur = d_height;
            float* hnext = sc + 8 * n;
            const float* fin = nullptr;  // zero before the first cycle
            float* fout = fA;
            for (int c = 0; c < erosive_iterations; c++) {
                int32_t rc = launch_flow_tile_cycle(hcur, hnext, fin, fout, n, width, rows, c + 1, norm_min, norm_max, erosive_factor, s);
                if (rc != NZ_OK) return rc;
                float* t = hcur; hcur = hnext; hnext = t;
                fin = fout;
                fout = fout == fA ? fB : fA;
            }
            if (hcur != d_height) NZ_CUDA(cudaMemcpyAsync(d_height, hcur, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
            return NZ_OK;
        }
    }
    FlowFields f = {sc, sc + n, sc + 2 * n, sc + 3 * n, sc + 4 * n};
    dim3 grid(cdiv(width, TX), rows);
    for (int c = 0; c < erosive_iterations; c++) {
        for (int it = 0; it <= c; it++) {
            if (it == 0 && c == 0) flow_step_kernel<true, true><<<grid, TX, 0, s>>>(d_height, f, width, rows);
            else if (it == 0) flow_step_kernel<true, false><<<grid, TX, 0, s>>>(d_height, f, width, rows);
            else flow_step_kernel<false, false><<<grid, TX, 0, s>>>(d_height, f, width, rows);
            NZ_LAUNCHED();
            if (it == 0) water_step_kernel<true><<<grid, TX, 0, s>>>(f, width, rows);
            else water_step_kernel<false><<<grid, TX, 0, s>>>(f, width, rows);
            NZ_LAUNCHED();
        }
        velocity_erode_kernel<<<grid, TX, 0, s>>>(d_height, f, width, rows, norm_min, norm_range, erosive_factor);
        NZ_LAUNCHED();
    }
    return NZ_OK;
}

}  // namespace nz
