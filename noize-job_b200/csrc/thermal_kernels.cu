// thermal_kernels.cu — thermal (talus) erosion, SURVEY.md section 8f rank 1.
//
// ThermalErosionFilter (Filter/Kernel/Blur/ThermalErosionFilter.cs:21-147; stage StageThermalErosion.cs:13-29;
// called per cycle by LiveErosion.cs:385): `iterations` x 4 phases ("flip" 0..3).  A phase relaxes disjoint 2x2
// blocks in place: block origin x = 1 + (flip & 1) + 2i  (x < res-1),  z = 2(j+1) - (flip > 1)  (j < res/2 - 1);
// inside a block the six cell pairs (xy, xz, xw, yz, yw, zw) are rectified IN THAT ORDER: if the two heights differ
// by more than maxDiff = tan(talus) * heightRatio / res, each moves increment * excess towards the other.
//
// The four phases tile the grid with 2x2 blocks at the four (odd/even x, odd/even z) alignments, so blocks of one
// phase never overlap (it is the reference's own 4-colouring, :137-144) and every block is one thread here.  A
// phase reads and writes each touched cell once: 8 B/cell per phase, HBM-bound; phases are separate launches because
// each depends on the previous one everywhere (blocks overlap across phases).  Row 0 and column 0 are never
// touched, as in the reference.
#include <math.h>
#include <stdlib.h>
#include "nz_common.cuh"

namespace nz {
namespace {

// rectify(float2), ThermalErosionFilter.cs:84-98.  a*b+c is written as fmaf (the oracle's canonical form).
__device__ __forceinline__ void rectify(float& a, float& b, float max_diff, float inc) {
    const float diff = fabsf(a - b);
    if (diff > max_diff) {
        const float excess = diff - max_diff;
        if (a > b) {
            b = fmaf(inc, excess, b);
            a = fmaf(-inc, excess, a);
        } else {
            a = fmaf(inc, excess, a);
            b = fmaf(-inc, excess, b);
        }
    }
}

// the same update without divergence: both candidates are formed and selected (a lane that does not move keeps its bits)
__device__ __forceinline__ void rectify_sel(float& a, float& b, float max_diff, float inc) {
    const float diff = fabsf(a - b);
    const float excess = diff - max_diff;
    const float sa = a > b ? -inc : inc;             // a moves down when it is the higher one
    const float na = fmaf(sa, excess, a), nb = fmaf(-sa, excess, b);
    const bool mv = diff > max_diff;
    a = mv ? na : a;
    b = mv ? nb : b;
}

constexpr int TH_THREADS = 128;

__global__ void __launch_bounds__(TH_THREADS) thermal_phase_kernel(float* __restrict__ data, int res, int x0, int z0, int jobs,
                                                                   float max_diff, float inc) {
    const int x = x0 + 2 * (blockIdx.x * TH_THREADS + threadIdx.x);
    const int j = blockIdx.y;
    if (x >= res - 1 || j >= jobs) return;
    const int z = z0 + 2 * j;
    float* r0 = data + (size_t)z * res + x;
    float* r1 = r0 + res;                 // z + 1 <= res - 1 for every scheduled job (getIdx's clamp never engages)
    float vx, vy, vz, vw;
    const bool aligned = ((((uintptr_t)r0) | ((uintptr_t)r1)) & 7) == 0;
    if (aligned) {
        const float2 a = *reinterpret_cast<const float2*>(r0), b = *reinterpret_cast<const float2*>(r1);
        vx = a.x; vy = a.y; vz = b.x; vw = b.y;
    } else {
        vx = r0[0]; vy = r0[1]; vz = r1[0]; vw = r1[1];
    }
    // rectifyNeighborhood, :73-80: x=(x,z) y=(x+1,z) z=(x,z+1) w=(x+1,z+1)
    rectify(vx, vy, max_diff, inc);
    rectify(vx, vz, max_diff, inc);
    rectify(vx, vw, max_diff, inc);
    rectify(vy, vz, max_diff, inc);
    rectify(vy, vw, max_diff, inc);
    rectify(vz, vw, max_diff, inc);
    if (aligned) {
        *reinterpret_cast<float2*>(r0) = make_float2(vx, vy);
        *reinterpret_cast<float2*>(r1) = make_float2(vz, vw);
    } else {
        r0[0] = vx; r0[1] = vy; r1[0] = vz; r1[1] = vw;
    }
}

// One whole iteration (the 4 phases) on a shared-memory tile: 8 B/cell of HBM traffic per iteration instead of 32.
// A CTA stages a (TT_W + 8) x (TT_H + 8) region (4-cell halo, clipped to the grid), runs the four phases on it in place
// (a thread per 2x2 block, exactly the blocks the reference schedules, restricted to those lying wholly inside the
// region), and writes back the TT_W x TT_H interior.  A block that straddles the region's edge is skipped, which leaves a
// stale cell; staleness spreads by one cell per phase (through the 2x2 block that contains it), so after four phases
// it has reached at most 4 cells inwards: exactly the halo.  Reads `src`, writes `dst` (tiles read their neighbours'
// cells, so the update cannot be in place).
constexpr int TT_W = 120, TT_H = 64, TT_HALO = 4, TT_THREADS = 256;   // region width 128 = 64 blocks = two full warp passes
constexpr int TT_PITCH = TT_W + 2 * TT_HALO + 1;   // odd pitch: the two rows of a block fall into different banks

__global__ void __launch_bounds__(TT_THREADS) thermal_tile_kernel(const float* __restrict__ src, float* __restrict__ dst, int res,
                                                                  int jobs, float max_diff, float inc) {
    __shared__ float t[(TT_H + 2 * TT_HALO) * TT_PITCH];
    const int tx = blockIdx.x * TT_W, tz = blockIdx.y * TT_H;
    const int X0 = max(tx - TT_HALO, 0), X1 = min(tx + TT_W + TT_HALO, res);
    const int Z0 = max(tz - TT_HALO, 0), Z1 = min(tz + TT_H + TT_HALO, res);
    const int rw = X1 - X0, rh = Z1 - Z0;
    for (int i = threadIdx.x; i < rw * rh; i += TT_THREADS) {
        const int z = i / rw, x = i - z * rw;
        t[z * TT_PITCH + x] = __ldg(src + (size_t)(Z0 + z) * res + X0 + x);
    }
    __syncthreads();
#pragma unroll 1
    for (int flip = 0; flip < 4; flip++) {
        const int x0 = 1 + (flip & 1), z0 = flip > 1 ? 1 : 2;
        // first block origin >= the region origin with the phase's parity; blocks need x+1 < min(X1, res) and the
        // reference's bounds x < res-1 (the same thing at the grid edge), z = z0 + 2j with j < jobs
        int xs = max(X0, x0); xs += (xs - x0) & 1;
        int zs = max(Z0, z0); zs += (zs - z0) & 1;
        const int zlast = min(Z1 - 2, z0 + 2 * (jobs - 1));
        const int nx = xs + 1 < X1 ? (X1 - 2 - xs) / 2 + 1 : 0;
        const int nz = zs <= zlast ? (zlast - zs) / 2 + 1 : 0;
        for (int bz = threadIdx.x >> 5; bz < nz; bz += TT_THREADS / 32)
            for (int bx = threadIdx.x & 31; bx < nx; bx += 32) {
                float* r0 = t + (zs + 2 * bz - Z0) * TT_PITCH + (xs + 2 * bx - X0);
                float* r1 = r0 + TT_PITCH;
                float vx = r0[0], vy = r0[1], vz = r1[0], vw = r1[1];
                rectify_sel(vx, vy, max_diff, inc);
                rectify_sel(vx, vz, max_diff, inc);
                rectify_sel(vx, vw, max_diff, inc);
                rectify_sel(vy, vz, max_diff, inc);
                rectify_sel(vy, vw, max_diff, inc);
                rectify_sel(vz, vw, max_diff, inc);
                r0[0] = vx; r0[1] = vy; r1[0] = vz; r1[1] = vw;
            }
        __syncthreads();
    }
    const int ow = min(TT_W, res - tx), oh = min(TT_H, res - tz);
    for (int i = threadIdx.x; i < ow * oh; i += TT_THREADS) {
        const int z = i / ow, x = i - z * ow;
        dst[(size_t)(tz + z) * res + tx + x] = t[(tz + z - Z0) * TT_PITCH + (tx + x - X0)];
    }
}

}  // namespace

// ThermalErosionFilter.Schedule, :111-133 (host float arithmetic in the reference's order)
float thermal_max_diff(float talus_deg, float height_ratio, int resolution) {
    const float talus = (talus_deg / 90.0f) * 3.14159f / 2.0f;
    return (tanf(talus) * height_ratio) / (float)resolution;
}

// d_tmp (optional): ping-pong partner of d_data.  With it every iteration is ONE launch of the tile kernel and the result
// lands in *d_result (d_data or d_tmp; copied back to d_data when d_result is null); without it the four phases of an
// iteration are four in-place launches.  NZ_THERMAL_PATH=phase forces the latter (the tests compare the two bit for bit).
int32_t launch_thermal_erosion(float* d_data, float* d_tmp, int res, float talus_deg, float increment, float height_ratio,
                               int iterations, float** d_result, cudaStream_t s) {
    if (d_result) *d_result = d_data;
    const int jobs = res / 2 - 1;           // ScheduleParallel(((int) resolution / 2) - 1, ...)
    if (jobs <= 0 || iterations <= 0) return NZ_OK;
    const float max_diff = thermal_max_diff(talus_deg, height_ratio, res);
    const char* tp = getenv("NZ_THERMAL_PATH");
    if (d_tmp && !(tp && tp[0] == 'p')) {
        NZ_REQUIRE(cdiv(res, TT_H) <= 65535, "nz_thermal_erosion: resolution %d too large", res);
        dim3 grid(cdiv(res, TT_W), cdiv(res, TT_H));
        float *a = d_data, *b = d_tmp;
        for (int it = 0; it < iterations; it++) {
            thermal_tile_kernel<<<grid, TT_THREADS, 0, s>>>(a, b, res, jobs, max_diff, increment);
            NZ_LAUNCHED();
            float* t = a; a = b; b = t;
        }
        if (d_result) *d_result = a;
        else if (a != d_data) NZ_CUDA(cudaMemcpyAsync(d_data, a, (size_t)res * res * sizeof(float), cudaMemcpyDeviceToDevice, s));
        return NZ_OK;
    }
    for (int it = 0; it < iterations; it++)
        for (int flip = 0; flip < 4; flip++) {
            const int x0 = 1 + (flip & 1), z0 = flip > 1 ? 1 : 2;
            const int nx = (res - x0) / 2;                  // x = x0 + 2i < res - 1
            if (nx <= 0) continue;
            dim3 grid(cdiv(nx, TH_THREADS), jobs);
            if (jobs > 65535) {
                set_error("nz_thermal_erosion: resolution %d too large", res);
                return NZ_E_UNSUPPORTED;
            }
            thermal_phase_kernel<<<grid, TH_THREADS, 0, s>>>(d_data, res, x0, z0, jobs, max_diff, increment);
            NZ_LAUNCHED();
        }
    return NZ_OK;
}

}  // namespace nz
