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

}  // namespace

// ThermalErosionFilter.Schedule, :111-133 (host float arithmetic in the reference's order)
float thermal_max_diff(float talus_deg, float height_ratio, int resolution) {
    const float talus = (talus_deg / 90.0f) * 3.14159f / 2.0f;
    return (tanf(talus) * height_ratio) / (float)resolution;
}

int32_t launch_thermal_erosion(float* d_data, int res, float talus_deg, float increment, float height_ratio, int iterations,
                               cudaStream_t s) {
    const int jobs = res / 2 - 1;           // ScheduleParallel(((int) resolution / 2) - 1, ...)
    if (jobs <= 0 || iterations <= 0) return NZ_OK;
    const float max_diff = thermal_max_diff(talus_deg, height_ratio, res);
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
