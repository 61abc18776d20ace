// mesh_kernels.cu — heightmap -> interleaved vertex stream + uint32 triangle indices (hot loop 4).
//
// Replaces HeightMapMeshJob<{SquareGridHeightMap,OvershootSquareGridHeightMap}, PositionStream32>
// (Mesh/Job/HeightMapMeshJob.cs:9-52, Mesh/Generators/SquareGridHeightMap.cs:59-105,
// Mesh/Generators/OvershootSquareGridHeightMap.cs:54-103, Mesh/Streams/PositionStream.cs:75-134,
// Mesh/Streams/Triangle.cs:19-28).  Vertex = {float3 pos, float3 normal, float4 tangent, float2 uv}
// = 48 B; triangle = 3 x uint32.  NormalStrength is hard-wired to 8 (HeightMapMeshJob.cs:41).
//
// Write-bandwidth bound (72 B written per quad vs 4 B read).  One CTA produces a strip of MESH_TX
// vertices of one vertex row: the 12 floats per vertex are staged in shared memory and streamed
// out as consecutive float4 (a warp writes 512 contiguous bytes per store), likewise the indices.
#include "nz_common.cuh"

namespace nz {
namespace {

constexpr int MESH_TX = 128;

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return fmaf(az, bz, fmaf(ay, by, ax * bx));
}

struct MeshParams {
    int R, inRes, off;
    int h_row_first, h_rows;  // rows of the height grid resident in d_heights
    int vz_begin, vz_end;     // vertex rows this launch produces
    float Height, TileSize;
};

template <int MESH_TYPE>
__global__ void __launch_bounds__(MESH_TX) mesh_kernel(float4* __restrict__ vtx_out, uint2* __restrict__ idx_out,
                                                       const float* __restrict__ heights, MeshParams p) {
    __shared__ float4 sv[MESH_TX * 3];
    __shared__ uint2 si[MESH_TX * 3];
    const int z = p.vz_begin + blockIdx.y;
    const int xs = blockIdx.x * MESH_TX;  // first vertex column of this strip
    const int x = xs + threadIdx.x;
    const int R = p.R;
    const int nvert = min(MESH_TX, R + 1 - xs);
    if (x <= R) {
        const float Rf = (float)R;
        const int lo = MESH_TYPE == NZ_MESH_SQUARE_GRID ? 0 : -p.off;
        const int hi = MESH_TYPE == NZ_MESH_SQUARE_GRID ? R + 1 : R + p.off;
        auto H = [&](int xx, int zz) -> float {
            xx = clampi(xx, lo, hi);
            zz = clampi(zz, lo, hi);
            return __ldg(heights + (size_t)(zz + p.off - p.h_row_first) * p.inRes + xx + p.off);
        };
        const float t = H(x, z);
        float l, r, u, d, uvx, uvz;
        if (MESH_TYPE == NZ_MESH_SQUARE_GRID) {
            // InterpolateEdge(a,b) = a - (b - a), SquareGridHeightMap.cs:36-38,70-73
            if (x > 0) l = H(x - 1, z); else l = t - (H(x + 1, z) - t);
            if (x < R - 1) r = H(x + 1, z); else r = t - (H(x - 1, z) - t);
            if (z > 0) u = H(x, z - 1); else { const float a = H(x, z + 1); u = a - (t - a); }
            if (z < R - 1) d = H(x, z + 1); else { const float a = H(x, z - 1); d = a - (t - a); }
            uvx = (float)x / (Rf + 1.0f);
            uvz = (float)z / (Rf + 1.0f);
        } else {
            l = H(x - 1, z);
            r = H(x + 1, z);
            u = H(x, z - 1);
            d = H(x, z + 1);
            uvx = (float)x / (Rf - 0.5f);
            uvz = (float)z / (Rf - 0.5f);
        }
        const float px = x == 0 ? -(0.5f * p.TileSize / Rf) : (float)x * p.TileSize / Rf - 0.5f;
        const float pz = (float)z * p.TileSize / Rf - 0.5f;
        const float py = t * p.Height;
        const float t1y = (r - l) / 2.0f, t2y = (u - d) / 2.0f;
        // cross(t2,t1) with t1=(4,t1y,0), t2=(0,t2y,4): math.cross(a,b) = (a*b.yzx - a.yzx*b).yzx
        const float tx = t2y * 0.0f - 4.0f * t1y;
        const float ty = 4.0f * 4.0f - 0.0f * 0.0f;
        const float tz = 0.0f * t1y - t2y * 4.0f;
        const float nx = (l - r) / 2.0f * 8.0f, ny = 2.0f / p.Height, nz = (u - d) / 2.0f * 8.0f;
        const float inv = 1.0f / sqrtf(dot3(nx, ny, nz, nx, ny, nz));
        sv[threadIdx.x * 3 + 0] = make_float4(px, py, pz, inv * nx);
        sv[threadIdx.x * 3 + 1] = make_float4(inv * ny, inv * nz, tx, ty);
        sv[threadIdx.x * 3 + 2] = make_float4(tz, 0.0f, uvx, uvz);
        if (z > 0 && x >= 1) {
            const unsigned vi = (unsigned)((R + 1) * z + x);
            const int q = threadIdx.x - (xs == 0 ? 1 : 0);  // quad slot within this strip
            si[q * 3 + 0] = make_uint2(vi - R - 2, vi - 1);
            si[q * 3 + 1] = make_uint2(vi - R - 1, vi - R - 1);
            si[q * 3 + 2] = make_uint2(vi - 1, vi);
        }
    }
    __syncthreads();
    // vertices: row z starts at vertex (R+1)*(z - vz_begin) of this launch's output slice
    float4* vdst = vtx_out + ((size_t)(R + 1) * (z - p.vz_begin) + xs) * 3;
    for (int i = threadIdx.x; i < nvert * 3; i += MESH_TX) vdst[i] = sv[i];
    if (z > 0) {
        // quads of this strip: columns max(xs,1)..xs+nvert-1 ; triangle row z-1 holds quads x=1..R
        const int qfirst = max(xs, 1);
        const int nq = xs + nvert - qfirst;
        const int trow = z - max(p.vz_begin, 1);
        uint2* idst = idx_out + ((size_t)R * trow + (qfirst - 1)) * 3;
        for (int i = threadIdx.x; i < nq * 3; i += MESH_TX) idst[i] = si[i];
    }
}

}  // namespace

int32_t launch_mesh(int mesh_type, void* d_vtx, uint32_t* d_idx, int R, int inRes, float tile_height, float tile_size,
                    const float* d_heights, int h_row_first, int h_rows, int vz_begin, int vz_end, cudaStream_t s) {
    NZ_REQUIRE(d_vtx && d_idx && d_heights, "mesh: null device buffer");
    NZ_REQUIRE(mesh_type == NZ_MESH_SQUARE_GRID || mesh_type == NZ_MESH_OVERSHOOT_SQUARE_GRID, "mesh: bad mesh_type %d", mesh_type);
    NZ_REQUIRE(R > 0 && inRes > 0 && inRes >= R, "mesh: need 0 < resolution <= input_resolution (got %d, %d)", R, inRes);
    NZ_REQUIRE((long long)(R + 1) * (R + 1) < (1LL << 32), "mesh: vertex count overflows uint32 indices");
    const int off = (inRes - R) / 2;  // PixOffset, SquareGridHeightMap.cs:33
    // the reference would read outside the height grid for these shapes (its getIdx clamps to
    // [0,R+1] resp. [-off,R+off] in tile space, not to the grid)
    const int max_col = mesh_type == NZ_MESH_SQUARE_GRID ? R + off : (R + 1 < R + off ? R + 1 : R + off) + off;
    NZ_REQUIRE(max_col <= inRes - 1, "mesh: input_resolution %d too small for resolution %d (reads column %d)", inRes, R, max_col);
    NZ_REQUIRE(0 <= vz_begin && vz_begin < vz_end && vz_end <= R + 1, "mesh: bad vertex row range [%d,%d)", vz_begin, vz_end);
    // rows of the height grid the requested vertex rows touch
    {
        int zlo = vz_begin - 1, zhi = vz_end;  // tile-space rows read (before clamp)
        const int lo = mesh_type == NZ_MESH_SQUARE_GRID ? 0 : -off, hi = mesh_type == NZ_MESH_SQUARE_GRID ? R + 1 : R + off;
        zlo = zlo < lo ? lo : zlo;
        zhi = zhi > hi ? hi : zhi;
        if (mesh_type == NZ_MESH_SQUARE_GRID && zhi > R) zhi = R;  // row R+1 is only reached through z+1 with z<R-1
        NZ_REQUIRE(zlo + off >= h_row_first && zhi + off < h_row_first + h_rows,
                   "mesh: resident height rows [%d,%d) do not cover rows [%d,%d] needed", h_row_first, h_row_first + h_rows,
                   zlo + off, zhi + off);
    }
    NZ_REQUIRE(vz_end - vz_begin <= 65535, "mesh: too many vertex rows in one launch");
    MeshParams p = {R, inRes, off, h_row_first, h_rows, vz_begin, vz_end, tile_height, tile_size};
    dim3 grid(cdiv(R + 1, MESH_TX), vz_end - vz_begin);
    if (mesh_type == NZ_MESH_SQUARE_GRID)
        mesh_kernel<NZ_MESH_SQUARE_GRID><<<grid, MESH_TX, 0, s>>>((float4*)d_vtx, (uint2*)d_idx, d_heights, p);
    else
        mesh_kernel<NZ_MESH_OVERSHOOT_SQUARE_GRID><<<grid, MESH_TX, 0, s>>>((float4*)d_vtx, (uint2*)d_idx, d_heights, p);
    NZ_LAUNCHED();
    return NZ_OK;
}

}  // namespace nz
