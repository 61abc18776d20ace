// NoizeB200Worlds.cs — P/Invoke bindings and thin owners for the two "whole job" entry points of libnoize_b200.so:
//
//   GpuTileWorld  : nz_tile_world_*  — the reference's tiled world (Scripts/MeshTileGenerator.cs:94-138,184-192 hands tile
//                   (tx, tz) to the generator pipeline with xpos = tileResolution * tx, zpos = tileResolution * tz, ONE tile in
//                   flight) with several tiles in flight on one GPU; one GpuTileWorld per device shards a world over the
//                   GPUs of the box with no communication (tiles are independent).
//   GpuBandChain  : nz_band_chain_create_local / _run / _download — one large heightmap (BASELINE.json configs[4]) split into
//                   row bands over the devices of THIS process, ghost rows moved by peer copies inside the library.
//
// Both are optional: the drop-in path is the stage classes of NoizeB200.cs (and nz_set_bands for large grids).  Struct
// layouts mirror include/noize_b200.h field for field (nz_tile_config, nz_chain_config, nz_band_info).
// Not compilable in the build image (no Unity / mono there); kept in step with the header by tests/test_abi_cpu.py, which
// checks that every function named in a DllImport here is exported by the library.
using System;
using System.Runtime.InteropServices;
using Unity.Collections;
using Unity.Collections.LowLevel.Unsafe;

namespace xshazwar.noize.interop.b200 {

    [StructLayout(LayoutKind.Sequential)]
    public struct NzTileConfig {
        public int resolution;             // generator resolution of one tile (e.g. 1024)
        public int tileResolution;         // world cells between tile origins (MeshTileGenerator.tileResolution)
        public int noiseType; public float hurst, startingAmplitude, stepdown, detuneRate; public int octaves, noiseSize;
        public int filterType, filterIterations;
        public int edgeFilterType, edgeFilterIterations;   // on a copy of the filtered tile; 0 iterations: none
        public int meshType, meshResolution, meshMarginPix; public float tileHeight, tileSize;   // meshResolution 0: no mesh
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct NzChainConfig {
        public int resolution;
        public int noiseType; public float hurst, startingAmplitude, stepdown, detuneRate; public int octaves, xpos, zpos, noiseSize;
        public int filterType, filterIterations;
        public int flowIterations; public float normMin, normMax;
        public int erosionIterations;
        public int meshType, meshResolution, meshMarginPix; public float tileHeight, tileSize;
    }

    [StructLayout(LayoutKind.Sequential)]
    public unsafe struct NzBandInfo {
        public int rank, world, device;
        public int z0, z1;                 // owned heightmap rows [z0, z1)
        public int vz0, vz1;               // owned vertex rows [vz0, vz1)
        public float* dRows;               // DEVICE pointers of the band's result
        public void* dVertices;
        public uint* dIndices;
        public long haloBytesPerRun;
    }

    public static unsafe class NativeWorlds {
        const string LIB = "noize_b200";

        public const int NZ_BANDS_EXCHANGE = 0, NZ_BANDS_RECOMPUTE = 1;

        [DllImport(LIB)] public static extern long nz_tile_world_create(NzTileConfig* cfg, int device, int slots);
        [DllImport(LIB)] public static extern int nz_tile_world_run(long world, int* tilesXZ, int n, float* hHeights, float* hEdges,
                                                                    void* hVertices, uint* hIndices);
        [DllImport(LIB)] public static extern int nz_tile_world_slot(long world, int slot, float** dHeights, float** dEdges,
                                                                     void** dVertices, uint** dIndices);
        [DllImport(LIB)] public static extern int nz_tile_world_destroy(long world);

        [DllImport(LIB)] public static extern int nz_band_geometry(int resolution, int world, int rank, int meshResolution, NzBandInfo* info);
        [DllImport(LIB)] public static extern long nz_band_chain_create_local(NzChainConfig* cfg, int* devices, int nBands, int mode);
        [DllImport(LIB)] public static extern int nz_band_chain_run(long chain, int timed);
        [DllImport(LIB)] public static extern int nz_band_chain_sync(long chain);
        [DllImport(LIB)] public static extern int nz_band_chain_stage_ms(long chain, float* ms5);
        [DllImport(LIB)] public static extern int nz_band_chain_local_bands(long chain);
        [DllImport(LIB)] public static extern int nz_band_chain_info(long chain, int localBand, NzBandInfo* info);
        [DllImport(LIB)] public static extern int nz_band_chain_download(long chain, float* hHeights, void* hVertices, uint* hIndices);
        [DllImport(LIB)] public static extern int nz_band_chain_destroy(long chain);
    }

    /// The tiles one GPU owns of a tiled world.  Run() blocks (call it from a worker thread or an IJob); the outputs land in the
    /// caller's NativeArrays tile after tile, in the PositionStream32 layout the reference's MeshTileStage produces.
    public sealed unsafe class GpuTileWorld : IDisposable {
        long handle;
        public readonly NzTileConfig config;

        public GpuTileWorld(NzTileConfig cfg, int device = 0, int slots = 4) {
            config = cfg;
            handle = NativeWorlds.nz_tile_world_create(&cfg, device, slots);
            Native.Check((int)Math.Min(handle, 0), "nz_tile_world_create");
        }

        /// tile owner rule shared with tiles.py: rank r of `world` devices owns the tiles with (tz * tilesX + tx) % world == r
        public static bool Owns(int tx, int tz, int tilesX, int rank, int world) => (tz * tilesX + tx) % world == rank;

        /// tilesXZ = {tx0, tz0, tx1, tz1, ...}; any output may be default (not created): it then stays on the device.
        public void Run(NativeArray<int> tilesXZ, NativeArray<float> heights, NativeArray<float> edges,
                        NativeArray<byte> vertices, NativeArray<uint> indices) {
            int n = tilesXZ.Length / 2;
            int rc = NativeWorlds.nz_tile_world_run(handle, (int*)tilesXZ.GetUnsafeReadOnlyPtr(), n,
                heights.IsCreated ? (float*)heights.GetUnsafePtr() : null,
                edges.IsCreated ? (float*)edges.GetUnsafePtr() : null,
                vertices.IsCreated ? vertices.GetUnsafePtr() : null,
                indices.IsCreated ? (uint*)indices.GetUnsafePtr() : null);
            Native.Check(rc, "noize_b200 world call");
        }

        public void Dispose() {
            if (handle > 0) NativeWorlds.nz_tile_world_destroy(handle);
            handle = 0;
        }
    }

    /// One large heightmap on the row bands of several devices of this process.
    public sealed unsafe class GpuBandChain : IDisposable {
        long handle;

        public GpuBandChain(NzChainConfig cfg, int[] devices, int mode = NativeWorlds.NZ_BANDS_EXCHANGE) {
            fixed (int* d = devices) handle = NativeWorlds.nz_band_chain_create_local(&cfg, d, devices.Length, mode);
            Native.Check((int)Math.Min(handle, 0), "nz_band_chain_create_local");
        }

        /// enqueue one pass (asynchronous); Download() or Sync() waits for it
        public void Run(bool timed = false) {
            int rc = NativeWorlds.nz_band_chain_run(handle, timed ? 1 : 0);
            Native.Check(rc, "noize_b200 world call");
        }

        public void Sync() {
            int rc = NativeWorlds.nz_band_chain_sync(handle);
            Native.Check(rc, "noize_b200 world call");
        }

        /// {noise, filter, flow, erosion, mesh} milliseconds of the last timed run
        public float[] StageMs() {
            var ms = new float[5];
            fixed (float* p = ms) {
                int rc = NativeWorlds.nz_band_chain_stage_ms(handle, p);
                Native.Check(rc, "noize_b200 world call");
            }
            return ms;
        }

        /// full-grid host buffers; every band writes its own slice
        public void Download(NativeArray<float> heights, NativeArray<byte> vertices, NativeArray<uint> indices) {
            int rc = NativeWorlds.nz_band_chain_download(handle,
                heights.IsCreated ? (float*)heights.GetUnsafePtr() : null,
                vertices.IsCreated ? vertices.GetUnsafePtr() : null,
                indices.IsCreated ? (uint*)indices.GetUnsafePtr() : null);
            Native.Check(rc, "noize_b200 world call");
        }

        public void Dispose() {
            if (handle > 0) NativeWorlds.nz_band_chain_destroy(handle);
            handle = 0;
        }
    }
}
