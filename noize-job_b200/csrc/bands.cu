// bands.cu — one large heightmap as ROW BANDS over several GPUs, behind the C ABI (include/noize_b200.h, "multi-GPU").
//
// Band b of g owns rows [b*n/g, (b+1)*n/g) of the grid.  Every stage kernel of this library works on a rectangular window
// and a window border that is not a border of the full grid only corrupts r rows per stencil iteration, so a band runs the
// ORDINARY single-GPU kernels on (own rows + ghost rows) and keeps its own rows — bit-identical to one GPU:
//     noise            0 ghost rows     (pure function of position: no communication)
//     filter K x I     r*I above/below  (Pipeline/Tiles/TileData.cs:72-77 clamps only at the grid's own border)
//     flow map x I     2I+1 above/below (outflow + water step per iteration, + the velocity epilogue)
//     value erosion    I above          (trailing window)
//     mesh             1 above/below    (normals read z-1, z+1)
// The reference has no counterpart: it runs one tile at a time on the CPU (Scripts/MeshTileGenerator.cs:125-138).
//
// A BandSet is the set of bands THIS process drives.  Ghost rows travel
//   * between bands of one process by cudaMemcpyPeerAsync (NVLink peer copies; devices may even repeat), ordered by events;
//   * between processes by ncclSend/ncclRecv in one group per stage.  libnccl.so.2 is bound with dlopen at first use, so
//     the library itself links nothing but the CUDA runtime (inside a torch process the already-loaded NCCL is picked up).
// One exchange per stage, never per iteration.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>
#include <nvtx3/nvToolsExt.h>
#include "nz_common.cuh"
#include "bands.cuh"

namespace nz {

// =====================================================================================================================
// NCCL, bound at run time
// =====================================================================================================================
namespace {

struct NcclId { char internal[NZ_COMM_ID_BYTES]; };
struct Nccl {
    void* handle = nullptr;
    int (*GetVersion)(int*) = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*CommGetAsyncError)(void*, int*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
};
constexpr int NCCL_FLOAT32 = 7;   // ncclFloat32 (nccl.h; stable since NCCL 2.0)

std::mutex g_nccl_mu;
Nccl g_nccl;

int32_t nccl_load(Nccl** out) {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (!g_nccl.handle) {
        const char* names[] = {getenv("NZ_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        void* h = nullptr;
        for (const char* n : names)
            if (n && n[0] && (h = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
        if (!h) {
            set_error("NCCL is not available (dlopen libnccl.so.2: %s); bands in separate processes need it", dlerror());
            return NZ_E_UNSUPPORTED;
        }
        Nccl n;
        n.handle = h;
#define NZ_SYM(field, name)                                                        \
    do {                                                                           \
        *(void**)(&n.field) = dlsym(h, name);                                      \
        if (!n.field) {                                                            \
            set_error("libnccl lacks %s", name);                                   \
            dlclose(h);                                                            \
            return NZ_E_UNSUPPORTED;                                               \
        }                                                                          \
    } while (0)
        NZ_SYM(GetVersion, "ncclGetVersion");
        NZ_SYM(GetUniqueId, "ncclGetUniqueId");
        NZ_SYM(CommInitRank, "ncclCommInitRank");
        NZ_SYM(CommDestroy, "ncclCommDestroy");
        NZ_SYM(CommGetAsyncError, "ncclCommGetAsyncError");
        NZ_SYM(GetErrorString, "ncclGetErrorString");
        NZ_SYM(Send, "ncclSend");
        NZ_SYM(Recv, "ncclRecv");
        NZ_SYM(GroupStart, "ncclGroupStart");
        NZ_SYM(GroupEnd, "ncclGroupEnd");
#undef NZ_SYM
        g_nccl = n;
    }
    *out = &g_nccl;
    return NZ_OK;
}

#define NZ_NCCL(api, call)                                                                   \
    do {                                                                                     \
        int _r = (call);                                                                     \
        if (_r != 0) {                                                                       \
            set_error("NCCL error %d (%s) in %s", _r, (api)->GetErrorString(_r), #call);      \
            return NZ_E_CUDA;                                                                \
        }                                                                                    \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (dev != prev) cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

struct Range {   // NVTX range per stage: visible in nsys / ncu timelines, free when no tool is attached
    explicit Range(const char* name) { nvtxRangePushA(name); }
    ~Range() { nvtxRangePop(); }
};

}  // namespace

struct Comm {
    void* nccl = nullptr;
    int world = 1, rank = 0, device = 0;
};

// =====================================================================================================================
// BandSet
// =====================================================================================================================
void band_rows(int rows, int world, int rank, int* z0, int* z1) {
    *z0 = (int)((long long)rank * rows / world);
    *z1 = (int)((long long)(rank + 1) * rows / world);
}

BandSet::~BandSet() {
    for (Band& bd : b) {
        DeviceGuard g(bd.device);
        if (bd.s) cudaStreamSynchronize(bd.s);
        for (float*& p : bd.buf) {
            if (p) dev_free(p);
            p = nullptr;
        }
        if (bd.scratch) dev_free(bd.scratch);
        if (bd.vtx) dev_free(bd.vtx);
        if (bd.idx) dev_free(bd.idx);
        if (bd.done) cudaEventDestroy(bd.done);
        if (bd.pulled) cudaEventDestroy(bd.pulled);
        for (cudaEvent_t e : bd.ev)
            if (e) cudaEventDestroy(e);
        if (bd.s && bd.own_stream) cudaStreamDestroy(bd.s);
    }
}

// Creates the bands [first_rank, first_rank + n_local) of a `world`-band partition of a width x rows grid, band k on
// devices[k].  cap = ghost rows kept above and below every band (clipped at the grid border).
int32_t bandset_create(std::unique_ptr<BandSet>* out, int width, int rows, int world, int first_rank, int n_local,
                       const int* devices, int cap, Comm* comm, cudaStream_t stream) {
    NZ_REQUIRE(width > 0 && rows > 0 && world >= 1 && n_local >= 1 && first_rank >= 0 && first_rank + n_local <= world,
               "bands: bad partition (%d x %d, world %d, local [%d, %d))", width, rows, world, first_rank, first_rank + n_local);
    NZ_REQUIRE(n_local == world || (n_local == 1 && comm), "bands: a process drives either every band or one band of a communicator");
    int ndev = 0;
    NZ_CUDA(cudaGetDeviceCount(&ndev));
    std::unique_ptr<BandSet> bs(new BandSet);
    bs->width = width;
    bs->rows = rows;
    bs->world = world;
    bs->comm = comm;
    int min_own = rows;
    for (int r = 0; r < world; r++) {
        int z0, z1;
        band_rows(rows, world, r, &z0, &z1);
        if (z1 - z0 < min_own) min_own = z1 - z0;
    }
    NZ_REQUIRE(min_own >= 1, "bands: %d rows cannot be split into %d bands", rows, world);
    bs->cap = world > 1 ? (cap < min_own ? cap : min_own) : 0;
    bs->b.resize(n_local);
    for (int k = 0; k < n_local; k++) {
        Band& bd = bs->b[k];
        bd.device = devices[k];
        NZ_REQUIRE(bd.device >= 0 && bd.device < ndev, "bands: device %d not visible (%d devices)", bd.device, ndev);
        bd.rank = first_rank + k;
        band_rows(rows, world, bd.rank, &bd.z0, &bd.z1);
        bd.own = bd.z1 - bd.z0;
        bd.above = bs->cap < bd.z0 ? bs->cap : bd.z0;
        bd.below = bs->cap < rows - bd.z1 ? bs->cap : rows - bd.z1;
        DeviceGuard g(bd.device);
        const size_t bytes = (size_t)(bd.above + bd.own + bd.below) * width * sizeof(float);
        for (float*& p : bd.buf) {
            int32_t rc = dev_alloc((void**)&p, bytes);
            if (rc != NZ_OK) return rc;
        }
        if (stream != (cudaStream_t)NZ_STREAM_OWN && n_local == 1) {
            bd.s = stream;              // the caller's stream (NULL: the legacy default stream)
        } else {
            NZ_CUDA(cudaStreamCreateWithFlags(&bd.s, cudaStreamNonBlocking));
            bd.own_stream = true;
        }
        NZ_CUDA(cudaEventCreateWithFlags(&bd.done, cudaEventDisableTiming));
        NZ_CUDA(cudaEventCreateWithFlags(&bd.pulled, cudaEventDisableTiming));
    }
    // peer access between neighbouring bands of this process (without it the copies are staged through the host)
    for (int k = 0; k + 1 < n_local; k++) {
        const int a = bs->b[k].device, c = bs->b[k + 1].device;
        if (a == c) continue;
        int ok = 0;
        if (cudaDeviceCanAccessPeer(&ok, a, c) == cudaSuccess && ok) {
            DeviceGuard g(a);
            if (cudaDeviceEnablePeerAccess(c, 0) != cudaSuccess) cudaGetLastError();   // already enabled
        }
        if (cudaDeviceCanAccessPeer(&ok, c, a) == cudaSuccess && ok) {
            DeviceGuard g(c);
            if (cudaDeviceEnablePeerAccess(a, 0) != cudaSuccess) cudaGetLastError();
        }
    }
    *out = std::move(bs);
    return NZ_OK;
}

// the window of `bd` covering its own rows plus (above, below) ghost rows, clipped at the grid border
static inline void band_window(const Band& bd, int rows_total, int above, int below, size_t width, int* row0, int* nrows, int* a_out) {
    const int a = above < bd.z0 ? above : bd.z0;
    const int c = below < rows_total - bd.z1 ? below : rows_total - bd.z1;
    *row0 = bd.above - a;
    *nrows = a + bd.own + c;
    *a_out = a;
    (void)width;
}

// Before a band overwrites one of its buffers, its neighbours must have finished reading their ghost rows out of it.
static int32_t wait_neighbour_pulls(BandSet& bs, int k) {
    if (bs.b.size() < 2) return NZ_OK;
    Band& bd = bs.b[k];
    if (k > 0) NZ_CUDA(cudaStreamWaitEvent(bd.s, bs.b[k - 1].pulled, 0));
    if (k + 1 < (int)bs.b.size()) NZ_CUDA(cudaStreamWaitEvent(bd.s, bs.b[k + 1].pulled, 0));
    return NZ_OK;
}

// Fill `above` ghost rows over and `below` ghost rows under the own rows of every band's current buffer from the
// neighbours' own rows.  Row counts are the same for every band (symmetric), at most `cap`.
int32_t bandset_exchange(BandSet& bs, int above, int below) {
    if (bs.world == 1 || (above == 0 && below == 0)) return NZ_OK;
    NZ_REQUIRE(above <= bs.cap && below <= bs.cap, "bands: %d/%d ghost rows exceed the capacity %d", above, below, bs.cap);
    Range r("nz.halo_exchange");
    const size_t W = bs.width, row_bytes = W * sizeof(float);
    if (bs.comm) {
        // one band per process: ncclSend/ncclRecv with both neighbours in ONE group on the band's stream
        Nccl* api;
        int32_t rc = nccl_load(&api);
        if (rc != NZ_OK) return rc;
        Band& bd = bs.b[0];
        DeviceGuard g(bd.device);
        float* cur = bd.buf[bd.cur];
        const int top = bd.above, bot = bd.above + bd.own;
        const int up = bd.rank - 1, down = bd.rank + 1;
        NZ_NCCL(api, api->GroupStart());
        if (up >= 0) {
            if (below > 0) NZ_NCCL(api, api->Send(cur + (size_t)top * W, (size_t)below * W, NCCL_FLOAT32, up, bs.comm->nccl, bd.s));
            if (above > 0) NZ_NCCL(api, api->Recv(cur + (size_t)(top - above) * W, (size_t)above * W, NCCL_FLOAT32, up, bs.comm->nccl, bd.s));
        }
        if (down < bs.world) {
            if (above > 0) NZ_NCCL(api, api->Send(cur + (size_t)(bot - above) * W, (size_t)above * W, NCCL_FLOAT32, down, bs.comm->nccl, bd.s));
            if (below > 0) NZ_NCCL(api, api->Recv(cur + (size_t)bot * W, (size_t)below * W, NCCL_FLOAT32, down, bs.comm->nccl, bd.s));
        }
        NZ_NCCL(api, api->GroupEnd());
        bd.halo_bytes += ((up >= 0 ? above : 0) + (down < bs.world ? below : 0)) * (long long)row_bytes;
        return NZ_OK;
    }
    // every band in this process: each band PULLS its ghost rows out of its neighbours' buffers once they are done
    const int n = (int)bs.b.size();
    for (int k = 0; k < n; k++) {
        Band& bd = bs.b[k];
        DeviceGuard g(bd.device);
        float* cur = bd.buf[bd.cur];
        if (k > 0) {
            Band& up = bs.b[k - 1];
            NZ_CUDA(cudaStreamWaitEvent(bd.s, up.done, 0));
            if (above > 0) {
                const float* src = up.buf[up.cur] + (size_t)(up.above + up.own - above) * W;
                NZ_CUDA(cudaMemcpyPeerAsync(cur + (size_t)(bd.above - above) * W, bd.device, src, up.device, above * row_bytes, bd.s));
                bd.halo_bytes += above * (long long)row_bytes;
            }
        }
        if (k + 1 < n) {
            Band& dn = bs.b[k + 1];
            NZ_CUDA(cudaStreamWaitEvent(bd.s, dn.done, 0));
            if (below > 0) {
                const float* src = dn.buf[dn.cur] + (size_t)dn.above * W;
                NZ_CUDA(cudaMemcpyPeerAsync(cur + (size_t)(bd.above + bd.own) * W, bd.device, src, dn.device, below * row_bytes, bd.s));
                bd.halo_bytes += below * (long long)row_bytes;
            }
        }
        NZ_CUDA(cudaEventRecord(bd.pulled, bd.s));
    }
    return NZ_OK;
}

// Runs `fn` on every band's window (own rows + (above, below) ghost rows).  fn(band, cur, other, rows, row_first, &result)
// enqueues the stage on band.s and reports which of the two window pointers holds the result.
int32_t bandset_stage(BandSet& bs, int above, int below, const BandStageFn& fn) {
    const size_t W = bs.width;
    for (int k = 0; k < (int)bs.b.size(); k++) {
        Band& bd = bs.b[k];
        DeviceGuard g(bd.device);
        int32_t rc = wait_neighbour_pulls(bs, k);
        if (rc != NZ_OK) return rc;
        int row0, nrows, a;
        band_window(bd, bs.rows, above, below, W, &row0, &nrows, &a);
        float* cur = bd.buf[bd.cur] + (size_t)row0 * W;
        float* other = bd.buf[bd.cur ^ 1] + (size_t)row0 * W;
        float* result = cur;
        const int first = bd.z0 - a;
        GridEdgesScope edges((first == 0 ? 1 : 0) | (first + nrows == bs.rows ? 2 : 0));
        rc = fn(bd, cur, other, nrows, first, &result);
        if (rc != NZ_OK) return rc;
        if (result == other) bd.cur ^= 1;
        NZ_CUDA(cudaEventRecord(bd.done, bd.s));
    }
    return NZ_OK;
}

int32_t bandset_sync(BandSet& bs) {
    for (Band& bd : bs.b) {
        DeviceGuard g(bd.device);
        NZ_CUDA(cudaStreamSynchronize(bd.s));
    }
    return NZ_OK;
}

// scratch of at least `bytes` on the band's device (grown on demand, reused across stages)
int32_t band_scratch(Band& bd, size_t bytes, void** out) {
    if (bytes > bd.scratch_bytes) {
        if (bd.scratch) {
            NZ_CUDA(cudaStreamSynchronize(bd.s));
            dev_free(bd.scratch);
            bd.scratch = nullptr;
            bd.scratch_bytes = 0;
        }
        int32_t rc = dev_alloc(&bd.scratch, bytes);
        if (rc != NZ_OK) return rc;
        bd.scratch_bytes = bytes;
    }
    *out = bytes ? bd.scratch : nullptr;
    return NZ_OK;
}

// vertex rows whose height row (vz + off) the band owns; the first / last band also take the margin rows
void band_vertex_rows(const Band& bd, int world, int in_res, int R, int* vz0, int* vz1) {
    const int off = (in_res - R) / 2;
    int lo = bd.rank > 0 ? bd.z0 - off : 0;
    int hi = bd.rank < world - 1 ? bd.z1 - off : R + 1;
    if (lo < 0) lo = 0;
    if (hi > R + 1) hi = R + 1;
    if (hi < lo) hi = lo;
    *vz0 = lo;
    *vz1 = hi;
}

// Mesh slice of every band (1 ghost row each side must be current): vertex rows [vz0, vz1) and the triangle rows they close.
int32_t bandset_mesh(BandSet& bs, int mesh_type, int R, float tile_height, float tile_size) {
    const size_t W = bs.width;
    for (int k = 0; k < (int)bs.b.size(); k++) {
        Band& bd = bs.b[k];
        DeviceGuard g(bd.device);
        band_vertex_rows(bd, bs.world, bs.rows, R, &bd.vz0, &bd.vz1);
        const int t0 = bd.vz0 > 1 ? bd.vz0 : 1;
        const size_t vbytes = (size_t)(bd.vz1 - bd.vz0) * (R + 1) * NZ_MESH_VERTEX_BYTES;
        const size_t ibytes = (size_t)6 * R * (bd.vz1 > t0 ? bd.vz1 - t0 : 0) * sizeof(uint32_t);
        if (vbytes > bd.vtx_bytes) {
            if (bd.vtx) { NZ_CUDA(cudaStreamSynchronize(bd.s)); dev_free(bd.vtx); bd.vtx = nullptr; }
            int32_t rc = dev_alloc(&bd.vtx, vbytes);
            if (rc != NZ_OK) return rc;
            bd.vtx_bytes = vbytes;
        }
        if (ibytes > bd.idx_bytes) {
            if (bd.idx) { NZ_CUDA(cudaStreamSynchronize(bd.s)); dev_free(bd.idx); bd.idx = nullptr; }
            int32_t rc = dev_alloc((void**)&bd.idx, ibytes);
            if (rc != NZ_OK) return rc;
            bd.idx_bytes = ibytes;
        }
        if (bd.vz1 <= bd.vz0) continue;
        int row0, nrows, a;
        band_window(bd, bs.rows, 1, 1, W, &row0, &nrows, &a);
        const float* win = bd.buf[bd.cur] + (size_t)row0 * W;
        int32_t rc = launch_mesh(mesh_type, bd.vtx, bd.idx, R, bs.rows, tile_height, tile_size, win, bd.z0 - a, nrows, bd.vz0,
                                 bd.vz1, bd.s);
        if (rc != NZ_OK) return rc;
        NZ_CUDA(cudaEventRecord(bd.done, bd.s));
    }
    return NZ_OK;
}

// ---- stage wrappers shared by the chain below and the banded host layer (abi.cu) ----------------------------------
int32_t bandset_fractal(BandSet& bs, int noise_type, float hurst, float start_amp, float stepdown, float detune, int octaves,
                        int xpos, int zpos, int noise_size, int extra_above, int extra_below) {
    Range r("nz.noise");
    const int W = bs.width;
    for (int k = 0; k < (int)bs.b.size(); k++) {
        Band& bd = bs.b[k];
        DeviceGuard g(bd.device);
        int32_t rc = wait_neighbour_pulls(bs, k);
        if (rc != NZ_OK) return rc;
        int row0, nrows, a;
        band_window(bd, bs.rows, extra_above, extra_below, W, &row0, &nrows, &a);
        FractalParams p;
        rc = fractal_params(&p, W, nrows, bd.z0 - a, noise_type, hurst, start_amp, stepdown, detune, octaves, xpos, zpos, noise_size);
        if (rc != NZ_OK) return rc;
        rc = launch_fractal(bd.buf[bd.cur] + (size_t)row0 * W, noise_type, p, bd.s);
        if (rc != NZ_OK) return rc;
        NZ_CUDA(cudaEventRecord(bd.done, bd.s));
    }
    return NZ_OK;
}

int32_t bandset_separable(BandSet& bs, int ksize, const float* kx, const float* kz, float factor, int iterations,
                          int extra_above, int extra_below, bool exchange) {
    Range rg("nz.filter");
    const int r = ksize / 2, W = bs.width;
    // rounds: as many iterations per exchange as the ghost capacity allows (all of them for the configs of BASELINE.json)
    const int per_round = (bs.world == 1 || !exchange) ? iterations : (r == 0 ? iterations : bs.cap / r);
    NZ_REQUIRE(per_round >= 1 || iterations == 0, "bands: kernel radius %d exceeds the band height %d", r, bs.cap);
    int left = iterations;
    while (left > 0) {
        const int it = left < per_round ? left : per_round;
        const int h = r * it;
        int32_t rc = exchange ? bandset_exchange(bs, h, h) : NZ_OK;
        if (rc != NZ_OK) return rc;
        rc = bandset_stage(bs, h + extra_above, h + extra_below, [&](Band& bd, float* cur, float* other, int rows, int, float** res) {
            return launch_separable(cur, other, W, rows, ksize, kx, kz, factor, it, res, bd.s);
        });
        if (rc != NZ_OK) return rc;
        left -= it;
    }
    return NZ_OK;
}

int32_t bandset_sobel2d(BandSet& bs, int iterations, int extra_above, int extra_below, bool exchange) {
    Range rg("nz.sobel2d");
    const int W = bs.width;
    const int per_round = (bs.world == 1 || !exchange) ? iterations : bs.cap;
    int left = iterations;
    while (left > 0) {
        const int it = left < per_round ? left : per_round;
        int32_t rc = exchange ? bandset_exchange(bs, it, it) : NZ_OK;
        if (rc != NZ_OK) return rc;
        rc = bandset_stage(bs, it + extra_above, it + extra_below, [&](Band& bd, float* cur, float* other, int rows, int, float** res) {
            return launch_sobel2d(cur, other, W, rows, it, res, bd.s);
        });
        if (rc != NZ_OK) return rc;
        left -= it;
    }
    return NZ_OK;
}

int32_t bandset_min_erosion(BandSet& bs, int iterations, int extra_above, int extra_below, bool exchange) {
    Range rg("nz.value_erosion");
    const int W = bs.width;
    const int per_round = (bs.world == 1 || !exchange) ? iterations : bs.cap;
    int left = iterations;
    while (left > 0) {
        const int it = left < per_round ? left : per_round;
        int32_t rc = exchange ? bandset_exchange(bs, it, 0) : NZ_OK;
        if (rc != NZ_OK) return rc;
        rc = bandset_stage(bs, it + extra_above, extra_below, [&](Band& bd, float* cur, float* other, int rows, int, float** res) {
            return launch_min_erosion(cur, other, W, rows, it, res, bd.s);
        });
        if (rc != NZ_OK) return rc;
        left -= it;
    }
    return NZ_OK;
}

// The flow map cannot be split into rounds (water and flows live inside the stage): it needs 2*iterations+1 ghost rows at once.
bool bandset_flowmap_fits(const BandSet& bs, int iterations) { return bs.world == 1 || 2 * iterations + 1 <= bs.cap; }

int32_t bandset_flowmap(BandSet& bs, int iterations, float norm_min, float norm_max, int extra_above, int extra_below, bool exchange) {
    Range rg("nz.flowmap");
    const int W = bs.width, h = 2 * iterations + 1;
    NZ_REQUIRE(!exchange || bandset_flowmap_fits(bs, iterations), "bands: the flow map needs %d ghost rows, the bands hold %d", h, bs.cap);
    int32_t rc = exchange ? bandset_exchange(bs, h, h) : NZ_OK;
    if (rc != NZ_OK) return rc;
    return bandset_stage(bs, h + extra_above, h + extra_below, [&](Band& bd, float* cur, float* other, int rows, int, float** res) {
        void* scratch = nullptr;
        int32_t r2 = band_scratch(bd, flowmap_scratch_bytes(W, rows, iterations), &scratch);
        if (r2 != NZ_OK) return r2;
        return launch_flowmap(cur, other, scratch, W, rows, iterations, norm_min, norm_max, res, bd.s);
    });
}

// =====================================================================================================================
// the chain of BASELINE.json configs[4] on a BandSet
// =====================================================================================================================
namespace {

struct Chain {
    nz_chain_config cfg{};
    int mode = NZ_BANDS_EXCHANGE;
    std::unique_ptr<BandSet> bs;
    std::shared_ptr<Comm> comm;      // keeps the communicator alive as long as the chain
    float kx[NZ_MAX_KERNEL_WIDTH], kz[NZ_MAX_KERNEL_WIDTH], factor = 1.0f;
    int ksize = 0;       // 0: Sobel3_2D (not one separable kernel)
    int h_filter = 0, h_flow = 0, h_erosion = 0, h_mesh = 0;   // ghost rows per stage (above; below is the same except erosion: 0)
    bool timed = false;
    long long runs = 0;
    std::mutex mu;
};

std::mutex g_handles_mu;
std::unordered_map<long long, std::shared_ptr<Chain>> g_chains;
std::unordered_map<long long, std::shared_ptr<Comm>> g_comms;
long long g_next_handle = 1;

int32_t chain_setup(Chain& c, const nz_chain_config* cfg, int mode) {
    NZ_REQUIRE(cfg, "band chain: null config");
    NZ_REQUIRE(mode == NZ_BANDS_EXCHANGE || mode == NZ_BANDS_RECOMPUTE, "band chain: mode %d out of range", mode);
    NZ_REQUIRE(cfg->resolution > 0 && cfg->resolution <= 46340, "band chain: resolution %d out of range", cfg->resolution);
    NZ_REQUIRE(cfg->filter_iterations >= 0 && cfg->flow_iterations >= 0 && cfg->erosion_iterations >= 0, "band chain: negative iteration count");
    c.cfg = *cfg;
    c.mode = mode;
    if (cfg->filter_iterations > 0) {
        NZ_REQUIRE(cfg->filter_type >= 0 && cfg->filter_type < NZ_FILTER__COUNT, "band chain: filter_type %d out of range", cfg->filter_type);
        if (cfg->filter_type == NZ_FILTER_SOBEL3_2D) {
            c.ksize = 0;
            c.h_filter = cfg->filter_iterations;
        } else {
            int32_t rc = kernel_filter_table(cfg->filter_type, c.kx, c.kz, &c.ksize, &c.factor);
            if (rc != NZ_OK) return rc;
            c.h_filter = (c.ksize / 2) * cfg->filter_iterations;
        }
    }
    c.h_flow = cfg->flow_iterations > 0 ? 2 * cfg->flow_iterations + 1 : 0;
    c.h_erosion = cfg->erosion_iterations;
    if (cfg->mesh_resolution > 0) {
        NZ_REQUIRE(cfg->mesh_type >= 0 && cfg->mesh_type < NZ_MESH__COUNT, "band chain: mesh_type %d out of range", cfg->mesh_type);
        c.h_mesh = 1;       // launch_mesh validates resolution against input resolution
    }
    return NZ_OK;
}

// ghost rows a band holds on either side: everything the chain consumes below the noise stage (see chain_run)
int chain_cap(const Chain& c) { return c.h_filter + c.h_flow + c.h_erosion + c.h_mesh; }

int32_t chain_check_fit(const Chain& c) {
    const BandSet& bs = *c.bs;
    if (bs.world == 1) return NZ_OK;
    const int need = chain_cap(c);
    NZ_REQUIRE(bs.cap >= need, "band chain: bands of %d rows are smaller than the ghost zone (%d rows) of this chain at %d bands",
               bs.cap, need, bs.world);
    return NZ_OK;
}

int32_t chain_mark(Chain& c, int stage) {
    if (!c.timed) return NZ_OK;
    for (Band& bd : c.bs->b) {
        DeviceGuard g(bd.device);
        if (!bd.ev[stage]) NZ_CUDA(cudaEventCreate(&bd.ev[stage]));
        NZ_CUDA(cudaEventRecord(bd.ev[stage], bd.s));
    }
    return NZ_OK;
}

int32_t chain_run(Chain& c, bool timed) {
    const nz_chain_config& f = c.cfg;
    BandSet& bs = *c.bs;
    const bool exch = c.mode == NZ_BANDS_EXCHANGE;
    c.timed = timed;
    // Both modes run every stage below the noise on SHRINKING windows: a stage with r ghost rows of stencil reach leaves the
    // outer r rows of its window wrong, which is fine as long as the stages after it need that many rows less.  Ghost rows
    // still needed AFTER the filter (above / below): the flow map's 2I+1, the value erosion's (above only), the mesh's 1.
    const int rem_filter_a = c.h_flow + c.h_erosion + c.h_mesh, rem_filter_b = c.h_flow + c.h_mesh;
    const int noise_a = c.h_filter + rem_filter_a, noise_b = c.h_filter + rem_filter_b;
    int32_t rc;
    if ((rc = chain_mark(c, 0)) != NZ_OK) return rc;
    // recompute mode: the noise stage evaluates the ghost rows itself (noise is a pure function of position)
    rc = bandset_fractal(bs, f.noise_type, f.hurst, f.starting_amplitude, f.stepdown, f.detune_rate, f.octaves, f.xpos, f.zpos,
                         f.noise_size, exch ? 0 : noise_a, exch ? 0 : noise_b);
    if (rc != NZ_OK) return rc;
    // Exchange mode moves ghost rows ONCE per pass: the noise rows everything downstream needs (51 above / 46 below at the
    // C5 parameters, 3.3 MB each way).  Round 2 first exchanged twice (34 rows before the filter, 17 / 12 of filtered rows
    // before the flow map) and round 1 once per stage; every NCCL round trip costs ~40 us on a band that takes 2.3 ms, the
    // 17 extra rows through the filter 6 us.
    if (exch && (rc = bandset_exchange(bs, noise_a, noise_b)) != NZ_OK) return rc;
    if ((rc = chain_mark(c, 1)) != NZ_OK) return rc;
    if (f.filter_iterations > 0) {
        rc = c.ksize ? bandset_separable(bs, c.ksize, c.kx, c.kz, c.factor, f.filter_iterations, rem_filter_a, rem_filter_b, false)
                     : bandset_sobel2d(bs, f.filter_iterations, rem_filter_a, rem_filter_b, false);
        if (rc != NZ_OK) return rc;
    }
    if ((rc = chain_mark(c, 2)) != NZ_OK) return rc;
    const int xfa = c.h_erosion + c.h_mesh, xfb = c.h_mesh, xea = c.h_mesh, xeb = c.h_mesh;
    if (f.flow_iterations > 0) {
        rc = bandset_flowmap(bs, f.flow_iterations, f.norm_min, f.norm_max, xfa, xfb, false);
        if (rc != NZ_OK) return rc;
    }
    if ((rc = chain_mark(c, 3)) != NZ_OK) return rc;
    if (f.erosion_iterations > 0) {
        rc = bandset_min_erosion(bs, f.erosion_iterations, xea, xeb, false);
        if (rc != NZ_OK) return rc;
    }
    if ((rc = chain_mark(c, 4)) != NZ_OK) return rc;
    if (f.mesh_resolution > 0) {
        Range rg("nz.mesh");
        rc = bandset_mesh(bs, f.mesh_type, f.mesh_resolution, f.tile_height, f.tile_size);
        if (rc != NZ_OK) return rc;
    }
    if ((rc = chain_mark(c, 5)) != NZ_OK) return rc;
    c.runs++;
    return NZ_OK;
}

int32_t find_chain(long long h, std::shared_ptr<Chain>* out) {
    std::lock_guard<std::mutex> lk(g_handles_mu);
    auto it = g_chains.find(h);
    if (it == g_chains.end()) {
        set_error("unknown band chain handle %lld", h);
        return NZ_E_INVALID;
    }
    *out = it->second;
    return NZ_OK;
}

}  // namespace
}  // namespace nz

using namespace nz;

extern "C" {

NZ_API int32_t nz_band_geometry(int32_t resolution, int32_t world, int32_t rank, int32_t mesh_resolution, nz_band_info* out) {
    NZ_REQUIRE(out, "nz_band_geometry: null output");
    NZ_REQUIRE(resolution > 0 && world >= 1 && world <= resolution && rank >= 0 && rank < world && mesh_resolution >= 0,
               "nz_band_geometry: bad partition (resolution %d, world %d, rank %d)", resolution, world, rank);
    memset(out, 0, sizeof(*out));
    Band bd;
    bd.rank = rank;
    band_rows(resolution, world, rank, &bd.z0, &bd.z1);
    out->rank = rank;
    out->world = world;
    out->z0 = bd.z0;
    out->z1 = bd.z1;
    if (mesh_resolution > 0) band_vertex_rows(bd, world, resolution, mesh_resolution, &out->vz0, &out->vz1);
    return NZ_OK;
}

NZ_API int32_t nz_comm_unique_id(void* id_bytes, int32_t capacity) {
    NZ_REQUIRE(id_bytes && capacity >= NZ_COMM_ID_BYTES, "nz_comm_unique_id: the buffer must hold %d bytes", NZ_COMM_ID_BYTES);
    Nccl* api;
    int32_t rc = nccl_load(&api);
    if (rc != NZ_OK) return rc;
    NcclId id;
    NZ_NCCL(api, api->GetUniqueId(&id));
    memcpy(id_bytes, id.internal, NZ_COMM_ID_BYTES);
    return NZ_OK;
}

NZ_API int64_t nz_comm_create(const void* id_bytes, int32_t world, int32_t rank, int32_t device) {
    NZ_REQUIRE(id_bytes && world >= 1 && rank >= 0 && rank < world, "nz_comm_create: bad arguments (world %d, rank %d)", world, rank);
    Nccl* api;
    int32_t rc = nccl_load(&api);
    if (rc != NZ_OK) return rc;
    int ndev = 0;
    NZ_CUDA(cudaGetDeviceCount(&ndev));
    NZ_REQUIRE(device >= 0 && device < ndev, "nz_comm_create: device %d not visible (%d devices)", device, ndev);
    auto c = std::make_shared<Comm>();
    c->world = world;
    c->rank = rank;
    c->device = device;
    NcclId id;
    memcpy(id.internal, id_bytes, NZ_COMM_ID_BYTES);
    {
        DeviceGuard g(device);
        NZ_NCCL(api, api->CommInitRank(&c->nccl, world, id, rank));
    }
    std::lock_guard<std::mutex> lk(g_handles_mu);
    const long long h = g_next_handle++;
    g_comms[h] = c;
    return h;
}

NZ_API int32_t nz_comm_async_error(int64_t comm) {
    std::shared_ptr<Comm> c;
    {
        std::lock_guard<std::mutex> lk(g_handles_mu);
        auto it = g_comms.find(comm);
        NZ_REQUIRE(it != g_comms.end(), "nz_comm_async_error: unknown communicator %lld", (long long)comm);
        c = it->second;
    }
    Nccl* api;
    int32_t rc = nccl_load(&api);
    if (rc != NZ_OK) return rc;
    int err = 0;
    NZ_NCCL(api, api->CommGetAsyncError(c->nccl, &err));
    if (err != 0 && err != 7 /* ncclInProgress */) {
        set_error("NCCL asynchronous error %d (%s) on rank %d of %d", err, api->GetErrorString(err), c->rank, c->world);
        return NZ_E_CUDA;
    }
    return NZ_OK;
}

NZ_API int32_t nz_comm_destroy(int64_t comm) {
    std::shared_ptr<Comm> c;
    {
        std::lock_guard<std::mutex> lk(g_handles_mu);
        auto it = g_comms.find(comm);
        NZ_REQUIRE(it != g_comms.end(), "nz_comm_destroy: unknown communicator %lld", (long long)comm);
        c = it->second;
        g_comms.erase(it);
    }
    Nccl* api;
    int32_t rc = nccl_load(&api);
    if (rc != NZ_OK) return rc;
    DeviceGuard g(c->device);
    NZ_NCCL(api, api->CommDestroy(c->nccl));
    return NZ_OK;
}

NZ_API int64_t nz_band_chain_create(const nz_chain_config* cfg, int64_t comm, int32_t device, int32_t mode, void* stream) {
    auto c = std::make_shared<Chain>();
    int32_t rc = chain_setup(*c, cfg, mode);
    if (rc != NZ_OK) return rc;
    Comm* cm = nullptr;
    if (comm != 0) {
        std::lock_guard<std::mutex> lk(g_handles_mu);
        auto it = g_comms.find(comm);
        NZ_REQUIRE(it != g_comms.end(), "nz_band_chain_create: unknown communicator %lld", (long long)comm);
        c->comm = it->second;
        cm = c->comm.get();
        device = cm->device;
    }
    const int world = cm ? cm->world : 1, rank = cm ? cm->rank : 0;
    const int dev = device;
    rc = bandset_create(&c->bs, cfg->resolution, cfg->resolution, world, rank, 1, &dev, chain_cap(*c), world > 1 ? cm : nullptr,
                        (cudaStream_t)stream);
    if (rc != NZ_OK) return rc;
    if ((rc = chain_check_fit(*c)) != NZ_OK) return rc;
    std::lock_guard<std::mutex> lk(g_handles_mu);
    const long long h = g_next_handle++;
    g_chains[h] = c;
    return h;
}

NZ_API int64_t nz_band_chain_create_local(const nz_chain_config* cfg, const int32_t* devices, int32_t n_bands, int32_t mode) {
    NZ_REQUIRE(devices && n_bands >= 1 && n_bands <= 64, "nz_band_chain_create_local: bad band count %d", n_bands);
    auto c = std::make_shared<Chain>();
    int32_t rc = chain_setup(*c, cfg, mode);
    if (rc != NZ_OK) return rc;
    rc = bandset_create(&c->bs, cfg->resolution, cfg->resolution, n_bands, 0, n_bands, devices, chain_cap(*c), nullptr,
                        (cudaStream_t)NZ_STREAM_OWN);
    if (rc != NZ_OK) return rc;
    if ((rc = chain_check_fit(*c)) != NZ_OK) return rc;
    std::lock_guard<std::mutex> lk(g_handles_mu);
    const long long h = g_next_handle++;
    g_chains[h] = c;
    return h;
}

NZ_API int32_t nz_band_chain_run(int64_t chain, int32_t timed) {
    std::shared_ptr<Chain> c;
    int32_t rc = find_chain(chain, &c);
    if (rc != NZ_OK) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    return chain_run(*c, timed != 0);
}

NZ_API int32_t nz_band_chain_sync(int64_t chain) {
    std::shared_ptr<Chain> c;
    int32_t rc = find_chain(chain, &c);
    if (rc != NZ_OK) return rc;
    return bandset_sync(*c->bs);
}

NZ_API int32_t nz_band_chain_stage_ms(int64_t chain, float* ms5) {
    NZ_REQUIRE(ms5, "nz_band_chain_stage_ms: null output");
    std::shared_ptr<Chain> c;
    int32_t rc = find_chain(chain, &c);
    if (rc != NZ_OK) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->timed) {
        set_error("nz_band_chain_stage_ms: the last run was not timed");
        return NZ_E_STATE;
    }
    for (int i = 0; i < 5; i++) ms5[i] = 0.0f;
    for (Band& bd : c->bs->b) {
        DeviceGuard g(bd.device);
        NZ_CUDA(cudaEventSynchronize(bd.ev[5]));
        for (int i = 0; i < 5; i++) {
            float ms = 0.0f;
            NZ_CUDA(cudaEventElapsedTime(&ms, bd.ev[i], bd.ev[i + 1]));
            if (ms > ms5[i]) ms5[i] = ms;
        }
    }
    return NZ_OK;
}

NZ_API int32_t nz_band_chain_local_bands(int64_t chain) {
    std::shared_ptr<Chain> c;
    int32_t rc = find_chain(chain, &c);
    if (rc != NZ_OK) return rc;
    return (int32_t)c->bs->b.size();
}

NZ_API int32_t nz_band_chain_info(int64_t chain, int32_t local_band, nz_band_info* out) {
    NZ_REQUIRE(out, "nz_band_chain_info: null output");
    std::shared_ptr<Chain> c;
    int32_t rc = find_chain(chain, &c);
    if (rc != NZ_OK) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    BandSet& bs = *c->bs;
    NZ_REQUIRE(local_band >= 0 && local_band < (int)bs.b.size(), "nz_band_chain_info: band %d out of range", local_band);
    Band& bd = bs.b[local_band];
    out->rank = bd.rank;
    out->world = bs.world;
    out->device = bd.device;
    out->z0 = bd.z0;
    out->z1 = bd.z1;
    if (c->cfg.mesh_resolution > 0) band_vertex_rows(bd, bs.world, bs.rows, c->cfg.mesh_resolution, &out->vz0, &out->vz1);
    else out->vz0 = out->vz1 = 0;
    out->d_rows = bd.buf[bd.cur] + (size_t)bd.above * bs.width;
    out->d_vertices = bd.vtx;
    out->d_indices = bd.idx;
    out->halo_bytes_per_run = c->runs ? bd.halo_bytes / c->runs : 0;
    return NZ_OK;
}

NZ_API int32_t nz_band_chain_download(int64_t chain, float* h_heights, void* h_vertices, uint32_t* h_indices) {
    std::shared_ptr<Chain> c;
    int32_t rc = find_chain(chain, &c);
    if (rc != NZ_OK) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    BandSet& bs = *c->bs;
    const size_t W = bs.width;
    const int R = c->cfg.mesh_resolution;
    for (Band& bd : bs.b) {
        DeviceGuard g(bd.device);
        if (h_heights)
            NZ_CUDA(cudaMemcpyAsync(h_heights + (size_t)bd.z0 * W, bd.buf[bd.cur] + (size_t)bd.above * W, (size_t)bd.own * W * sizeof(float),
                                    cudaMemcpyDeviceToHost, bd.s));
        if (R > 0 && bd.vz1 > bd.vz0) {
            const int t0 = bd.vz0 > 1 ? bd.vz0 : 1;
            if (h_vertices)
                NZ_CUDA(cudaMemcpyAsync((char*)h_vertices + (size_t)bd.vz0 * (R + 1) * NZ_MESH_VERTEX_BYTES, bd.vtx,
                                        (size_t)(bd.vz1 - bd.vz0) * (R + 1) * NZ_MESH_VERTEX_BYTES, cudaMemcpyDeviceToHost, bd.s));
            if (h_indices && bd.vz1 > t0)
                NZ_CUDA(cudaMemcpyAsync(h_indices + (size_t)(t0 - 1) * 6 * R, bd.idx, (size_t)(bd.vz1 - t0) * 6 * R * sizeof(uint32_t),
                                        cudaMemcpyDeviceToHost, bd.s));
        }
    }
    return bandset_sync(bs);
}

NZ_API int32_t nz_band_chain_destroy(int64_t chain) {
    std::shared_ptr<Chain> c;
    {
        std::lock_guard<std::mutex> lk(g_handles_mu);
        auto it = g_chains.find(chain);
        NZ_REQUIRE(it != g_chains.end(), "nz_band_chain_destroy: unknown band chain handle %lld", (long long)chain);
        c = it->second;
        g_chains.erase(it);
    }
    std::lock_guard<std::mutex> lk(c->mu);
    c->bs.reset();
    return NZ_OK;
}

}  // extern "C"

// =====================================================================================================================
// tiled worlds: several tiles in flight on one GPU (BASELINE config C4)
// =====================================================================================================================
namespace nz {
namespace {

struct TileSlot {
    cudaStream_t s = nullptr;
    float *a = nullptr, *b = nullptr, *edge = nullptr;
    void* vtx = nullptr;
    uint32_t* idx = nullptr;
    float* heights = nullptr;      // which of a / b holds the last tile's filtered heights
    float* edges = nullptr;        // which of edge / (the other of a, b) holds its edge map
};
struct TileWorld {
    nz_tile_config cfg{};
    int device = 0;
    float kx[NZ_MAX_KERNEL_WIDTH], kz[NZ_MAX_KERNEL_WIDTH], factor = 1.0f;
    int ksize = 0;
    std::vector<TileSlot> slots;
    std::mutex mu;
    ~TileWorld() {
        DeviceGuard g(device);
        for (TileSlot& t : slots) {
            if (t.s) cudaStreamSynchronize(t.s);
            dev_free(t.a); dev_free(t.b); dev_free(t.edge); dev_free(t.vtx); dev_free(t.idx);
            if (t.s) cudaStreamDestroy(t.s);
        }
    }
};
std::unordered_map<long long, std::shared_ptr<TileWorld>> g_tile_worlds;

int32_t find_tile_world(long long h, std::shared_ptr<TileWorld>* out) {
    std::lock_guard<std::mutex> lk(g_handles_mu);
    auto it = g_tile_worlds.find(h);
    if (it == g_tile_worlds.end()) {
        set_error("unknown tile world handle %lld", h);
        return NZ_E_INVALID;
    }
    *out = it->second;
    return NZ_OK;
}

}  // namespace
}  // namespace nz

extern "C" {

NZ_API int64_t nz_tile_world_create(const nz_tile_config* cfg, int32_t device, int32_t slots) {
    NZ_REQUIRE(cfg, "nz_tile_world_create: null config");
    NZ_REQUIRE(cfg->resolution > 0 && cfg->resolution <= 16384 && cfg->tile_resolution > 0, "nz_tile_world_create: bad tile resolution");
    NZ_REQUIRE(slots >= 1 && slots <= 64, "nz_tile_world_create: slots %d out of range [1,64]", slots);
    NZ_REQUIRE(cfg->filter_iterations >= 0 && cfg->edge_filter_iterations >= 0, "nz_tile_world_create: negative iteration count");
    int ndev = 0;
    NZ_CUDA(cudaGetDeviceCount(&ndev));
    NZ_REQUIRE(device >= 0 && device < ndev, "nz_tile_world_create: device %d not visible (%d devices)", device, ndev);
    auto w = std::make_shared<TileWorld>();
    w->cfg = *cfg;
    w->device = device;
    if (cfg->filter_iterations > 0) {
        NZ_REQUIRE(cfg->filter_type >= 0 && cfg->filter_type < NZ_FILTER__COUNT, "nz_tile_world_create: filter_type %d out of range", cfg->filter_type);
        if (cfg->filter_type != NZ_FILTER_SOBEL3_2D) {
            int32_t rc = kernel_filter_table(cfg->filter_type, w->kx, w->kz, &w->ksize, &w->factor);
            if (rc != NZ_OK) return rc;
        }
    }
    if (cfg->edge_filter_iterations > 0)
        NZ_REQUIRE(cfg->edge_filter_type >= 0 && cfg->edge_filter_type < NZ_FILTER__COUNT, "nz_tile_world_create: edge_filter_type %d out of range",
                   cfg->edge_filter_type);
    if (cfg->mesh_resolution > 0)
        NZ_REQUIRE(cfg->mesh_type >= 0 && cfg->mesh_type < NZ_MESH__COUNT, "nz_tile_world_create: mesh_type %d out of range", cfg->mesh_type);
    DeviceGuard g(device);
    const size_t n = (size_t)cfg->resolution * cfg->resolution, R = (size_t)cfg->mesh_resolution;
    w->slots.resize(slots);
    for (TileSlot& t : w->slots) {
        NZ_CUDA(cudaStreamCreateWithFlags(&t.s, cudaStreamNonBlocking));
        int32_t rc = dev_alloc((void**)&t.a, n * sizeof(float));
        if (rc == NZ_OK) rc = dev_alloc((void**)&t.b, n * sizeof(float));
        if (rc == NZ_OK && cfg->edge_filter_iterations > 0) rc = dev_alloc((void**)&t.edge, n * sizeof(float));
        if (rc == NZ_OK && R > 0) rc = dev_alloc(&t.vtx, (R + 1) * (R + 1) * NZ_MESH_VERTEX_BYTES);
        if (rc == NZ_OK && R > 0) rc = dev_alloc((void**)&t.idx, 6 * R * R * sizeof(uint32_t));
        if (rc != NZ_OK) return rc;          // ~TileWorld frees what was allocated
    }
    std::lock_guard<std::mutex> lk(g_handles_mu);
    const long long h = g_next_handle++;
    g_tile_worlds[h] = w;
    return h;
}

NZ_API int32_t nz_tile_world_run(int64_t world, const int32_t* tiles_xz, int32_t n, float* h_heights, float* h_edges,
                                 void* h_vertices, uint32_t* h_indices) {
    NZ_REQUIRE(n >= 0 && (tiles_xz || n == 0), "nz_tile_world_run: bad tile list");
    std::shared_ptr<TileWorld> w;
    int32_t rc = find_tile_world(world, &w);
    if (rc != NZ_OK) return rc;
    std::lock_guard<std::mutex> lk(w->mu);
    const nz_tile_config& c = w->cfg;
    NZ_REQUIRE(!h_edges || c.edge_filter_iterations > 0, "nz_tile_world_run: this world has no edge filter");
    NZ_REQUIRE((!h_vertices && !h_indices) || c.mesh_resolution > 0, "nz_tile_world_run: this world has no mesh");
    DeviceGuard g(w->device);
    Range rg("nz.tile_world");
    const int res = c.resolution, R = c.mesh_resolution;
    const size_t cells = (size_t)res * res;
    const size_t vbytes = (size_t)(R + 1) * (R + 1) * NZ_MESH_VERTEX_BYTES, icount = (size_t)6 * R * R;
    for (int k = 0; k < n; k++) {
        TileSlot& t = w->slots[k % w->slots.size()];       // stream order makes the slot's reuse safe
        const int tx = tiles_xz[2 * k], tz = tiles_xz[2 * k + 1];
        FractalParams p;
        // tile -> noise domain: xpos = tileResolution * tx, zpos = tileResolution * tz (MeshTileGenerator.cs:188-189)
        rc = fractal_params(&p, res, res, 0, c.noise_type, c.hurst, c.starting_amplitude, c.stepdown, c.detune_rate, c.octaves,
                            c.tile_resolution * tx, c.tile_resolution * tz, c.noise_size);
        if (rc == NZ_OK) rc = launch_fractal(t.a, c.noise_type, p, t.s);
        float* cur = t.a;
        if (rc == NZ_OK && c.filter_iterations > 0)
            rc = c.filter_type == NZ_FILTER_SOBEL3_2D ? launch_sobel2d(t.a, t.b, res, res, c.filter_iterations, &cur, t.s)
                                                      : launch_separable(t.a, t.b, res, res, w->ksize, w->kx, w->kz, w->factor, c.filter_iterations, &cur, t.s);
        if (rc != NZ_OK) return rc;
        float* other = cur == t.a ? t.b : t.a;
        float* edges = nullptr;
        if (c.edge_filter_iterations == 1 && c.edge_filter_type == NZ_FILTER_SOBEL3_2D) {
            // one Sobel3_2D pass reads its input and writes its partner: straight from the heights into the edge buffer,
            // no copy (a D2D copy between two kernels of a 70 us tile costs a copy-engine hand-off each way)
            rc = launch_sobel2d(cur, t.edge, res, res, 1, &edges, t.s);
            if (rc != NZ_OK) return rc;
        } else if (c.edge_filter_iterations > 0) {
            NZ_CUDA(cudaMemcpyAsync(t.edge, cur, cells * sizeof(float), cudaMemcpyDeviceToDevice, t.s));
            if (c.edge_filter_type == NZ_FILTER_SOBEL3_2D) {
                rc = launch_sobel2d(t.edge, other, res, res, c.edge_filter_iterations, &edges, t.s);
            } else {
                float ekx[9], ekz[9], ef;
                int eks;
                rc = kernel_filter_table(c.edge_filter_type, ekx, ekz, &eks, &ef);
                if (rc == NZ_OK) rc = launch_separable(t.edge, other, res, res, eks, ekx, ekz, ef, c.edge_filter_iterations, &edges, t.s);
            }
            if (rc != NZ_OK) return rc;
        }
        if (R > 0) {
            rc = launch_mesh(c.mesh_type, t.vtx, t.idx, R, res, c.tile_height, c.tile_size, cur, 0, res, 0, R + 1, t.s);
            if (rc != NZ_OK) return rc;
        }
        t.heights = cur;
        t.edges = edges;
        if (h_heights) NZ_CUDA(cudaMemcpyAsync(h_heights + (size_t)k * cells, cur, cells * sizeof(float), cudaMemcpyDeviceToHost, t.s));
        if (h_edges) NZ_CUDA(cudaMemcpyAsync(h_edges + (size_t)k * cells, edges, cells * sizeof(float), cudaMemcpyDeviceToHost, t.s));
        if (h_vertices) NZ_CUDA(cudaMemcpyAsync((char*)h_vertices + (size_t)k * vbytes, t.vtx, vbytes, cudaMemcpyDeviceToHost, t.s));
        if (h_indices) NZ_CUDA(cudaMemcpyAsync(h_indices + (size_t)k * icount, t.idx, icount * sizeof(uint32_t), cudaMemcpyDeviceToHost, t.s));
    }
    for (TileSlot& t : w->slots) NZ_CUDA(cudaStreamSynchronize(t.s));
    return NZ_OK;
}

NZ_API int32_t nz_tile_world_slot(int64_t world, int32_t slot, float** d_heights, float** d_edges, void** d_vertices, uint32_t** d_indices) {
    std::shared_ptr<TileWorld> w;
    int32_t rc = find_tile_world(world, &w);
    if (rc != NZ_OK) return rc;
    std::lock_guard<std::mutex> lk(w->mu);
    NZ_REQUIRE(slot >= 0 && slot < (int)w->slots.size(), "nz_tile_world_slot: slot %d out of range", slot);
    const TileSlot& t = w->slots[slot];
    if (d_heights) *d_heights = t.heights;
    if (d_edges) *d_edges = t.edges;
    if (d_vertices) *d_vertices = t.vtx;
    if (d_indices) *d_indices = t.idx;
    return NZ_OK;
}

NZ_API int32_t nz_tile_world_destroy(int64_t world) {
    std::shared_ptr<TileWorld> w;
    {
        std::lock_guard<std::mutex> lk(g_handles_mu);
        auto it = g_tile_worlds.find(world);
        NZ_REQUIRE(it != g_tile_worlds.end(), "nz_tile_world_destroy: unknown tile world handle %lld", (long long)world);
        w = it->second;
        g_tile_worlds.erase(it);
    }
    return NZ_OK;      // the last reference frees the slots
}

}  // extern "C"
