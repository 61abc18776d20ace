// abi.cu — the C ABI of libnoize_b200.so (include/noize_b200.h): status/error plumbing, host-side
// stage logic (tables, normalisation value, tile geometry), the device layer (nz_dev_*) and the host
// layer (nz_*) with its per-thread stream, device-buffer pool and residency map.
//
// There is deliberately no CPU path in this file: every compute entry point ends in a kernel
// launch from noise/filter/flow/mesh_kernels.cu or fails with NZ_E_CUDA.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>
#include <nvtx3/nvToolsExt.h>
#include "nz_common.cuh"
#include "bands.cuh"

namespace nz {

// ---- errors -----------------------------------------------------------------------------------
static thread_local char t_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}
int32_t cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return e == cudaErrorMemoryAllocation ? NZ_E_NOMEM : NZ_E_CUDA;
}

// ---- host-side stage logic --------------------------------------------------------------------
// Normalised sampled Gaussian; equals every literal table of Filter/Kernel/Blur/BlurKernels.cs:59-316
// and gauss*_s* of Filter/Kernel/KernelJob.cs:97-105 after rounding to float (tests/test_host_logic.py).
void gauss_table(double sigma, int width, float* out) {
    const int r = width / 2;
    double g[NZ_MAX_KERNEL_WIDTH + 1], sum = 0.0;
    for (int i = -r; i <= r; i++) {
        g[i + r] = exp(-(double)(i * i) / (2.0 * sigma * sigma));
        sum += g[i + r];
    }
    for (int i = 0; i < width; i++) out[i] = (float)(g[i] / sum);
}

// SeparableKernelFilter.Schedule switch, KernelJob.cs:217-292
int32_t kernel_filter_table(int filter, float* kx, float* kz, int* ksize, float* factor) {
    static const float k_m101[3] = {-1.f, 0.f, 1.f}, k_121[3] = {1.f, 2.f, 1.f}, k_10m1[3] = {1.f, 0.f, -1.f},
                       k_111[3] = {1.f, 1.f, 1.f};
    float g[9];
    const float *x = k_111, *z = k_111;
    int size = 3;
    float f = 1.0f;
    switch (filter) {
        case NZ_FILTER_GAUSS9_S1: size = 9; gauss_table(1.0, 9, g); x = z = g; break;
        case NZ_FILTER_GAUSS7_S1: size = 7; gauss_table(1.0, 7, g); x = z = g; break;
        case NZ_FILTER_GAUSS5_S1: size = 5; gauss_table(1.0, 5, g); x = z = g; break;
        case NZ_FILTER_GAUSS3_S1: size = 3; gauss_table(1.0, 3, g); x = z = g; break;
        case NZ_FILTER_GAUSS9_S2: size = 9; gauss_table(2.0, 9, g); x = z = g; break;
        case NZ_FILTER_GAUSS7_S2: size = 7; gauss_table(2.0, 7, g); x = z = g; break;
        case NZ_FILTER_GAUSS5_S2: size = 5; gauss_table(2.0, 5, g); x = z = g; break;
        case NZ_FILTER_GAUSS3_S2: size = 3; gauss_table(2.0, 3, g); x = z = g; break;
        case NZ_FILTER_SMOOTH3: f = 1.0f / 3.0f; break;
        case NZ_FILTER_SOBEL3_HORIZONTAL: x = k_m101; z = k_121; break;
        case NZ_FILTER_SOBEL3_VERTICAL: x = k_121; z = k_10m1; break;
        case NZ_FILTER_PREWITT3_HORIZONTAL: x = k_10m1; z = k_111; break;
        case NZ_FILTER_PREWITT3_VERTICAL: x = k_111; z = k_m101; break;
        case NZ_FILTER_SOBEL3_2D:
            set_error("kernel_filter_table: Sobel3_2D is a two-branch reduce, not one separable kernel");
            return NZ_E_UNSUPPORTED;
        default:
            set_error("kernel_filter_table: filter_type %d out of range", filter);
            return NZ_E_INVALID;
    }
    memcpy(kx, x, size * sizeof(float));
    memcpy(kz, z, size * sizeof(float));
    *ksize = size;
    *factor = f;
    return NZ_OK;
}

// ---- process context ----------------------------------------------------------------------------
struct Context {
    std::mutex mu;
    bool ready = false;
    int n_visible = 0;
    std::vector<int> devices{0};   // nz_init's list; devices[0] is the device of the un-banded host layer
    int n_bands = 0;               // nz_set_bands: > 1 splits large host-layer grids over devices[0..n_bands)
    // device-buffer pool: freed blocks are kept, keyed by (device, size) (stage buffers repeat the same sizes)
    std::multimap<std::pair<int, size_t>, void*> free_blocks;
    std::unordered_map<void*, std::pair<int, size_t>> live;
    long long fail_skip = 0, fail_allocs = 0;   // fault injection (nz_test_fail_allocs): after fail_skip allocations the next fail_allocs fail
};
static Context g_ctx;

// The CUDA runtime is usable and nz_init's devices exist.  Does NOT touch the calling thread's current device: the
// device layer (nz_dev_*) runs on caller-owned buffers and streams, on whatever device the caller has made current.
static int32_t ensure_runtime() {
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    if (!g_ctx.ready) {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n <= 0) {
            set_error("no CUDA device available (%s); libnoize_b200 has no CPU fallback",
                      e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
            return NZ_E_CUDA;
        }
        for (int d : g_ctx.devices)
            if (d < 0 || d >= n) {
                set_error("device %d requested but only %d visible", d, n);
                return NZ_E_CUDA;
            }
        g_ctx.n_visible = n;
        g_ctx.ready = true;
    }
    return NZ_OK;
}

static int primary_device() {
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    return g_ctx.devices[0];
}

// Host layer: the runtime is usable and the calling thread's current device is the primary device of nz_init.
static int32_t ensure_init() {
    int32_t rc = ensure_runtime();
    if (rc != NZ_OK) return rc;
    NZ_CUDA(cudaSetDevice(primary_device()));
    return NZ_OK;
}

// The device layer trusts the caller's current device; a pointer that lives on ANOTHER device is the one mistake that
// would otherwise surface as an illegal address inside a kernel, so it is checked here (cheap: one attribute query).
static int32_t check_device_pointer(const void* p, const char* who) {
    if (!p) return NZ_OK;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return NZ_OK;   // not a pointer the runtime knows (e.g. a VMM mapping): let the launch decide
    }
    if (a.type != cudaMemoryTypeDevice) return NZ_OK;
    int cur = 0;
    NZ_CUDA(cudaGetDevice(&cur));
    if (a.device != cur) {
        set_error("%s: the buffer lives on device %d but the calling thread's current device is %d (cudaSetDevice first)", who,
                  a.device, cur);
        return NZ_E_INVALID;
    }
    return NZ_OK;
}

// Allocates on the CURRENT device.
int32_t dev_alloc(void** p, size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    int dev = 0;
    NZ_CUDA(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lk(g_ctx.mu);
        if (g_ctx.fail_allocs > 0 && g_ctx.fail_skip-- <= 0) {
            g_ctx.fail_skip = 0;
            g_ctx.fail_allocs--;
            set_error("device allocation of %zu bytes failed (injected fault)", bytes);
            return NZ_E_NOMEM;
        }
        auto it = g_ctx.free_blocks.find({dev, bytes});
        if (it != g_ctx.free_blocks.end()) {
            *p = it->second;
            g_ctx.free_blocks.erase(it);
            g_ctx.live[*p] = {dev, bytes};
            return NZ_OK;
        }
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation) {
        // give this device's cached blocks back to the driver and retry once
        cudaGetLastError();
        std::vector<void*> drop;
        {
            std::lock_guard<std::mutex> lk(g_ctx.mu);
            for (auto it = g_ctx.free_blocks.begin(); it != g_ctx.free_blocks.end();) {
                if (it->first.first == dev) {
                    drop.push_back(it->second);
                    it = g_ctx.free_blocks.erase(it);
                } else {
                    ++it;
                }
            }
        }
        for (void* d : drop) cudaFree(d);
        e = cudaMalloc(p, bytes);
    }
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    g_ctx.live[*p] = {dev, bytes};
    return NZ_OK;
}

// Returns a block to the pool.  The caller has ordered every use of the block before this call (stream synchronised).
void dev_free(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    auto it = g_ctx.live.find(p);
    if (it == g_ctx.live.end()) return;
    g_ctx.free_blocks.emplace(it->second, p);
    g_ctx.live.erase(it);
}
static inline int32_t pool_alloc(void** p, size_t bytes) { return dev_alloc(p, bytes); }
static inline void pool_free(void* p) { dev_free(p); }

// ---- per-thread state: stream, timing events, residency map ------------------------------------
struct Mirror {
    float* d = nullptr;    // device copy of the slice (contiguous n floats)
    float* d_tmp = nullptr;  // ping-pong partner, allocated on demand, same size
    size_t n = 0;
    nz_slice_f32 host{};
    bool dirty = false;  // device newer than host
    cudaEvent_t ready = nullptr;        // recorded after the last stage that touched the mirror (cross-thread ordering)
    cudaStream_t last_stream = nullptr;
    // nz_set_bands: the mirror of a large square grid is a set of row bands, one per device, instead of d / d_tmp.  All
    // work on a banded mirror runs on the bands' own streams, so stages issued from different threads stay ordered.
    std::shared_ptr<BandSet> bands;
    // element-wise stages issued inside a scope and not applied yet (see materialize()): d holds the values BEFORE them
    std::vector<PointwiseOp> pending;
};

// A residency scope: the device mirrors of the host slices one chain of stages works on.  Scopes are process-wide
// objects so that the stages of a chain may run on different threads (Unity runs every IJob on an arbitrary worker):
// a thread ENTERS a scope, makes its stage call, and leaves; the mirror's `ready` event orders the streams.
struct Scope {
    std::mutex mu;
    std::unordered_map<const void*, Mirror> mirrors;
    // A scope that is closed while another thread is still inside it can receive mirrors afterwards (that thread's stage
    // calls keep working on the object it entered); they are released when the last thread leaves.  Their contents are not
    // brought home: the scope was closed, the host was told so.
    ~Scope();
};
static std::mutex g_scopes_mu;
static std::unordered_map<long long, std::shared_ptr<Scope>> g_scopes;
static std::atomic<long long> g_next_scope{1};

struct ThreadState {
    cudaStream_t stream = nullptr;
    int device = 0;                 // device `stream` and `ev` belong to
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // start, after h2d, after kernels, after d2h
    cudaStream_t copy_stream = nullptr;                         // D2H of mesh chunks while the next chunk is built
    cudaEvent_t mesh_ev[4] = {nullptr, nullptr, nullptr, nullptr};   // chunk built [slot], chunk downloaded [2 + slot]
    bool timed = false;
    long long launches_at_start = 0;
    int launches = 0;
    std::shared_ptr<Scope> scope;   // scope this thread is inside of (nz_scope_enter / nz_pipeline_begin), if any
    long long scope_id = 0;
    bool owns_scope = false;        // entered through nz_pipeline_begin: nz_pipeline_end closes it
    Scope local;                    // mirrors of a call made outside any scope (dropped when the call returns)
    // A host with a job system creates and retires worker threads: the thread's stream and events go with it.  (At process
    // exit the runtime may be unloading already; the calls then fail harmlessly.)
    ~ThreadState() {
        if (!stream) return;
        int prev = 0;
        const bool have = cudaGetDevice(&prev) == cudaSuccess;
        if (cudaSetDevice(device) == cudaSuccess) {
            cudaStreamSynchronize(stream);
            cudaStreamDestroy(stream);
            for (auto& e : ev) if (e) cudaEventDestroy(e);
            if (copy_stream) cudaStreamDestroy(copy_stream);
            for (auto& e : mesh_ev) if (e) cudaEventDestroy(e);
            if (have) cudaSetDevice(prev);
        }
        cudaGetLastError();
        stream = nullptr;
    }
};
static thread_local ThreadState t_state;
static inline bool in_scope() { return (bool)t_state.scope; }
static inline Scope& cur_scope() { return t_state.scope ? *t_state.scope : t_state.local; }

static int32_t thread_ready() {
    int32_t rc = ensure_init();
    if (rc != NZ_OK) return rc;
    const int dev = primary_device();
    if (t_state.stream && t_state.device != dev) {
        // nz_init re-pointed the library at another device: this thread's stream and events belong to the old one
        cudaSetDevice(t_state.device);
        cudaStreamSynchronize(t_state.stream);
        cudaStreamDestroy(t_state.stream);
        for (auto& e : t_state.ev) {
            if (e) cudaEventDestroy(e);
            e = nullptr;
        }
        if (t_state.copy_stream) {
            cudaStreamDestroy(t_state.copy_stream);
            t_state.copy_stream = nullptr;
            for (auto& e : t_state.mesh_ev) {
                if (e) cudaEventDestroy(e);
                e = nullptr;
            }
        }
        t_state.stream = nullptr;
        t_state.timed = false;
        NZ_CUDA(cudaSetDevice(dev));
    }
    if (!t_state.stream) {
        NZ_CUDA(cudaStreamCreateWithFlags(&t_state.stream, cudaStreamNonBlocking));
        for (auto& e : t_state.ev) NZ_CUDA(cudaEventCreate(&e));
        t_state.device = dev;
    }
    return NZ_OK;
}

static int32_t check_slice(const nz_slice_f32& s, long long expect, const char* who) {
    NZ_REQUIRE(s.ptr != nullptr, "%s: slice pointer is null", who);
    NZ_REQUIRE(s.stride_bytes >= 4, "%s: slice stride %d < 4", who, s.stride_bytes);
    NZ_REQUIRE((long long)s.length == expect, "%s: slice length %d != %lld (resolution^2)", who, s.length, expect);
    return NZ_OK;
}

static int32_t upload(Mirror& m) {
    cudaStream_t s = t_state.stream;
    if (m.host.stride_bytes == 4) {
        NZ_CUDA(cudaMemcpyAsync(m.d, m.host.ptr, m.n * sizeof(float), cudaMemcpyHostToDevice, s));
        return NZ_OK;
    }
    // strided NativeSlice (e.g. one channel of an RGBAFloat texture): move the span, gather on device
    const size_t span = (m.n - 1) * (size_t)m.host.stride_bytes + sizeof(float);
    void* raw = nullptr;
    int32_t rc = pool_alloc(&raw, span);
    if (rc != NZ_OK) return rc;
    cudaError_t e = cudaMemcpyAsync(raw, m.host.ptr, span, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
        rc = launch_gather_strided(m.d, (const unsigned char*)raw, m.host.stride_bytes, m.n, s);
        if (rc == NZ_OK) e = cudaStreamSynchronize(s);
    }
    pool_free(raw);
    if (e != cudaSuccess) return cuda_fail(e, "strided upload");
    return rc;
}

// Element-wise stages (nz_constant, nz_normalize) called inside a residency scope are DEFERRED: the call appends a step to
// the mirror's `pending` list and returns.  The list is applied as one pass (launch_pointwise_chain: 8 B of HBM traffic per
// cell for the whole run instead of 8 B per stage) together with the next nz_curve, or when anything else needs the values:
// another stage (acquire), the download at scope close / flush, a named-buffer write.  NZ_POINTWISE_FUSE=0 applies every
// stage at once.  The values are those of the separate passes bit for bit.  Banded mirrors do not defer.
static bool pointwise_defer_enabled() {
    static const bool on = [] {
        const char* e = getenv("NZ_POINTWISE_FUSE");
        return !(e && e[0] == '0');
    }();
    return on;
}
static int32_t materialize(Mirror& m, const PointwiseOp* last = nullptr) {
    if (last) m.pending.push_back(*last);
    if (m.pending.empty()) return NZ_OK;
    int32_t rc = launch_pointwise_chain(m.d, m.n, m.pending.data(), (int)m.pending.size(), t_state.stream);
    m.pending.clear();
    m.dirty = true;
    return rc;
}

static int32_t download(Mirror& m) {
    cudaStream_t s = t_state.stream;
    int32_t prc = materialize(m);
    if (prc != NZ_OK) return prc;
    if (m.host.stride_bytes == 4) {
        NZ_CUDA(cudaMemcpyAsync(m.host.ptr, m.d, m.n * sizeof(float), cudaMemcpyDeviceToHost, s));
    } else {
        // scatter back without touching the other channels of the host span
        NZ_CUDA(cudaMemcpy2DAsync(m.host.ptr, (size_t)m.host.stride_bytes, m.d, sizeof(float), sizeof(float), m.n,
                                  cudaMemcpyDeviceToHost, s));
    }
    m.dirty = false;
    return NZ_OK;
}

// ---- banded mirrors (nz_set_bands) ------------------------------------------------------------------------------
constexpr int BAND_GHOST_CAP = 288;     // ghost rows kept per band side: Gauss9 x 32 iterations = 128, flow x 128 = 257
constexpr int BAND_MIN_ROWS = 64;       // a band narrower than this is not worth a device

// the calling thread's stream waits for every band's last stage (before it reads the bands, or to time them)
static int32_t banded_join(Mirror& m) {
    for (Band& bd : m.bands->b) NZ_CUDA(cudaStreamWaitEvent(t_state.stream, bd.done, 0));
    return NZ_OK;
}

static int32_t banded_copy_rows(Mirror& m, bool to_host) {
    BandSet& bs = *m.bands;
    const size_t W = bs.width;
    int prev = 0;
    NZ_CUDA(cudaGetDevice(&prev));
    cudaError_t e = cudaSuccess;
    for (Band& bd : bs.b) {
        e = cudaSetDevice(bd.device);
        if (e != cudaSuccess) break;
        float* dev = bd.buf[bd.cur] + (size_t)bd.above * W;
        float* host = m.host.ptr + (size_t)bd.z0 * W;
        const size_t bytes = (size_t)bd.own * W * sizeof(float);
        e = to_host ? cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, bd.s)
                    : cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, bd.s);
        if (e == cudaSuccess) e = cudaEventRecord(bd.done, bd.s);
        if (e != cudaSuccess) break;
    }
    cudaSetDevice(prev);
    if (e != cudaSuccess) return cuda_fail(e, to_host ? "banded download" : "banded upload");
    if (to_host) m.dirty = false;
    return NZ_OK;
}

// Gather a banded mirror into an ordinary one on the primary device (for a stage that has no banded form).
static int32_t unband(Mirror& m) {
    BandSet& bs = *m.bands;
    const size_t W = bs.width;
    int32_t rc = pool_alloc((void**)&m.d, m.n * sizeof(float));
    if (rc != NZ_OK) return rc;
    rc = banded_join(m);
    cudaError_t e = cudaSuccess;
    if (rc == NZ_OK)
        for (Band& bd : bs.b) {
            e = cudaMemcpyPeerAsync(m.d + (size_t)bd.z0 * W, t_state.device, bd.buf[bd.cur] + (size_t)bd.above * W, bd.device,
                                    (size_t)bd.own * W * sizeof(float), t_state.stream);
            if (e != cudaSuccess) break;
        }
    if (rc == NZ_OK && e == cudaSuccess) e = cudaStreamSynchronize(t_state.stream);   // the bands are freed next
    if (rc != NZ_OK || e != cudaSuccess) {
        pool_free(m.d);
        m.d = nullptr;
        return rc != NZ_OK ? rc : cuda_fail(e, "gathering a banded mirror");
    }
    m.bands.reset();
    return NZ_OK;
}

static void release_mirror(Mirror& m) {
    pool_free(m.d);
    pool_free(m.d_tmp);
    m.d = m.d_tmp = nullptr;
    m.bands.reset();        // ~BandSet synchronises its streams and frees its buffers
    if (m.ready) cudaEventDestroy(m.ready);
    m.ready = nullptr;
}

Scope::~Scope() {
    if (mirrors.empty()) return;
    // whatever still runs on these buffers (any stream of the primary device).  At process exit the runtime may already be
    // unloading: then nothing is released (the process is going away with its memory)
    if (cudaDeviceSynchronize() != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    for (auto& kv : mirrors) release_mirror(kv.second);
    mirrors.clear();
}

// Obtain the device mirror of a host slice.  `need_contents`: the stage reads the slice (H2D unless a resident mirror is
// already newer than the host).  `band_res` > 0: the slice is a band_res^2 grid and the calling stage has a banded form
// (nz_set_bands); 0: the stage needs an ordinary single-device mirror (a banded one is gathered first).
static int32_t acquire(const nz_slice_f32& s, bool need_contents, Mirror** out, int band_res = 0, bool keep_pending = false) {
    Scope& sc = cur_scope();
    std::lock_guard<std::mutex> lk(sc.mu);
    auto& mirrors = sc.mirrors;
    auto it = mirrors.find(s.ptr);
    if (it != mirrors.end() && it->second.ready && it->second.last_stream != t_state.stream)
        NZ_CUDA(cudaStreamWaitEvent(t_state.stream, it->second.ready, 0));   // previous stage ran on another thread's stream
    if (it != mirrors.end() && (it->second.n != (size_t)s.length || it->second.host.stride_bytes != s.stride_bytes)) {
        // same base pointer, different shape: drop the stale mirror (after everything in flight on it has finished: the
        // blocks go back to a pool other threads allocate from)
        Mirror& old = it->second;
        if (old.dirty) {
            int32_t rc = old.bands ? banded_copy_rows(old, /*to_host=*/true) : download(old);
            if (rc != NZ_OK) return rc;
        }
        if (old.ready && old.last_stream != t_state.stream) cudaEventSynchronize(old.ready);
        NZ_CUDA(cudaStreamSynchronize(t_state.stream));
        release_mirror(old);
        mirrors.erase(it);
        it = mirrors.end();
    }
    if (it == mirrors.end()) {
        Mirror m;
        m.n = (size_t)s.length;
        m.host = s;
        int n_bands = 0;
        std::vector<int> devs;
        {
            std::lock_guard<std::mutex> lk2(g_ctx.mu);
            n_bands = g_ctx.n_bands;
            devs = g_ctx.devices;
        }
        const bool banded = band_res >= NZ_BANDS_MIN_RESOLUTION && n_bands > 1 && s.stride_bytes == 4 &&
                            band_res / n_bands >= BAND_MIN_ROWS;
        int32_t rc;
        if (banded) {
            std::unique_ptr<BandSet> bs;
            rc = bandset_create(&bs, band_res, band_res, n_bands, 0, n_bands, devs.data(), BAND_GHOST_CAP, nullptr,
                                (cudaStream_t)NZ_STREAM_OWN);
            if (rc != NZ_OK) return rc;
            m.bands = std::move(bs);
            if (need_contents && (rc = banded_copy_rows(m, /*to_host=*/false)) != NZ_OK) return rc;   // m frees its bands
        } else {
            rc = pool_alloc((void**)&m.d, m.n * sizeof(float));
            if (rc != NZ_OK) return rc;
            if (need_contents && (rc = upload(m)) != NZ_OK) {
                cudaStreamSynchronize(t_state.stream);
                pool_free(m.d);
                return rc;
            }
        }
        // the mirror enters the map only once it is complete: a failed call must not leave a half-made mirror that a
        // later call with the same host pointer would trust
        it = mirrors.emplace(s.ptr, m).first;
    } else if (it->second.bands && band_res == 0) {
        int32_t rc = unband(it->second);
        if (rc != NZ_OK) return rc;
    }
    if (!keep_pending && !it->second.pending.empty()) {
        // the caller reads (or overwrites) the values: apply the deferred element-wise run first (the stream already waits
        // for the mirror's previous stage)
        if (!need_contents) it->second.pending.clear();      // the caller overwrites every value
        int32_t rc = materialize(it->second);
        if (rc != NZ_OK) return rc;
    }
    *out = &it->second;
    return NZ_OK;
}

// Outside a scope a mirror lives for one call.  Every host-layer entry point holds one of these: whatever a failed (or
// finished) call left in the thread-local map is released when the call returns, so a retry starts from the host data.
struct UnscopedCleanup {
    ~UnscopedCleanup() {
        if (t_state.scope || t_state.local.mirrors.empty()) return;
        if (t_state.stream) cudaStreamSynchronize(t_state.stream);
        for (auto& kv : t_state.local.mirrors) release_mirror(kv.second);
        t_state.local.mirrors.clear();
    }
};

static int32_t ensure_tmp(Mirror& m) {
    if (!m.d_tmp) return pool_alloc((void**)&m.d_tmp, m.n * sizeof(float));
    return NZ_OK;
}

// A stage left its result in `result` (either m.d or m.d_tmp): make it the mirror's primary buffer.
static void adopt_result(Mirror& m, float* result) {
    if (result == m.d_tmp) {
        m.d_tmp = m.d;
        m.d = result;
    }
    m.dirty = true;
}

// End of a host-layer stage: outside a pipeline, bring the result home and drop the mirror.
static int32_t finish(Mirror* m) {
    cudaStream_t s = t_state.stream;
    if (m->bands) {
        // the stage ran on the bands' streams: the thread's stream joins them (timing; and the download below)
        int32_t rc = banded_join(*m);
        if (rc != NZ_OK) return rc;
        NZ_CUDA(cudaEventRecord(t_state.ev[2], s));
        if (!in_scope()) {
            if (m->dirty) rc = banded_copy_rows(*m, /*to_host=*/true);   // every band brings its own rows home
            if (rc == NZ_OK) rc = bandset_sync(*m->bands);
            release_mirror(*m);
            t_state.local.mirrors.erase(m->host.ptr);
            if (rc != NZ_OK) return rc;
        }
        NZ_CUDA(cudaEventRecord(t_state.ev[3], s));
        t_state.timed = true;
        t_state.launches = (int)(g_launches.load() - t_state.launches_at_start);
        return NZ_OK;
    }
    NZ_CUDA(cudaEventRecord(t_state.ev[2], s));
    if (!in_scope()) {
        int32_t rc = NZ_OK;
        if (m->dirty) rc = download(*m);
        cudaError_t e = cudaEventRecord(t_state.ev[3], s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        release_mirror(*m);
        t_state.local.mirrors.erase(m->host.ptr);
        if (rc != NZ_OK) return rc;
        if (e != cudaSuccess) return cuda_fail(e, "stage completion");
    } else {
        NZ_CUDA(cudaEventRecord(t_state.ev[3], s));
        // publish: whoever touches this mirror next (possibly another thread, another stream) waits for this stage
        if (!m->ready) NZ_CUDA(cudaEventCreateWithFlags(&m->ready, cudaEventDisableTiming));
        NZ_CUDA(cudaEventRecord(m->ready, s));
        m->last_stream = s;
    }
    t_state.timed = true;
    t_state.launches = (int)(g_launches.load() - t_state.launches_at_start);
    return NZ_OK;
}

// A stage only READ this mirror (the right operand of a reduce, a curve's samples, the input of a crop, the heights of a
// mesh): inside a scope whoever touches it next — possibly on another thread's stream, possibly to overwrite it — must wait
// for that read as well.  The reader's stream already waited for the last writer (acquire), so re-recording `ready` here
// keeps the order transitive.
static void publish_read(Mirror* m) {
    if (!m || !in_scope() || m->bands) return;
    if (!m->ready && cudaEventCreateWithFlags(&m->ready, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        m->ready = nullptr;
        cudaStreamSynchronize(t_state.stream);      // no event to hand on: finish the read now
        return;
    }
    cudaEventRecord(m->ready, t_state.stream);
    m->last_stream = t_state.stream;
}

static int32_t begin_stage() {
    int32_t rc = thread_ready();
    if (rc != NZ_OK) return rc;
    t_state.timed = false;
    t_state.launches_at_start = g_launches.load();
    NZ_CUDA(cudaEventRecord(t_state.ev[0], t_state.stream));
    return NZ_OK;
}
static int32_t mark_uploaded() {
    NZ_CUDA(cudaEventRecord(t_state.ev[1], t_state.stream));
    return NZ_OK;
}

int32_t fractal_params(FractalParams* p, int width, int rows, int z_first, int noise_type, float hurst,
                              float start_amp, float stepdown, float detune, int octaves, int xpos, int zpos,
                              int noise_size) {
    NZ_REQUIRE(noise_type >= 0 && noise_type < NZ_NOISE__COUNT, "fractal: noise_type %d out of range", noise_type);
    NZ_REQUIRE(width > 0 && rows > 0, "fractal: bad grid %d x %d", width, rows);
    NZ_REQUIRE(octaves >= 1 && octaves <= 64, "fractal: octaves %d out of range [1,64]", octaves);
    NZ_REQUIRE(noise_size != 0, "fractal: noise_size must be non-zero");
    p->width = width;
    p->rows = rows;
    p->z_first = z_first;
    p->octaves = octaves;
    p->posx = (float)xpos;  // FractalGenerator.SetPosition, Fractal.cs:109-112
    p->posz = (float)zpos;
    p->noise_size = (float)noise_size;
    p->start_amp = start_amp;
    p->stepdown = stepdown;
    p->detune_rate = detune;
    p->G = exp2f(-hurst);                                // Fractal.cs:118
    p->norm = nz_fractal_norm_value(hurst, octaves);     // Fractal.cs:31-40
    // bound on the simplex lattice index |floor(v + (vx+vy)*0.366)| <= 1.74*max|v| over the tile and all octaves
    {
        double f = 1.0, fmax = 1.0, det = 0.0;
        bool positive = noise_size > 0 && xpos >= 0 && (long long)zpos + z_first >= 0;
        {   // the kernels' own float recurrence decides the sign of every octave frequency
            float ff = 1.0f, dd = 0.0f;
            for (int i = 0; i < octaves; i++) {
                dd += detune;
                ff *= (stepdown - dd);
                if (!(ff > 0.0f) && i + 1 < octaves) positive = false;
            }
        }
        p->nonneg = positive;
        for (int i = 0; i < octaves; i++) {
            det += detune;
            f *= ((double)stepdown - det);
            if (fabs(f) > fmax) fmax = fabs(f);
        }
        const double cx = fabs((double)xpos) + width, cz = fabs((double)zpos) + fabs((double)z_first) + rows;
        const double vmax = fmax * (cx > cz ? cx : cz) / fabs((double)noise_size);
        p->fast_hash = (1.74 * vmax + 2.0) < 2097152.0;   // also false for NaN/inf parameters
        // domain-rotated bases: |xr|, |zr| <= 1.42 vmax, |yr| <= 1.16 vmax, simplex skew adds (xr+yr+zr)/3: below 3 vmax
        p->fast_hash3d = (3.0 * vmax + 2.0) < 2097152.0;
    }
    return NZ_OK;
}

}  // namespace nz

using namespace nz;

// =================================================================================================
// lifecycle / introspection
// =================================================================================================
extern "C" {

NZ_API int32_t nz_init(const int32_t* devices, int32_t n) {
    NZ_REQUIRE(n >= 0 && n <= 64, "nz_init: device count %d out of range", n);
    {
        std::lock_guard<std::mutex> lk(g_scopes_mu);
        if (!g_scopes.empty()) {
            set_error("nz_init: %zu residency scope(s) are open; close them before re-initialising", g_scopes.size());
            return NZ_E_STATE;
        }
    }
    std::vector<int> old;
    {
        std::lock_guard<std::mutex> lk(g_ctx.mu);
        old = g_ctx.devices;
        g_ctx.devices.assign(1, 0);
        if (devices && n > 0) g_ctx.devices.assign(devices, devices + n);
        if (g_ctx.n_bands > (int)g_ctx.devices.size()) g_ctx.n_bands = 0;
        g_ctx.ready = false;
    }
    int32_t rc = ensure_init();
    if (rc != NZ_OK) {
        std::lock_guard<std::mutex> lk(g_ctx.mu);   // keep the previous, working configuration
        g_ctx.devices = old;
        g_ctx.ready = false;
    }
    return rc;
}

NZ_API int32_t nz_set_bands(int32_t n_bands) {
    int32_t rc = ensure_runtime();
    if (rc != NZ_OK) return rc;
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    NZ_REQUIRE(n_bands >= 0 && n_bands <= (int)g_ctx.devices.size(),
               "nz_set_bands: %d bands but nz_init was given %zu device(s)", n_bands, g_ctx.devices.size());
    g_ctx.n_bands = n_bands;
    return NZ_OK;
}

NZ_API int32_t nz_shutdown(void) {
    std::vector<std::pair<int, void*>> drop;
    {
        std::lock_guard<std::mutex> lk(g_ctx.mu);
        for (auto& kv : g_ctx.free_blocks) drop.push_back({kv.first.first, kv.second});
        g_ctx.free_blocks.clear();
    }
    int prev = 0;
    cudaGetDevice(&prev);
    for (auto& d : drop) {
        cudaSetDevice(d.first);
        cudaFree(d.second);
    }
    cudaSetDevice(prev);
    return NZ_OK;
}

/* Fault injection for the tests: after `skip` more allocations the next `count` device allocations fail with NZ_E_NOMEM. */
NZ_API int32_t nz_test_fail_allocs(int32_t skip, int32_t count) {
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    g_ctx.fail_skip = skip > 0 ? skip : 0;
    g_ctx.fail_allocs = count > 0 ? count : 0;
    return NZ_OK;
}

NZ_API const char* nz_last_error(void) { return t_err; }
NZ_API const char* nz_version(void) { return "noize_b200 0.1.0 (sm_100a)"; }
NZ_API int64_t nz_kernel_launch_count(void) { return (int64_t)g_launches.load(); }

NZ_API int32_t nz_last_timing(nz_timing* out) {
    NZ_REQUIRE(out, "nz_last_timing: null output");
    if (!t_state.timed) {
        set_error("nz_last_timing: no completed host-layer call on this thread");
        return NZ_E_STATE;
    }
    NZ_CUDA(cudaEventSynchronize(t_state.ev[3]));
    NZ_CUDA(cudaEventElapsedTime(&out->ms_h2d, t_state.ev[0], t_state.ev[1]));
    NZ_CUDA(cudaEventElapsedTime(&out->ms_kernel, t_state.ev[1], t_state.ev[2]));
    NZ_CUDA(cudaEventElapsedTime(&out->ms_d2h, t_state.ev[2], t_state.ev[3]));
    out->kernel_launches = t_state.launches;
    return NZ_OK;
}

// =================================================================================================
// host-side helpers (no GPU)
// =================================================================================================
NZ_API float nz_fractal_norm_value(float hurst, int32_t octaves) {
    float G = exp2f(-hurst);
    float a = 1.0f, t = 0.0f;
    for (int i = 0; i < octaves; i++) {
        t += a * 1.0f;
        a *= G;
    }
    return t;
}

NZ_API int32_t nz_limit_width(int32_t width) {
    if (width % 2 == 0) width += 1;
    if (width > NZ_MAX_KERNEL_WIDTH) width = NZ_MAX_KERNEL_WIDTH;
    return width < 3 ? 3 : width;
}

NZ_API int32_t nz_gauss_kernel(int32_t sigma, int32_t width, float* out, int32_t* width_out) {
    NZ_REQUIRE(sigma >= 0 && sigma < NZ_SIGMA__COUNT, "nz_gauss_kernel: sigma index %d out of range", sigma);
    NZ_REQUIRE(out, "nz_gauss_kernel: null output");
    width = nz_limit_width(width);
    gauss_table(0.5 * (sigma + 1), width, out);
    if (width_out) *width_out = width;
    return NZ_OK;
}

NZ_API int32_t nz_kernel_filter_table(int32_t filter_type, float* kx, float* kz, int32_t* ksize, float* factor) {
    NZ_REQUIRE(kx && kz && ksize && factor, "nz_kernel_filter_table: null output");
    return kernel_filter_table(filter_type, kx, kz, ksize, factor);
}

NZ_API int32_t nz_tile_geometry(int32_t tile_resolution, int32_t tile_size, int32_t margin, int32_t* mesh_resolution,
                                int32_t* margin_pix, float* mesh_tile_size) {
    NZ_REQUIRE(tile_resolution > 0 && tile_size > 0, "nz_tile_geometry: bad tile");
    // calcTotalResolution / calcMarginVerts / calculateMarginWS, Scripts/MeshTileGenerator.cs:166-177
    double patchRes = (tile_resolution * 1.0) / tile_size;
    int total = tile_resolution + (2 * (int)(float)(margin * patchRes));
    int mv = (int)((total - tile_resolution) / 2);
    float marginWS = mv * (float)((tile_size * 1.0) / tile_resolution);
    if (mesh_resolution) *mesh_resolution = total;
    if (margin_pix) *margin_pix = mv;
    if (mesh_tile_size) *mesh_tile_size = tile_size + (2 * marginWS);  // RequestMesh, :197-206
    return NZ_OK;
}

// =================================================================================================
// device layer
// =================================================================================================
NZ_API int32_t nz_dev_fractal(float* d_dst, int32_t width, int32_t rows, int32_t z_first, int32_t noise_type,
                              float hurst, float starting_amplitude, float stepdown, float detune_rate,
                              int32_t octaves, int32_t xpos, int32_t zpos, int32_t noise_size, void* stream) {
    NZ_REQUIRE(d_dst, "nz_dev_fractal: null destination");
    FractalParams p;
    int32_t rc = fractal_params(&p, width, rows, z_first, noise_type, hurst, starting_amplitude, stepdown, detune_rate,
                                octaves, xpos, zpos, noise_size);
    if (rc != NZ_OK) return rc;
    rc = ensure_runtime();
    if (rc != NZ_OK) return rc;
    if ((rc = check_device_pointer(d_dst, "nz_dev_fractal")) != NZ_OK) return rc;
    return launch_fractal(d_dst, noise_type, p, (cudaStream_t)stream);
}

NZ_API int32_t nz_dev_separable(float* d_data, float* d_tmp, int32_t width, int32_t rows, int32_t ksize,
                                const float* h_kx, const float* h_kz, float factor, int32_t iterations,
                                float** d_result, void* stream) {
    int32_t rc = ensure_runtime();
    if (rc != NZ_OK) return rc;
    if ((rc = check_device_pointer(d_data, "nz_dev_separable")) != NZ_OK) return rc;
    return launch_separable(d_data, d_tmp, width, rows, ksize, h_kx, h_kz, factor, iterations, d_result,
                            (cudaStream_t)stream);
}

NZ_API int32_t nz_dev_kernel_filter(float* d_data, float* d_tmp, int32_t width, int32_t rows, int32_t filter_type,
                                    int32_t iterations, float** d_result, void* stream) {
    int32_t rc = ensure_runtime();
    if (rc != NZ_OK) return rc;
    if ((rc = check_device_pointer(d_data, "nz_dev_kernel_filter")) != NZ_OK) return rc;
    NZ_REQUIRE(filter_type >= 0 && filter_type < NZ_FILTER__COUNT, "kernel_filter: filter_type %d out of range", filter_type);
    if (filter_type == NZ_FILTER_SOBEL3_2D)
        return launch_sobel2d(d_data, d_tmp, width, rows, iterations, d_result, (cudaStream_t)stream);
    float kx[9], kz[9], factor;
    int ksize;
    rc = kernel_filter_table(filter_type, kx, kz, &ksize, &factor);
    if (rc != NZ_OK) return rc;
    return launch_separable(d_data, d_tmp, width, rows, ksize, kx, kz, factor, iterations, d_result, (cudaStream_t)stream);
}

NZ_API int32_t nz_dev_min_erosion(float* d_data, float* d_tmp, int32_t width, int32_t rows, int32_t iterations,
                                  float** d_result, void* stream) {
    int32_t rc = ensure_runtime();
    if (rc != NZ_OK) return rc;
    if ((rc = check_device_pointer(d_data, "nz_dev_min_erosion")) != NZ_OK) return rc;
    return launch_min_erosion(d_data, d_tmp, width, rows, iterations, d_result, (cudaStream_t)stream);
}

NZ_API size_t nz_dev_flowmap_scratch_bytes(int32_t width, int32_t rows, int32_t iterations) {
    if (width <= 0 || rows <= 0) return 0;
    return flowmap_scratch_bytes(width, rows, iterations);
}

NZ_API int32_t nz_dev_flowmap(float* d_height, float* d_tmp, void* d_scratch, int32_t width, int32_t rows,
                              int32_t iterations, float norm_min, float norm_max, float** d_result, void* stream) {
    int32_t rc = ensure_runtime();
    if (rc != NZ_OK) return rc;
    if ((rc = check_device_pointer(d_height, "nz_dev_flowmap")) != NZ_OK) return rc;
    return launch_flowmap(d_height, d_tmp, d_scratch, width, rows, iterations, norm_min, norm_max, d_result,
                          (cudaStream_t)stream);
}

NZ_API int32_t nz_dev_flow_walk_reruns(uint64_t* count) {
    NZ_REQUIRE(count != nullptr, "nz_dev_flow_walk_reruns: null result pointer");
    int32_t rc = ensure_runtime();
    if (rc != NZ_OK) return rc;
    unsigned long long c = 0;
    rc = flow_walk_reruns(&c);
    *count = c;
    return rc;
}

NZ_API int32_t nz_dev_heightmap_mesh(int32_t mesh_type, void* d_vertices, uint32_t* d_indices, int32_t resolution,
                                     int32_t input_resolution, int32_t margin_pix, float tile_height, float tile_size,
                                     const float* d_heights, int32_t h_row_first, int32_t h_rows, int32_t vz_begin,
                                     int32_t vz_end, void* stream) {
    (void)margin_pix;  // consumed only by MarginScale(), which nothing calls (SquareGridHeightMap.cs:41-56)
    int32_t rc = ensure_runtime();
    if (rc != NZ_OK) return rc;
    if ((rc = check_device_pointer(d_heights, "nz_dev_heightmap_mesh")) != NZ_OK) return rc;
    return launch_mesh(mesh_type, d_vertices, d_indices, resolution, input_resolution, tile_height, tile_size, d_heights,
                       h_row_first, h_rows, vz_begin, vz_end, (cudaStream_t)stream);
}

NZ_API int32_t nz_dev_thermal_erosion(float* d_data, float* d_tmp, int32_t resolution, float talus, float increment_ratio,
                                      float mesh_height_width_ratio, int32_t iterations, float** d_result, void* stream) {
    NZ_REQUIRE(d_data && resolution > 0, "nz_dev_thermal_erosion: bad arguments");
    NZ_REQUIRE(iterations >= 0, "nz_dev_thermal_erosion: iterations %d < 0", iterations);
    NZ_REQUIRE(d_tmp != d_data, "nz_dev_thermal_erosion: d_tmp aliases d_data");
    int32_t rc = ensure_runtime();
    if (rc != NZ_OK) return rc;
    if ((rc = check_device_pointer(d_data, "nz_dev_thermal_erosion")) != NZ_OK) return rc;
    return launch_thermal_erosion(d_data, d_tmp, resolution, talus, increment_ratio, mesh_height_width_ratio, iterations, d_result,
                                  (cudaStream_t)stream);
}
NZ_API size_t nz_dev_subtractive_flow_scratch_bytes(int32_t width, int32_t rows) {
    if (width <= 0 || rows <= 0) return 0;
    return subtractive_flow_scratch_bytes(width, rows);
}
NZ_API int32_t nz_dev_subtractive_flow_erosion(float* d_height, void* d_scratch, int32_t width, int32_t rows,
                                               int32_t erosive_iterations, float erosive_factor, float norm_min,
                                               float norm_max, void* stream) {
    int32_t rc = ensure_runtime();
    if (rc != NZ_OK) return rc;
    if ((rc = check_device_pointer(d_height, "nz_dev_subtractive_flow_erosion")) != NZ_OK) return rc;
    return launch_subtractive_flow_erosion(d_height, d_scratch, width, rows, erosive_iterations, erosive_factor, norm_min,
                                           norm_max, (cudaStream_t)stream);
}
NZ_API int32_t nz_dev_constant(float* d_data, size_t n, int32_t operation, float constant_value, void* stream) {
    NZ_REQUIRE(d_data || n == 0, "nz_dev_constant: null grid");
    int32_t rc = ensure_runtime();
    if (rc != NZ_OK) return rc;
    if ((rc = check_device_pointer(d_data, "nz_dev_constant")) != NZ_OK) return rc;
    return launch_constant(d_data, n, operation, constant_value, (cudaStream_t)stream);
}
NZ_API int32_t nz_dev_reduce(float* d_left, const float* d_right, size_t n, int32_t operation, void* stream) {
    NZ_REQUIRE((d_left && d_right) || n == 0, "nz_dev_reduce: null grid");
    int32_t rc = ensure_runtime();
    if (rc != NZ_OK) return rc;
    if ((rc = check_device_pointer(d_left, "nz_dev_reduce")) != NZ_OK) return rc;
    return launch_reduce(d_left, d_right, n, operation, (cudaStream_t)stream);
}
NZ_API int32_t nz_dev_curve(float* d_data, size_t n, const float* d_curve, int32_t curve_size, void* stream) {
    NZ_REQUIRE(d_data || n == 0, "nz_dev_curve: null grid");
    NZ_REQUIRE(d_curve && curve_size >= 2, "nz_dev_curve: the curve needs at least 2 samples (got %d)", curve_size);
    int32_t rc = ensure_runtime();
    if (rc != NZ_OK) return rc;
    if ((rc = check_device_pointer(d_data, "nz_dev_curve")) != NZ_OK) return rc;
    return launch_curve(d_data, n, d_curve, curve_size, (cudaStream_t)stream);
}
NZ_API int32_t nz_dev_crop(const float* d_input, int32_t input_resolution, float* d_output, int32_t output_resolution,
                           int32_t offset, void* stream) {
    NZ_REQUIRE(d_input && d_output && input_resolution > 0 && output_resolution > 0 && output_resolution <= 65535,
               "nz_dev_crop: bad arguments");
    int32_t rc = ensure_runtime();
    if (rc != NZ_OK) return rc;
    if ((rc = check_device_pointer(d_input, "nz_dev_crop")) != NZ_OK) return rc;
    return launch_crop(d_input, input_resolution, d_output, output_resolution, offset, (cudaStream_t)stream);
}
NZ_API size_t nz_dev_map_range_scratch_bytes(void) { return map_range_scratch_bytes(); }
NZ_API int32_t nz_dev_map_range(const float* d_map, size_t n, float lim_min, float lim_max, float* d_res3, void* d_scratch,
                                void* stream) {
    NZ_REQUIRE(d_map && d_res3 && d_scratch && n > 0, "nz_dev_map_range: bad arguments");
    int32_t rc = ensure_runtime();
    if (rc != NZ_OK) return rc;
    if ((rc = check_device_pointer(d_map, "nz_dev_map_range")) != NZ_OK) return rc;
    return launch_map_range(d_map, n, lim_min, lim_max, d_res3, d_scratch, (cudaStream_t)stream);
}
NZ_API int32_t nz_dev_normalize(float* d_data, size_t n, float vmin, float range, void* stream) {
    NZ_REQUIRE(d_data || n == 0, "nz_dev_normalize: null grid");
    int32_t rc = ensure_runtime();
    if (rc != NZ_OK) return rc;
    if ((rc = check_device_pointer(d_data, "nz_dev_normalize")) != NZ_OK) return rc;
    return launch_normalize(d_data, n, vmin, range, (cudaStream_t)stream);
}

NZ_API int32_t nz_dev_fma_peak(float* d_sink, int32_t grid, int32_t iters, double* flops, void* stream) {
    int32_t rc = ensure_runtime();
    if (rc != NZ_OK) return rc;
    return launch_fma_peak(d_sink, grid, iters, flops, (cudaStream_t)stream);
}

// =================================================================================================
// host layer
// =================================================================================================
// ---- residency scopes ---------------------------------------------------------------------------------
static int32_t scope_close(const std::shared_ptr<Scope>& sc) {
    int32_t rc = thread_ready();
    if (rc != NZ_OK) return rc;
    std::lock_guard<std::mutex> lk(sc->mu);
    cudaStream_t s = t_state.stream;
    for (auto& kv : sc->mirrors) {
        Mirror& m = kv.second;
        if (m.bands) {
            if (m.dirty && rc == NZ_OK) rc = banded_copy_rows(m, /*to_host=*/true);   // one D2H per band, in parallel
            continue;
        }
        if (m.ready && m.last_stream != s) cudaStreamWaitEvent(s, m.ready, 0);
        if (m.dirty && rc == NZ_OK) rc = download(m);
    }
    cudaError_t e = cudaStreamSynchronize(s);
    for (auto& kv : sc->mirrors) release_mirror(kv.second);   // ~BandSet synchronises the bands' streams
    sc->mirrors.clear();
    if (rc != NZ_OK) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "scope close");
    return NZ_OK;
}

NZ_API int64_t nz_scope_create(void) {
    int32_t rc = ensure_init();
    if (rc != NZ_OK) return rc;
    const long long id = g_next_scope.fetch_add(1);
    std::lock_guard<std::mutex> lk(g_scopes_mu);
    g_scopes[id] = std::make_shared<Scope>();
    return id;
}

NZ_API int32_t nz_scope_enter(int64_t id) {
    int32_t rc = thread_ready();
    if (rc != NZ_OK) return rc;
    if (in_scope()) {
        set_error("nz_scope_enter: this thread is already inside scope %lld", t_state.scope_id);
        return NZ_E_STATE;
    }
    std::lock_guard<std::mutex> lk(g_scopes_mu);
    auto it = g_scopes.find(id);
    if (it == g_scopes.end()) {
        set_error("nz_scope_enter: unknown scope %lld", (long long)id);
        return NZ_E_INVALID;
    }
    t_state.scope = it->second;
    t_state.scope_id = id;
    t_state.owns_scope = false;
    return NZ_OK;
}

NZ_API int32_t nz_scope_leave(void) {
    if (!in_scope()) {
        set_error("nz_scope_leave: this thread is in no scope");
        return NZ_E_STATE;
    }
    t_state.scope.reset();
    t_state.scope_id = 0;
    t_state.owns_scope = false;
    return NZ_OK;
}

NZ_API int32_t nz_scope_close(int64_t id) {
    std::shared_ptr<Scope> sc;
    {
        std::lock_guard<std::mutex> lk(g_scopes_mu);
        auto it = g_scopes.find(id);
        if (it == g_scopes.end()) {
            set_error("nz_scope_close: unknown scope %lld", (long long)id);
            return NZ_E_INVALID;
        }
        sc = it->second;
        g_scopes.erase(it);
    }
    if (t_state.scope_id == id) {
        t_state.scope.reset();
        t_state.scope_id = 0;
        t_state.owns_scope = false;
    }
    return scope_close(sc);
}

NZ_API int32_t nz_pipeline_begin(void) {
    int32_t rc = thread_ready();
    if (rc != NZ_OK) return rc;
    if (in_scope()) {
        set_error("nz_pipeline_begin: already inside a pipeline on this thread");
        return NZ_E_STATE;
    }
    const int64_t id = nz_scope_create();
    if (id < 0) return (int32_t)id;
    rc = nz_scope_enter(id);
    if (rc == NZ_OK) t_state.owns_scope = true;
    return rc;
}

NZ_API int32_t nz_flush_to_host(const float* host_ptr) {
    int32_t rc = thread_ready();
    if (rc != NZ_OK) return rc;
    Scope& sc = cur_scope();
    std::lock_guard<std::mutex> lk(sc.mu);
    auto it = sc.mirrors.find(host_ptr);
    if (it == sc.mirrors.end()) return NZ_OK;  // nothing resident: host is current
    if (it->second.bands) {
        if (it->second.dirty && (rc = banded_copy_rows(it->second, /*to_host=*/true)) != NZ_OK) return rc;
        return bandset_sync(*it->second.bands);
    }
    if (it->second.ready && it->second.last_stream != t_state.stream) NZ_CUDA(cudaStreamWaitEvent(t_state.stream, it->second.ready, 0));
    if (it->second.dirty) {
        rc = download(it->second);
        if (rc != NZ_OK) return rc;
    }
    NZ_CUDA(cudaStreamSynchronize(t_state.stream));
    return NZ_OK;
}

NZ_API int32_t nz_pipeline_end(void) {
    if (!in_scope() || !t_state.owns_scope) {
        set_error("nz_pipeline_end: no pipeline open on this thread");
        return NZ_E_STATE;
    }
    return nz_scope_close(t_state.scope_id);
}

NZ_API int32_t nz_pin(void* host_ptr, size_t bytes) {
    int32_t rc = ensure_init();
    if (rc != NZ_OK) return rc;
    NZ_REQUIRE(host_ptr && bytes, "nz_pin: null/empty range");
    NZ_CUDA(cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable));   // pinned for every device (banded mirrors)
    return NZ_OK;
}
NZ_API int32_t nz_unpin(void* host_ptr) {
    int32_t rc = ensure_init();
    if (rc != NZ_OK) return rc;
    NZ_CUDA(cudaHostUnregister(host_ptr));
    return NZ_OK;
}

NZ_API int32_t nz_fractal(nz_slice_f32 dst, int32_t resolution, int32_t noise_type, float hurst,
                          float starting_amplitude, float stepdown, float detune_rate, int32_t octaves, int32_t xpos,
                          int32_t zpos, int32_t noise_size) {
    NZ_REQUIRE(resolution > 0 && resolution <= 46340, "nz_fractal: resolution %d out of range", resolution);
    int32_t rc = check_slice(dst, (long long)resolution * resolution, "nz_fractal");
    if (rc != NZ_OK) return rc;
    FractalParams p;
    rc = fractal_params(&p, resolution, resolution, 0, noise_type, hurst, starting_amplitude, stepdown, detune_rate,
                        octaves, xpos, zpos, noise_size);
    if (rc != NZ_OK) return rc;
    UnscopedCleanup cleanup;
    nvtxRangePushA("nz_fractal");
    struct Pop { ~Pop() { nvtxRangePop(); } } pop;
    if ((rc = begin_stage()) != NZ_OK) return rc;
    Mirror* m;
    if ((rc = acquire(dst, /*need_contents=*/false, &m, resolution)) != NZ_OK) return rc;
    if ((rc = mark_uploaded()) != NZ_OK) return rc;
    if (m->bands)
        rc = bandset_fractal(*m->bands, noise_type, hurst, starting_amplitude, stepdown, detune_rate, octaves, xpos, zpos, noise_size, 0, 0);
    else
        rc = launch_fractal(m->d, noise_type, p, t_state.stream);
    if (rc == NZ_OK) m->dirty = true;
    int32_t rc2 = finish(m);
    return rc != NZ_OK ? rc : rc2;
}

}  // extern "C"

// ghost rows a banded mirror of a resolution^2 grid keeps per band side under the current nz_set_bands (0: not banded)
static int band_cap_for(int resolution) {
    int n_bands;
    {
        std::lock_guard<std::mutex> lk(g_ctx.mu);
        n_bands = g_ctx.n_bands;
    }
    if (n_bands < 2 || resolution < NZ_BANDS_MIN_RESOLUTION || resolution / n_bands < BAND_MIN_ROWS) return 0;
    const int min_own = resolution / n_bands;
    return BAND_GHOST_CAP < min_own ? BAND_GHOST_CAP : min_own;
}

// one element-wise step on an ordinary mirror: deferred inside a scope (see materialize()), applied at once outside
static int32_t pointwise_step(Mirror& m, const PointwiseOp& op) {
    if (in_scope() && pointwise_defer_enabled() && m.pending.size() + 1 < (size_t)PW_CHAIN_MAX) {
        m.pending.push_back(op);
        m.dirty = true;
        return NZ_OK;
    }
    return materialize(m, &op);
}

// shared body of the in-place two-buffer stages.  `banded_body` (BandSet&) -> status is the stage's form on row bands
// (nz_set_bands); pass can_band = false for a stage (or a parameter combination) that has none: a banded mirror is then
// gathered onto the primary device first.
template <typename F, typename B>
static int32_t run_inplace_stage(nz_slice_f32 src, int32_t resolution, const char* who, bool need_tmp, F&& body, bool can_band,
                                 B&& banded_body, bool keep_pending = false) {
    NZ_REQUIRE(resolution > 0 && resolution <= 46340, "%s: resolution %d out of range", who, resolution);
    int32_t rc = check_slice(src, (long long)resolution * resolution, who);
    if (rc != NZ_OK) return rc;
    UnscopedCleanup cleanup;
    nvtxRangePushA(who);
    struct Pop { ~Pop() { nvtxRangePop(); } } pop;
    if ((rc = begin_stage()) != NZ_OK) return rc;
    Mirror* m;
    if ((rc = acquire(src, /*need_contents=*/true, &m, can_band ? resolution : 0, keep_pending)) != NZ_OK) return rc;
    if (m->bands) {
        if ((rc = mark_uploaded()) != NZ_OK) return rc;
        rc = banded_body(*m->bands);
        if (rc == NZ_OK) m->dirty = true;
        int32_t rc2 = finish(m);
        return rc != NZ_OK ? rc : rc2;
    }
    if (need_tmp && (rc = ensure_tmp(*m)) != NZ_OK) return rc;
    if ((rc = mark_uploaded()) != NZ_OK) return rc;
    float* result = m->d;
    rc = body(*m, &result);
    if (rc == NZ_OK) adopt_result(*m, result);
    int32_t rc2 = finish(m);
    return rc != NZ_OK ? rc : rc2;
}
template <typename F>
static int32_t run_inplace_stage(nz_slice_f32 src, int32_t resolution, const char* who, bool need_tmp, F&& body) {
    return run_inplace_stage(src, resolution, who, need_tmp, body, false, [](BandSet&) { return (int32_t)NZ_E_UNSUPPORTED; });
}

extern "C" {

NZ_API int32_t nz_separable(nz_slice_f32 src, nz_slice_f32 tmp, int32_t ksize, const float* kx, const float* kz,
                            float factor, int32_t resolution, int32_t iterations) {
    (void)tmp;
    NZ_REQUIRE(ksize >= 1 && (ksize & 1) && ksize <= NZ_MAX_KERNEL_WIDTH && kx && kz && iterations >= 0,
               "nz_separable: bad kernel (ksize %d, iterations %d)", ksize, iterations);
    return run_inplace_stage(src, resolution, "nz_separable", true, [&](Mirror& m, float** res) {
        return launch_separable(m.d, m.d_tmp, resolution, resolution, ksize, kx, kz, factor, iterations, res, t_state.stream);
    }, true, [&](BandSet& bs) { return bandset_separable(bs, ksize, kx, kz, factor, iterations, 0, 0, true); });
}

NZ_API int32_t nz_kernel_filter(nz_slice_f32 src, nz_slice_f32 tmp, int32_t filter_type, int32_t resolution,
                                int32_t iterations) {
    (void)tmp;
    NZ_REQUIRE(filter_type >= 0 && filter_type < NZ_FILTER__COUNT, "nz_kernel_filter: filter_type %d out of range", filter_type);
    NZ_REQUIRE(iterations >= 1, "nz_kernel_filter: iterations %d < 1", iterations);
    return run_inplace_stage(src, resolution, "nz_kernel_filter", true, [&](Mirror& m, float** res) {
        return nz_dev_kernel_filter(m.d, m.d_tmp, resolution, resolution, filter_type, iterations, res, t_state.stream);
    }, true, [&](BandSet& bs) {
        if (filter_type == NZ_FILTER_SOBEL3_2D) return bandset_sobel2d(bs, iterations, 0, 0, true);
        float tkx[9], tkz[9], factor;
        int ksize;
        int32_t r = kernel_filter_table(filter_type, tkx, tkz, &ksize, &factor);
        if (r != NZ_OK) return r;
        return bandset_separable(bs, ksize, tkx, tkz, factor, iterations, 0, 0, true);
    });
}

NZ_API int32_t nz_gauss_filter(nz_slice_f32 src, nz_slice_f32 tmp, int32_t width, int32_t sigma, int32_t resolution,
                               int32_t iterations) {
    (void)tmp;
    float k[NZ_MAX_KERNEL_WIDTH];
    int32_t w = 0;
    int32_t rc = nz_gauss_kernel(sigma, width, k, &w);
    if (rc != NZ_OK) return rc;
    return nz_separable(src, tmp, w, k, k, 1.0f, resolution, iterations);
}

NZ_API int32_t nz_smooth_filter(nz_slice_f32 src, nz_slice_f32 tmp, int32_t width, int32_t resolution, int32_t iterations) {
    // SmoothBlur.GetKernel, BlurKernels.cs:38-44 (after StageSmoothBlur's limitWidth)
    int32_t w = nz_limit_width(width);
    float k[NZ_MAX_KERNEL_WIDTH];
    for (int i = 0; i < w; i++) k[i] = 1.0f / w;
    return nz_separable(src, tmp, w, k, k, 1.0f, resolution, iterations);
}

NZ_API int32_t nz_min_erosion(nz_slice_f32 src, int32_t resolution, int32_t iterations) {
    NZ_REQUIRE(iterations >= 0, "nz_min_erosion: iterations %d < 0", iterations);
    return run_inplace_stage(src, resolution, "nz_min_erosion", true, [&](Mirror& m, float** res) {
        return launch_min_erosion(m.d, m.d_tmp, resolution, resolution, iterations, res, t_state.stream);
    }, true, [&](BandSet& bs) { return bandset_min_erosion(bs, iterations, 0, 0, true); });
}

NZ_API int32_t nz_flowmap(nz_slice_f32 height, int32_t resolution, int32_t iterations, float norm_min, float norm_max) {
    NZ_REQUIRE(iterations >= 0, "nz_flowmap: iterations %d < 0", iterations);
    void* scratch = nullptr;
    int32_t rc = run_inplace_stage(height, resolution, "nz_flowmap", true, [&](Mirror& m, float** res) {
        const size_t need = flowmap_scratch_bytes(resolution, resolution, iterations);
        if (need) {
            int32_t r = pool_alloc(&scratch, need);
            if (r != NZ_OK) return r;
        }
        return launch_flowmap(m.d, m.d_tmp, scratch, resolution, resolution, iterations, norm_min, norm_max, res, t_state.stream);
    }, /*can_band=*/2 * iterations + 1 <= band_cap_for(resolution), [&](BandSet& bs) {
        return bandset_flowmap(bs, iterations, norm_min, norm_max, 0, 0, true);
    });
    if (scratch) {
        // the stream may still be using it inside a pipeline: order the release after the work
        if (in_scope()) cudaStreamSynchronize(t_state.stream);
        pool_free(scratch);
    }
    return rc;
}

// The mesh of a large tile is 72 bytes per cell going to the HOST, i.e. PCIe time, not kernel time.  The mesh is therefore
// built in chunks of vertex rows on the compute stream while the previous chunk goes home on a copy stream: two staging
// slots of ~128 MB instead of a (R+1)^2 x 48 B + 6 R^2 x 4 B device buffer (19.3 GB at R = 16376), and the kernel time
// disappears under the download.
static int32_t mesh_chunked(int mesh_type, void* vertices, uint32_t* indices, int R, int in_res, float tile_height, float tile_size,
                            const float* d_heights) {
    cudaStream_t s = t_state.stream;
    if (!t_state.copy_stream) {
        NZ_CUDA(cudaStreamCreateWithFlags(&t_state.copy_stream, cudaStreamNonBlocking));
        for (auto& e : t_state.mesh_ev) NZ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    cudaStream_t cs = t_state.copy_stream;
    const int VR = R + 1;
    const size_t row_v = (size_t)VR * NZ_MESH_VERTEX_BYTES, row_i = (size_t)6 * R * sizeof(uint32_t);
    long long per = (128ll << 20) / (long long)(row_v + row_i);
    const int rows_per_chunk = (int)(per < 1 ? 1 : (per > VR ? VR : per));
    const int n_chunks = cdiv(VR, rows_per_chunk);
    void* d_v[2] = {nullptr, nullptr};
    void* d_i[2] = {nullptr, nullptr};
    int32_t rc = NZ_OK;
    const int n_slots = n_chunks > 1 ? 2 : 1;
    for (int k = 0; k < n_slots && rc == NZ_OK; k++) {
        rc = pool_alloc(&d_v[k], rows_per_chunk * row_v);
        if (rc == NZ_OK) rc = pool_alloc(&d_i[k], rows_per_chunk * row_i);
    }
    cudaError_t e = cudaSuccess;
    for (int c = 0; c < n_chunks && rc == NZ_OK && e == cudaSuccess; c++) {
        const int slot = c & 1;
        const int vz0 = c * rows_per_chunk, vz1 = vz0 + rows_per_chunk < VR ? vz0 + rows_per_chunk : VR;
        const int t0 = vz0 > 1 ? vz0 : 1;
        if (c >= 2) e = cudaStreamWaitEvent(s, t_state.mesh_ev[2 + slot], 0);      // the slot's previous chunk has gone home
        if (e != cudaSuccess) break;
        rc = launch_mesh(mesh_type, d_v[slot], (uint32_t*)d_i[slot], R, in_res, tile_height, tile_size, d_heights, 0, in_res, vz0, vz1, s);
        if (rc != NZ_OK) break;
        e = cudaEventRecord(t_state.mesh_ev[slot], s);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(cs, t_state.mesh_ev[slot], 0);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync((char*)vertices + (size_t)vz0 * row_v, d_v[slot], (size_t)(vz1 - vz0) * row_v, cudaMemcpyDeviceToHost, cs);
        if (e == cudaSuccess && vz1 > t0)
            e = cudaMemcpyAsync((char*)indices + (size_t)(t0 - 1) * row_i, d_i[slot], (size_t)(vz1 - t0) * row_i, cudaMemcpyDeviceToHost, cs);
        if (e == cudaSuccess) e = cudaEventRecord(t_state.mesh_ev[2 + slot], cs);
    }
    if (e == cudaSuccess && rc == NZ_OK) e = cudaEventRecord(t_state.ev[2], s);
    cudaError_t e2 = cudaStreamSynchronize(cs);
    cudaError_t e3 = cudaStreamSynchronize(s);
    for (int k = 0; k < 2; k++) {
        pool_free(d_v[k]);
        pool_free(d_i[k]);
    }
    if (rc != NZ_OK) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "nz_heightmap_mesh");
    if (e2 != cudaSuccess) return cuda_fail(e2, "nz_heightmap_mesh (download)");
    if (e3 != cudaSuccess) return cuda_fail(e3, "nz_heightmap_mesh");
    return NZ_OK;
}

// heights live on row bands (nz_set_bands): every band builds its slice of the mesh and downloads it over its own PCIe link
static int32_t mesh_banded(Mirror& m, int mesh_type, void* vertices, uint32_t* indices, int R, float tile_height, float tile_size) {
    BandSet& bs = *m.bands;
    int32_t rc = bandset_exchange(bs, 1, 1);
    if (rc == NZ_OK) rc = bandset_mesh(bs, mesh_type, R, tile_height, tile_size);
    if (rc != NZ_OK) return rc;
    const size_t row_v = (size_t)(R + 1) * NZ_MESH_VERTEX_BYTES, row_i = (size_t)6 * R * sizeof(uint32_t);
    int prev = 0;
    NZ_CUDA(cudaGetDevice(&prev));
    cudaError_t e = cudaSuccess;
    for (Band& bd : bs.b) {
        if (bd.vz1 <= bd.vz0) continue;
        const int t0 = bd.vz0 > 1 ? bd.vz0 : 1;
        e = cudaSetDevice(bd.device);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync((char*)vertices + (size_t)bd.vz0 * row_v, bd.vtx, (size_t)(bd.vz1 - bd.vz0) * row_v, cudaMemcpyDeviceToHost, bd.s);
        if (e == cudaSuccess && bd.vz1 > t0)
            e = cudaMemcpyAsync((char*)indices + (size_t)(t0 - 1) * row_i, bd.idx, (size_t)(bd.vz1 - t0) * row_i, cudaMemcpyDeviceToHost, bd.s);
        if (e == cudaSuccess) e = cudaEventRecord(bd.done, bd.s);
        if (e != cudaSuccess) break;
    }
    cudaSetDevice(prev);
    if (e != cudaSuccess) return cuda_fail(e, "nz_heightmap_mesh (banded download)");
    rc = bandset_sync(bs);
    // the mesh slices are per-call staging: give them back (19.3 GB / n_bands per device at R = 16376)
    for (Band& bd : bs.b) {
        dev_free(bd.vtx);
        dev_free(bd.idx);
        bd.vtx = nullptr;
        bd.idx = nullptr;
        bd.vtx_bytes = bd.idx_bytes = 0;
    }
    return rc;
}

NZ_API int32_t nz_heightmap_mesh(int32_t mesh_type, void* vertices, uint32_t* indices, int32_t resolution,
                                 int32_t input_resolution, int32_t margin_pix, float tile_height, float tile_size,
                                 nz_slice_f32 heights) {
    (void)margin_pix;
    NZ_REQUIRE(vertices && indices, "nz_heightmap_mesh: null output buffer");
    NZ_REQUIRE(resolution > 0 && input_resolution > 0 && input_resolution <= 46340, "nz_heightmap_mesh: bad resolution");
    int32_t rc = check_slice(heights, (long long)input_resolution * input_resolution, "nz_heightmap_mesh");
    if (rc != NZ_OK) return rc;
    UnscopedCleanup cleanup;
    nvtxRangePushA("nz_heightmap_mesh");
    struct Pop { ~Pop() { nvtxRangePop(); } } pop;
    if ((rc = begin_stage()) != NZ_OK) return rc;
    Mirror* m;
    if ((rc = acquire(heights, /*need_contents=*/true, &m, input_resolution)) != NZ_OK) return rc;
    if ((rc = mark_uploaded()) != NZ_OK) return rc;
    if (m->bands) {
        rc = mesh_banded(*m, mesh_type, vertices, indices, resolution, tile_height, tile_size);
        if (rc == NZ_OK) rc = banded_join(*m);
        if (rc == NZ_OK) NZ_CUDA(cudaEventRecord(t_state.ev[2], t_state.stream));
    } else {
        rc = mesh_chunked(mesh_type, vertices, indices, resolution, input_resolution, tile_height, tile_size, m->d);
    }
    if (rc != NZ_OK) return rc;     // heights were only read: outside a scope the cleanup guard drops their mirror
    NZ_CUDA(cudaEventRecord(t_state.ev[3], t_state.stream));
    t_state.timed = true;
    t_state.launches = (int)(g_launches.load() - t_state.launches_at_start);
    return NZ_OK;
}

}  // extern "C"

// ---- the section-8f rows ----------------------------------------------------------------------------------
extern "C" {

NZ_API int32_t nz_thermal_erosion(nz_slice_f32 src, float talus, float increment_ratio, float mesh_height_width_ratio,
                                  int32_t iterations, int32_t resolution) {
    NZ_REQUIRE(iterations >= 0, "nz_thermal_erosion: iterations %d < 0", iterations);
    return run_inplace_stage(src, resolution, "nz_thermal_erosion", true, [&](Mirror& m, float** res) {
        return launch_thermal_erosion(m.d, m.d_tmp, resolution, talus, increment_ratio, mesh_height_width_ratio, iterations, res,
                                      t_state.stream);
    });
}

NZ_API int32_t nz_subtractive_flow_erosion(nz_slice_f32 height, int32_t resolution, int32_t erosive_iterations,
                                           float erosive_factor, float norm_min, float norm_max) {
    NZ_REQUIRE(erosive_iterations >= 0, "nz_subtractive_flow_erosion: erosiveIterations %d < 0", erosive_iterations);
    void* scratch = nullptr;
    int32_t rc = run_inplace_stage(height, resolution, "nz_subtractive_flow_erosion", false, [&](Mirror& m, float**) {
        if (erosive_iterations > 0) {
            int32_t r = pool_alloc(&scratch, subtractive_flow_scratch_bytes(resolution, resolution));
            if (r != NZ_OK) return r;
        }
        return launch_subtractive_flow_erosion(m.d, scratch, resolution, resolution, erosive_iterations, erosive_factor,
                                               norm_min, norm_max, t_state.stream);
    });
    if (scratch) {
        if (in_scope()) cudaStreamSynchronize(t_state.stream);   // order the release after the work (see nz_flowmap)
        pool_free(scratch);
    }
    return rc;
}

NZ_API int32_t nz_constant(nz_slice_f32 src, nz_slice_f32 tmp, int32_t operation, float constant_value, int32_t resolution) {
    (void)tmp;
    NZ_REQUIRE(operation >= 0 && operation < NZ_CONSTANT__COUNT, "nz_constant: operation %d out of range", operation);
    const PointwiseOp op{operation == NZ_CONSTANT_MULTIPLY ? PW_MUL : PW_BINARIZE, constant_value, 0.0f, nullptr};
    return run_inplace_stage(src, resolution, "nz_constant", false, [&](Mirror& m, float**) {
        return pointwise_step(m, op);
    }, true, [&](BandSet& bs) {
        return bandset_stage(bs, 0, 0, [&](Band& bd, float* cur, float*, int rows, int, float**) {
            return launch_constant(cur, (size_t)rows * bs.width, operation, constant_value, bd.s);
        });
    }, /*keep_pending=*/true);
}

NZ_API int32_t nz_normalize(nz_slice_f32 src, nz_slice_f32 tmp, const float* args3, int32_t resolution) {
    (void)tmp;
    NZ_REQUIRE(args3 != nullptr, "nz_normalize: args is null");
    const float vmin = args3[0], range = args3[2];
    const PointwiseOp op{PW_NORMALIZE, vmin, range, nullptr};
    return run_inplace_stage(src, resolution, "nz_normalize", false, [&](Mirror& m, float**) {
        return pointwise_step(m, op);
    }, true, [&](BandSet& bs) {
        return bandset_stage(bs, 0, 0, [&](Band& bd, float* cur, float*, int rows, int, float**) {
            return launch_normalize(cur, (size_t)rows * bs.width, vmin, range, bd.s);
        });
    }, /*keep_pending=*/true);
}

NZ_API int32_t nz_reduce(nz_slice_f32 left, nz_slice_f32 right, nz_slice_f32 tmp, int32_t operation, int32_t resolution) {
    (void)tmp;
    NZ_REQUIRE(operation >= 0 && operation < NZ_REDUCE__COUNT, "nz_reduce: operation %d out of range", operation);
    NZ_REQUIRE(resolution > 0 && resolution <= 46340, "nz_reduce: resolution %d out of range", resolution);
    int32_t rc = check_slice(right, (long long)resolution * resolution, "nz_reduce(right)");
    if (rc != NZ_OK) return rc;
    NZ_REQUIRE(left.ptr != right.ptr, "nz_reduce: left and right are the same slice");
    Mirror* r = nullptr;
    rc = run_inplace_stage(left, resolution, "nz_reduce", false, [&](Mirror& m, float**) {
        int32_t rr = acquire(right, /*need_contents=*/true, &r);
        if (rr != NZ_OK) return rr;
        rr = launch_reduce(m.d, r->d, m.n, operation, t_state.stream);
        publish_read(r);
        return rr;
    });
    return rc;              // outside a scope run_inplace_stage's cleanup guard has released the right operand's mirror
}

NZ_API int32_t nz_curve(nz_slice_f32 src, nz_slice_f32 tmp, nz_slice_f32 curve, int32_t resolution) {
    (void)tmp;
    NZ_REQUIRE(curve.ptr && curve.length >= 2 && curve.stride_bytes >= 4, "nz_curve: the curve needs at least 2 samples");
    NZ_REQUIRE(curve.ptr != src.ptr, "nz_curve: curve and data are the same slice");
    Mirror* c = nullptr;
    // the curve closes a deferred run: its samples live in another mirror whose lifetime the data mirror does not control,
    // so the run (with the curve as its last step) is applied now, in one pass
    int32_t rc = run_inplace_stage(src, resolution, "nz_curve", false, [&](Mirror& m, float**) {
        int32_t rr = acquire(curve, /*need_contents=*/true, &c);
        if (rr != NZ_OK) return rr;
        const PointwiseOp op{PW_CURVE, (float)curve.length, 0.0f, c->d};
        rr = materialize(m, &op);
        publish_read(c);
        return rr;
    }, false, [](BandSet&) { return (int32_t)NZ_E_UNSUPPORTED; }, /*keep_pending=*/true);
    return rc;
}

NZ_API int32_t nz_crop(nz_slice_f32 input, int32_t input_resolution, nz_slice_f32 output, int32_t output_resolution,
                       int32_t offset) {
    NZ_REQUIRE(input_resolution > 0 && input_resolution <= 46340 && output_resolution > 0 && output_resolution <= 46340,
               "nz_crop: resolution out of range");
    int32_t rc = check_slice(input, (long long)input_resolution * input_resolution, "nz_crop(input)");
    if (rc != NZ_OK) return rc;
    if ((rc = check_slice(output, (long long)output_resolution * output_resolution, "nz_crop(output)")) != NZ_OK) return rc;
    NZ_REQUIRE(input.ptr != output.ptr, "nz_crop: input and output are the same slice");
    UnscopedCleanup cleanup;
    if ((rc = begin_stage()) != NZ_OK) return rc;
    Mirror *in = nullptr, *out = nullptr;
    if ((rc = acquire(input, /*need_contents=*/true, &in)) != NZ_OK) return rc;
    if ((rc = acquire(output, /*need_contents=*/false, &out)) != NZ_OK) return rc;
    if ((rc = mark_uploaded()) != NZ_OK) return rc;
    rc = launch_crop(in->d, input_resolution, out->d, output_resolution, offset, t_state.stream);
    publish_read(in);
    if (rc == NZ_OK) out->dirty = true;
    int32_t rc2 = finish(out);
    return rc != NZ_OK ? rc : rc2;
}

// ---- named device-resident buffers (PipelineStateManager on the GPU) ------------------------------------------------
}  // extern "C"

namespace {
struct NamedBuffer {
    float* d = nullptr;
    size_t n = 0;
    int device = 0;
    cudaEvent_t last = nullptr;     // recorded after every access: the next access, on whatever stream, waits for it
};
std::mutex g_named_mu;
std::unordered_map<std::string, NamedBuffer> g_named;

// called with g_named_mu held; (re)allocates the buffer for n floats on the current (primary) device
int32_t named_prepare(NamedBuffer& b, size_t n) {
    if (b.d && b.n != n) {
        if (b.last) cudaEventSynchronize(b.last);
        pool_free(b.d);
        b.d = nullptr;
    }
    if (!b.d) {
        int32_t rc = pool_alloc((void**)&b.d, n * sizeof(float));
        if (rc != NZ_OK) return rc;
        b.n = n;
        b.device = t_state.device;
    }
    if (!b.last) NZ_CUDA(cudaEventCreateWithFlags(&b.last, cudaEventDisableTiming));
    else NZ_CUDA(cudaStreamWaitEvent(t_state.stream, b.last, 0));
    return NZ_OK;
}
}  // namespace

extern "C" {

NZ_API int32_t nz_context_write(const char* name, nz_slice_f32 src) {
    NZ_REQUIRE(name && name[0], "nz_context_write: empty buffer name");
    NZ_REQUIRE(src.ptr && src.length > 0 && src.stride_bytes >= 4, "nz_context_write: empty slice");
    UnscopedCleanup cleanup;
    int32_t rc = begin_stage();
    if (rc != NZ_OK) return rc;
    Mirror* m = nullptr;
    if ((rc = acquire(src, /*need_contents=*/true, &m)) != NZ_OK) return rc;     // resident: no copy; else one H2D
    if ((rc = mark_uploaded()) != NZ_OK) return rc;
    {
        std::lock_guard<std::mutex> lk(g_named_mu);
        NamedBuffer& b = g_named[name];
        rc = named_prepare(b, m->n);
        if (rc == NZ_OK) {
            cudaError_t e = cudaMemcpyAsync(b.d, m->d, m->n * sizeof(float), cudaMemcpyDeviceToDevice, t_state.stream);
            if (e == cudaSuccess) e = cudaEventRecord(b.last, t_state.stream);
            if (e != cudaSuccess) rc = cuda_fail(e, "nz_context_write");
        } else if (!b.d) {
            g_named.erase(name);
        }
    }
    int32_t rc2 = finish(m);
    return rc != NZ_OK ? rc : rc2;
}

NZ_API int32_t nz_context_read(const char* name, nz_slice_f32 dst) {
    NZ_REQUIRE(name && name[0], "nz_context_read: empty buffer name");
    NZ_REQUIRE(dst.ptr && dst.length > 0 && dst.stride_bytes >= 4, "nz_context_read: empty slice");
    UnscopedCleanup cleanup;
    int32_t rc = begin_stage();
    if (rc != NZ_OK) return rc;
    {
        std::lock_guard<std::mutex> lk(g_named_mu);
        auto it = g_named.find(name);
        if (it == g_named.end()) {
            set_error("nz_context_read: no buffer named '%s'", name);
            return NZ_E_STATE;
        }
        NZ_REQUIRE(it->second.n == (size_t)dst.length, "nz_context_read: buffer '%s' holds %zu floats, the slice %d", name,
                   it->second.n, dst.length);
    }
    Mirror* m = nullptr;
    if ((rc = acquire(dst, /*need_contents=*/false, &m)) != NZ_OK) return rc;
    if ((rc = mark_uploaded()) != NZ_OK) return rc;
    {
        std::lock_guard<std::mutex> lk(g_named_mu);
        auto it = g_named.find(name);
        if (it == g_named.end() || it->second.n != m->n) {
            set_error("nz_context_read: buffer '%s' changed during the call", name);
            rc = NZ_E_STATE;
        } else {
            NamedBuffer& b = it->second;
            cudaError_t e = cudaStreamWaitEvent(t_state.stream, b.last, 0);
            if (e == cudaSuccess) e = cudaMemcpyAsync(m->d, b.d, m->n * sizeof(float), cudaMemcpyDeviceToDevice, t_state.stream);
            if (e == cudaSuccess) e = cudaEventRecord(b.last, t_state.stream);
            if (e != cudaSuccess) rc = cuda_fail(e, "nz_context_read");
            else m->dirty = true;
        }
    }
    int32_t rc2 = finish(m);
    return rc != NZ_OK ? rc : rc2;
}

NZ_API int32_t nz_context_exists(const char* name, int32_t* length) {
    NZ_REQUIRE(name && length, "nz_context_exists: null argument");
    std::lock_guard<std::mutex> lk(g_named_mu);
    auto it = g_named.find(name);
    *length = it == g_named.end() ? -1 : (int32_t)it->second.n;
    return NZ_OK;
}

NZ_API int32_t nz_context_download(const char* name, float* h_dst, int32_t length) {
    NZ_REQUIRE(name && h_dst && length > 0, "nz_context_download: bad arguments");
    int32_t rc = thread_ready();
    if (rc != NZ_OK) return rc;
    std::lock_guard<std::mutex> lk(g_named_mu);
    auto it = g_named.find(name);
    if (it == g_named.end()) {
        set_error("nz_context_download: no buffer named '%s'", name);
        return NZ_E_STATE;
    }
    NZ_REQUIRE(it->second.n == (size_t)length, "nz_context_download: buffer '%s' holds %zu floats, not %d", name, it->second.n, length);
    NZ_CUDA(cudaStreamWaitEvent(t_state.stream, it->second.last, 0));
    NZ_CUDA(cudaMemcpyAsync(h_dst, it->second.d, (size_t)length * sizeof(float), cudaMemcpyDeviceToHost, t_state.stream));
    NZ_CUDA(cudaEventRecord(it->second.last, t_state.stream));
    NZ_CUDA(cudaStreamSynchronize(t_state.stream));
    return NZ_OK;
}

NZ_API int32_t nz_context_upload(const char* name, const float* h_src, int32_t length) {
    NZ_REQUIRE(name && name[0] && h_src && length > 0, "nz_context_upload: bad arguments");
    int32_t rc = thread_ready();
    if (rc != NZ_OK) return rc;
    std::lock_guard<std::mutex> lk(g_named_mu);
    NamedBuffer& b = g_named[name];
    rc = named_prepare(b, (size_t)length);
    if (rc != NZ_OK) {
        if (!b.d) g_named.erase(name);
        return rc;
    }
    NZ_CUDA(cudaMemcpyAsync(b.d, h_src, (size_t)length * sizeof(float), cudaMemcpyHostToDevice, t_state.stream));
    NZ_CUDA(cudaEventRecord(b.last, t_state.stream));
    NZ_CUDA(cudaStreamSynchronize(t_state.stream));      // h_src may be pageable and is the caller's to reuse
    return NZ_OK;
}

NZ_API int32_t nz_context_release(const char* name) {
    NZ_REQUIRE(name, "nz_context_release: null name");
    std::lock_guard<std::mutex> lk(g_named_mu);
    auto it = g_named.find(name);
    if (it == g_named.end()) return NZ_OK;
    if (it->second.last) {
        cudaEventSynchronize(it->second.last);
        cudaEventDestroy(it->second.last);
    }
    pool_free(it->second.d);
    g_named.erase(it);
    return NZ_OK;
}

NZ_API int32_t nz_map_range(nz_slice_f32 map, float* res3, float lim_min, float lim_max) {
    NZ_REQUIRE(res3 != nullptr, "nz_map_range: result pointer is null");
    NZ_REQUIRE(map.ptr && map.length > 0 && map.stride_bytes >= 4, "nz_map_range: empty map");
    UnscopedCleanup cleanup;
    int32_t rc = begin_stage();
    if (rc != NZ_OK) return rc;
    Mirror* m = nullptr;
    if ((rc = acquire(map, /*need_contents=*/true, &m)) != NZ_OK) return rc;
    if ((rc = mark_uploaded()) != NZ_OK) return rc;
    void* scratch = nullptr;
    rc = pool_alloc(&scratch, map_range_scratch_bytes() + 3 * sizeof(float));
    cudaError_t e = cudaSuccess;
    if (rc == NZ_OK) {
        float* d_res = (float*)((char*)scratch + map_range_scratch_bytes());
        cudaStream_t s = t_state.stream;
        rc = launch_map_range(m->d, m->n, lim_min, lim_max, d_res, scratch, s);
        if (rc == NZ_OK) {
            e = cudaEventRecord(t_state.ev[2], s);
            if (e == cudaSuccess) e = cudaMemcpyAsync(res3, d_res, 3 * sizeof(float), cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaEventRecord(t_state.ev[3], s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);   // the caller reads res3 on return
        }
    }
    pool_free(scratch);
    if (rc != NZ_OK) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "nz_map_range");
    t_state.timed = true;
    t_state.launches = (int)(g_launches.load() - t_state.launches_at_start);
    return NZ_OK;
}

}  // extern "C"
