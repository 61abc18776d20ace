// aux_stream.cu — a side stream per (host thread, device, caller stream) for a launch that is independent of the next launch on
// the caller's stream.  The register-walk flow map (flowwalk_kernels.cu) runs its border warps as a separate, short launch
// beside the interior launch (the register-walk filter did too until its border items moved into the interior launch; its
// two-launch form is still selectable).
//   aux_fork(main, &aux): everything enqueued on `main` so far happens before what is enqueued on `aux` next
//   aux_join(main)      : everything enqueued on `aux` so far happens before what is enqueued on `main` next
// Both are event record/wait pairs (legal inside stream capture).  NZ_NO_AUX_STREAM=1 makes aux == main.
#include <stdlib.h>
#include "nz_common.cuh"

namespace nz {
namespace {

struct Aux {
    cudaStream_t s = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
// One side stream per (host thread, device, caller stream): a thread that drives several streams (the tile world's slots) gets
// one side stream for each, so their border launches neither serialise on one stream nor make one slot's join wait for
// another slot's work.  At most AUX_MAX live entries per thread: beyond that the oldest is recycled (its stream is idle by
// then or will simply order a little more than necessary — fork and join on the same entry are always a matched pair).
constexpr int AUX_MAX = 16;
struct AuxEntry {
    int dev = -1;
    cudaStream_t main = nullptr;
    Aux a;
};
struct AuxTable {
    AuxEntry e[AUX_MAX];
    int next = 0;
    // worker threads come and go under a job system: their side streams go with them
    ~AuxTable() {
        int prev = 0;
        const bool have = cudaGetDevice(&prev) == cudaSuccess;
        for (AuxEntry& c : e)
            if (c.a.s && cudaSetDevice(c.dev) == cudaSuccess) {
                cudaStreamSynchronize(c.a.s);
                cudaStreamDestroy(c.a.s);
                cudaEventDestroy(c.a.fork);
                cudaEventDestroy(c.a.join);
                c.a = Aux{};
            }
        if (have) cudaSetDevice(prev);
        cudaGetLastError();
    }
};
thread_local AuxTable t_auxtab;
#define t_aux t_auxtab.e
#define t_aux_next t_auxtab.next

int32_t get_aux(cudaStream_t main, Aux** out) {
    int dev = 0;
    NZ_CUDA(cudaGetDevice(&dev));
    for (AuxEntry& e : t_aux)
        if (e.a.s && e.dev == dev && e.main == main) {
            *out = &e.a;
            return NZ_OK;
        }
    AuxEntry* e = nullptr;
    for (AuxEntry& c : t_aux)
        if (!c.a.s) { e = &c; break; }
    if (!e) {
        // recycle: prefer an entry of this device (its stream and events can be reused as they are)
        for (int k = 0; k < AUX_MAX && !e; k++) {
            AuxEntry& c = t_aux[(t_aux_next + k) % AUX_MAX];
            if (c.dev == dev) { e = &c; t_aux_next = (t_aux_next + k + 1) % AUX_MAX; }
        }
        if (!e) {
            e = &t_aux[t_aux_next];
            t_aux_next = (t_aux_next + 1) % AUX_MAX;
            int prev = dev;
            cudaSetDevice(e->dev);
            cudaStreamSynchronize(e->a.s);
            cudaStreamDestroy(e->a.s);
            cudaEventDestroy(e->a.fork);
            cudaEventDestroy(e->a.join);
            cudaSetDevice(prev);
            e->a = Aux{};
        }
    }
    if (!e->a.s) {
        // (a high-priority side stream was tried, to make the border CTAs win SM slots from the interior launch: no
        // measurable change, tools/band_scan3.py)
        NZ_CUDA(cudaStreamCreateWithFlags(&e->a.s, cudaStreamNonBlocking));
        NZ_CUDA(cudaEventCreateWithFlags(&e->a.fork, cudaEventDisableTiming));
        NZ_CUDA(cudaEventCreateWithFlags(&e->a.join, cudaEventDisableTiming));
    }
    e->dev = dev;
    e->main = main;
    *out = &e->a;
    return NZ_OK;
}

bool aux_disabled() {
    static const bool off = [] {
        const char* e = getenv("NZ_NO_AUX_STREAM");
        return e && e[0] == '1';
    }();
    return off;
}

}  // namespace

static thread_local int t_grid_edges = 3;
int grid_edges() {
    // NZ_GRID_EDGES=0..3 overrides the hint while profiling band-sized windows on one GPU (tools/band_scan.py)
    static const int forced = [] {
        const char* e = getenv("NZ_GRID_EDGES");
        return e ? atoi(e) & 3 : -1;
    }();
    return forced >= 0 ? forced : t_grid_edges;
}
GridEdgesScope::GridEdgesScope(int edges) : prev(t_grid_edges) { t_grid_edges = edges; }
GridEdgesScope::~GridEdgesScope() { t_grid_edges = prev; }

int sm_count() {
    static std::atomic<int> cache[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int n = cache[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cache[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

int32_t aux_fork(cudaStream_t main, cudaStream_t* aux) {
    if (aux_disabled()) {
        *aux = main;
        return NZ_OK;
    }
    Aux* a;
    int32_t rc = get_aux(main, &a);
    if (rc != NZ_OK) return rc;
    NZ_CUDA(cudaEventRecord(a->fork, main));
    NZ_CUDA(cudaStreamWaitEvent(a->s, a->fork, 0));
    *aux = a->s;
    return NZ_OK;
}

int32_t aux_join(cudaStream_t main) {
    if (aux_disabled()) return NZ_OK;
    Aux* a;
    int32_t rc = get_aux(main, &a);
    if (rc != NZ_OK) return rc;
    NZ_CUDA(cudaEventRecord(a->join, a->s));
    NZ_CUDA(cudaStreamWaitEvent(main, a->join, 0));
    return NZ_OK;
}

}  // namespace nz
