// aux_stream.cu — a side stream per (host thread, device) for launches that are independent of the next launch on the
// caller's stream.  The register-walk kernels (flowwalk_kernels.cu, sepwalk_kernels.cu) run their border warps as a
// separate, short launch; enqueued on the caller's stream it would serialise with the interior launch (one chunk walk of
// latency, ~35-90 us, per launch); forked onto the side stream it runs underneath it.
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
thread_local Aux t_aux[64];

int32_t get_aux(Aux** out) {
    int dev = 0;
    NZ_CUDA(cudaGetDevice(&dev));
    NZ_REQUIRE(dev >= 0 && dev < 64, "aux stream: device ordinal %d out of range", dev);
    Aux& a = t_aux[dev];
    if (!a.s) {
        // (a high-priority side stream was tried, to make the border CTAs win SM slots from the interior launch: no
        // measurable change, tools/band_scan3.py)
        NZ_CUDA(cudaStreamCreateWithFlags(&a.s, cudaStreamNonBlocking));
        NZ_CUDA(cudaEventCreateWithFlags(&a.fork, cudaEventDisableTiming));
        NZ_CUDA(cudaEventCreateWithFlags(&a.join, cudaEventDisableTiming));
    }
    *out = &a;
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
    int32_t rc = get_aux(&a);
    if (rc != NZ_OK) return rc;
    NZ_CUDA(cudaEventRecord(a->fork, main));
    NZ_CUDA(cudaStreamWaitEvent(a->s, a->fork, 0));
    *aux = a->s;
    return NZ_OK;
}

int32_t aux_join(cudaStream_t main) {
    if (aux_disabled()) return NZ_OK;
    Aux* a;
    int32_t rc = get_aux(&a);
    if (rc != NZ_OK) return rc;
    NZ_CUDA(cudaEventRecord(a->join, a->s));
    NZ_CUDA(cudaStreamWaitEvent(main, a->join, 0));
    return NZ_OK;
}

}  // namespace nz
