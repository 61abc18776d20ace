"""Why does the end-to-end (download) leg not scale with GPUs?  D2H probe for 1..N concurrent ranks.

    torchrun --nproc-per-node N tools/d2h_probe_multi.py [GiB per rank = 4]

Every rank downloads the same amount from its GPU into pinned host memory, (a) one rank at a time, (b) k = 1, 2, 4, .. N ranks
at the same time, for each kind of host buffer:
    torch-pinned          torch.empty(pin_memory=True)                 (what bench.py used in round 1)
    hostalloc             cudaHostAlloc(default)
    hostalloc-wc          cudaHostAlloc(write-combined)
    registered            malloc'ed, first-touched by this rank, then cudaHostRegister (what nz_pin does to Unity's buffers)
and with the process bound to the GPU's NUMA node (bands.bind_host_to_gpu_numa_node) or left unbound (NZ_PROBE_NO_BIND=1).
Rank 0 prints one table: GB/s per rank alone, aggregate GB/s at each concurrency.
"""
import ctypes as C
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from noize_job_b200 import bands  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
nbytes = int(gib * 2 ** 30)
torch.cuda.set_device(local)
bound = "" if os.environ.get("NZ_PROBE_NO_BIND") else bands.bind_host_to_gpu_numa_node(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def _loaded_cudart():
    """The libcudart torch has already loaded (one runtime instance in the process)."""
    for line in open("/proc/self/maps"):
        if "libcudart" in line:
            return C.CDLL(line.split()[-1])
    return C.CDLL("/usr/local/cuda/lib64/libcudart.so.12")


cudart = _loaded_cudart()
libc = C.CDLL("libc.so.6")
libc.malloc.restype = C.c_void_p
libc.malloc.argtypes = [C.c_size_t]
libc.memset.argtypes = [C.c_void_p, C.c_int, C.c_size_t]
libc.free.argtypes = [C.c_void_p]

src = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
src.fill_(1)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def host_buffer(kind):
    """(pointer, release()) of nbytes of page-locked host memory of the given kind."""
    if kind == "torch-pinned":
        t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        return t.data_ptr(), (lambda: None), t
    p = C.c_void_p()
    if kind.startswith("hostalloc"):
        flags = 0x04 if kind.endswith("wc") else 0x00          # cudaHostAllocWriteCombined / cudaHostAllocDefault
        rc = cudart.cudaHostAlloc(C.byref(p), C.c_size_t(nbytes), C.c_uint(flags))
        assert rc == 0, f"cudaHostAlloc failed: {rc}"
        return p.value, (lambda: cudart.cudaFreeHost(p)), None
    if kind == "registered":
        q = libc.malloc(nbytes)
        libc.memset(q, 0, nbytes)                              # first touch on this rank's (bound) NUMA node
        rc = cudart.cudaHostRegister(C.c_void_p(q), C.c_size_t(nbytes), C.c_uint(0))
        assert rc == 0, f"cudaHostRegister failed: {rc}"

        def rel():
            cudart.cudaHostUnregister(C.c_void_p(q))
            libc.free(q)
        return q, rel, None
    raise ValueError(kind)


def copy_time(ptr, active):
    """Seconds for this rank's D2H when `active` (else it idles); all ranks pass the barriers."""
    stream = torch.cuda.current_stream().cuda_stream
    barrier()
    t0 = time.perf_counter()
    if active:
        rc = cudart.cudaMemcpyAsync(C.c_void_p(ptr), C.c_void_p(src.data_ptr()), C.c_size_t(nbytes), C.c_int(2), C.c_void_p(stream))
        assert rc == 0, rc
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    return dt


def gather(v):
    t = torch.tensor([v], device="cuda", dtype=torch.float64)
    if world == 1:
        return [v]
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(x) for x in out]


levels = [k for k in (1, 2, 4, 8, 16) if k <= world]
if world not in levels:
    levels.append(world)
if rank == 0:
    print(f"# D2H probe: {world} rank(s), {gib:g} GiB per rank; binding: {bound or 'none'}")
    print(f"{'buffer':14s} {'alone GB/s (per rank)':40s} " + " ".join(f"{'x' + str(k) + ' aggregate':>14s}" for k in levels))
for kind in ("torch-pinned", "hostalloc", "hostalloc-wc", "registered"):
    ptr, release, keep = host_buffer(kind)
    copy_time(ptr, True)                                       # warm-up: page tables, first DMA
    alone = []
    for r in range(world):
        dt = copy_time(ptr, rank == r)
        alone.append(gather(nbytes / dt / 1e9 if rank == r else 0.0)[r])
    agg = []
    for k in levels:
        dt = copy_time(ptr, rank < k)
        times = gather(dt if rank < k else 0.0)
        agg.append(k * nbytes / max(times) / 1e9)
    if rank == 0:
        print(f"{kind:14s} {' '.join(f'{a:5.1f}' for a in alone):40s} " + " ".join(f"{a:14.1f}" for a in agg))
    release()
    del keep
if world > 1:
    dist.destroy_process_group()
