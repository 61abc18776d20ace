"""PCIe D2H probe: one stream vs two concurrent streams, pinned host memory (sizes of the C5 mesh download)."""
import time, torch
v = torch.empty(12_870_000_000 // 4, dtype=torch.float32, device="cuda")
i = torch.empty(6_440_000_000 // 4, dtype=torch.int32, device="cuda")
hv = torch.empty_like(v, device="cpu", pin_memory=True)
hi = torch.empty_like(i, device="cpu", pin_memory=True)
def run(two):
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.cuda.stream(s1): hv.copy_(v, non_blocking=True)
    with torch.cuda.stream(s2 if two else s1): hi.copy_(i, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return (v.numel() + i.numel()) * 4 / dt / 1e9
for two in (False, True, False, True):
    print("two streams" if two else "one stream ", f"{run(two):.1f} GB/s")
# chunked: 8 chunks alternating
