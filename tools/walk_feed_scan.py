"""Gauss5 x17 at 16384^2: per-lane cp.async row feed vs the bulk-copy (TMA engine) feed, best of 6."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import noize_job_b200 as nz
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
a = torch.rand(N, N, device="cuda"); b = torch.empty_like(a)
def t(fn, reps=6):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
for feed in ("cp.async", "bulk", "cp.async", "bulk"):
    if feed == "bulk": os.environ["NZ_WALK_FEED"] = "bulk"
    else: os.environ.pop("NZ_WALK_FEED", None)
    print(f"gauss5 x17 {N}^2, row feed {feed:8s}: {t(lambda: nz.device.kernel_filter(a, b, 2, 17)):.3f} ms   gauss3 x3: {t(lambda: nz.device.kernel_filter(a, b, 3, 3)):.3f} ms", flush=True)
