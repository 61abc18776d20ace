"""Gauss5 x17 at 16384^2 and on an 8-GPU band window: 4 stages per launch (4,4,3,3,3) vs 5 (5,4,4,4), best of 6."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import noize_job_b200 as nz
def t(fn, reps=6):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
for rows in (16384, 2116):
    a = torch.rand(rows, 16384, device="cuda"); b = torch.empty_like(a)
    for tmax in ("4", "5", "4", "5"):
        os.environ["NZ_WALK_TMAX"] = tmax
        print(f"rows {rows:5d}  gauss5 x17, up to {tmax} stages per launch: {t(lambda: nz.device.kernel_filter(a, b, 2, 17)):.3f} ms   x15: {t(lambda: nz.device.kernel_filter(a, b, 2, 15)):.3f} ms   x20: {t(lambda: nz.device.kernel_filter(a, b, 2, 20)):.3f} ms", flush=True)
