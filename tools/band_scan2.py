"""Where does a band-sized filter launch lose time?  Gauss5 x {4, 8, 12, 16, 17} on a 2116 x 16384 window (marginal cost per
launch), with the window's edges treated as grid edges or not (NZ_GRID_EDGES is read once per process: run twice)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import noize_job_b200 as nz
d = nz.device
W = 16384
def t(fn, reps=8):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
print("NZ_GRID_EDGES =", os.environ.get("NZ_GRID_EDGES"), " NZ_NO_AUX_STREAM =", os.environ.get("NZ_NO_AUX_STREAM"))
for rows in (2116, 16384):
    a = torch.rand(rows, W, device="cuda"); b = torch.empty_like(a)
    for it in (3, 4, 8, 12, 16, 17):
        print(f"rows {rows:5d}  gauss5 x{it:2d}: {t(lambda: d.kernel_filter(a, b, 2, it)):.4f} ms", flush=True)
    f = a[: rows - 46].contiguous(); g = torch.empty_like(f)
    print(f"rows {rows - 46:5d}  flow x5: {t(lambda: d.flowmap(f, g, None, 5, 0.0, 0.005)):.4f} ms", flush=True)
