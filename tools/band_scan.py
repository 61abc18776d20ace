"""Stencil stages on the windows an 8-GPU row band runs (2048 own rows of a 16384-wide grid + ghost rows), across chunk
heights: is the launch-granularity loss at band size a matter of the chunk heuristic?  (profiling aid)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import noize_job_b200 as nz
d = nz.device
W = 16384
def t(fn, reps=6):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
src = torch.empty(2116, W, device="cuda")
d.fractal(src, 3, 0.4, octaves=13, noise_size=1700)
a = src.clone(); b = torch.empty_like(a)
print("full-grid per-band share (t1/8): filter 0.311  flow 0.373  erosion 0.052 ms")
for zc in (None, 64, 86, 103, 137, 171, 206, 257, 342, 513):
    if zc is None: os.environ.pop("NZ_WALK_ZC", None)
    else: os.environ["NZ_WALK_ZC"] = str(zc)
    print(f"filter  2116 rows  NZ_WALK_ZC={zc}: {t(lambda: d.kernel_filter(a, b, 2, 17)):.4f} ms", flush=True)
os.environ.pop("NZ_WALK_ZC", None)
f = d.kernel_filter(src.clone(), b, 2, 17).clone()[:2070].contiguous(); g = torch.empty_like(f)
for grp in ("0", "4"):
    os.environ["NZ_FLOW_GROUP"] = grp
    for zc in (None, 64, 98, 128, 171, 256, 342, 496):
        if zc is None: os.environ.pop("NZ_FLOWWALK_ZC", None)
        else: os.environ["NZ_FLOWWALK_ZC"] = str(zc)
        print(f"flow    2070 rows  group={grp} NZ_FLOWWALK_ZC={zc}: {t(lambda: d.flowmap(f, g, None, 5, 0.0, 0.005)):.4f} ms", flush=True)
os.environ.pop("NZ_FLOWWALK_ZC", None); os.environ.pop("NZ_FLOW_GROUP", None)
e = f[:2054].contiguous(); h = torch.empty_like(e)
print(f"erosion 2054 rows: {t(lambda: d.min_erosion(e, h, 5)):.4f} ms")
n = torch.empty(2048, W, device="cuda")
print(f"noise   2048 rows: {t(lambda: d.fractal(n, 3, 0.4, octaves=13, noise_size=1700, z_first=4096)):.4f} ms  (t1/8 = 0.932)")
