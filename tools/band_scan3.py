"""Fixed cost of a band-sized filter / flow launch: Gauss5 x4 (ONE sep_walk launch, T = 4) and flow x5 over windows of
increasing height, window edges not grid edges (NZ_GRID_EDGES=0).  time(rows) = a + b * rows: `a` is what 8 bands pay 5x."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import noize_job_b200 as nz
d = nz.device
W = 16384
def t(fn, reps=10):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
print({k: v for k, v in os.environ.items() if k.startswith("NZ_")})
what = sys.argv[1] if len(sys.argv) > 1 else "both"
full = torch.empty(16384, W, device="cuda"); tmp = torch.empty_like(full)
d.fractal(full, 3, 0.4, octaves=13, noise_size=1700)         # the bench chain's input to the flow map: filtered simplex fBm
full = d.kernel_filter(full, tmp, 2, 17).clone()
del tmp
ROWS = [int(x) for x in os.environ["ROWS"].split(",")] if os.environ.get("ROWS") else (264, 529, 1058, 1587, 2116, 2645, 3174, 4232, 8464, 16384)
for rows in ROWS:
    a = full[:rows].clone(); b = torch.empty_like(a)
    line = f"rows {rows:5d}"
    if what in ("both", "filter"):
        t4 = t(lambda: d.kernel_filter(a, b, 2, 4))
        t8 = t(lambda: d.kernel_filter(a, b, 2, 8))
        t3 = t(lambda: d.kernel_filter(a, b, 2, 3))
        t17 = t(lambda: d.kernel_filter(a, b, 2, 17))
        line += f"  gauss5 x4 {t4*1e3:7.1f} us ({t4*1e6/rows:6.2f} ns/row)  x8 {t8*1e3:7.1f}  x3 {t3*1e3:7.1f}  x17 {t17*1e3:7.1f} us"
    if what in ("both", "flow"):
        tf = t(lambda: d.flowmap(a, b, None, 5, 0.0, 0.005))
        line += f"  flow x5 {tf*1e3:7.1f} us ({tf*1e6/rows:6.2f} ns/row)"
    print(line, flush=True)
