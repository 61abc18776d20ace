"""Flow map x5 at 16384^2: independent strips vs group strips (NZ_FLOW_GROUP = 0 / 4 / 6), best of 5, plus I = 3, 4."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import noize_job_b200 as nz
d = nz.device
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
a = torch.empty(N, N, device="cuda"); b = torch.empty_like(a)
d.fractal(a, 3, 0.4, octaves=13, noise_size=1700)
src = d.kernel_filter(a, b, 2, 17).clone()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
for I in (5, 4, 3):
    for g in ("0", "4", "6"):
        os.environ["NZ_FLOW_GROUP"] = g
        a.copy_(src)
        print(f"I={I} group={g}: {t(lambda: d.flowmap(a, b, None, I, 0.0, 0.005)):.3f} ms", flush=True)
