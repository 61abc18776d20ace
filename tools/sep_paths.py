"""Gauss5 x17 / Gauss3 x3: walk vs fused (shared-memory tile) separable paths across grid sizes (dispatch heuristic aid)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import noize_job_b200 as nz
for n in (256, 512, 1024, 2048, 4096):
    a = torch.rand(n, n, device="cuda"); b = torch.empty_like(a)
    for ft, it in ((2, 17), (3, 3)):
        res = {}
        for path in ("walk", "fused"):
            if path == "fused": os.environ["NZ_SEP_PATH"] = "fused"
            else: os.environ.pop("NZ_SEP_PATH", None)
            best = 1e9
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); nz.device.kernel_filter(a, b, ft, it); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            res[path] = best
        print(f"n={n:5d} filter={ft} x{it}: walk {res['walk']:.3f} ms  fused {res['fused']:.3f} ms")
os.environ.pop("NZ_SEP_PATH", None)
