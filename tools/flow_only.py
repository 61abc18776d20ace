"""Runs only the fused flow map at N (for ncu captures)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import noize_job_b200 as nz
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
what = sys.argv[2] if len(sys.argv) > 2 else "flow"
a = torch.rand(N, N, device="cuda") * 0.05
b = torch.empty_like(a)
for _ in range(2):
    if what == "flow":
        nz.device.flowmap(a, b, None, 5, 0.0, 0.005)
    elif what == "gauss":
        nz.device.kernel_filter(a, b, 2, 17)
    elif what == "noise":
        nz.device.fractal(a, 3, 0.4, octaves=13, noise_size=1700)
    elif what == "psr":
        nz.device.fractal(a, 4, 0.4, octaves=13, noise_size=1700)
    elif what == "erosion":
        nz.device.min_erosion(a, b, 5)
torch.cuda.synchronize()
print("ok")
