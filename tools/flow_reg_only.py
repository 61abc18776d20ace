"""Runs only the register-walk flow map on filtered noise at N (for ncu captures)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import noize_job_b200 as nz
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
os.environ["NZ_FLOW_PATH"] = sys.argv[2] if len(sys.argv) > 2 else "reg"
h = torch.empty(N, N, device="cuda"); t = torch.empty_like(h)
nz.device.fractal(h, 3, 0.4, octaves=13, noise_size=1700)
h = nz.device.kernel_filter(h, t, 2, 17).clone()
for _ in range(2):
    nz.device.flowmap(h, t, None, 5, 0.0, 0.005)
torch.cuda.synchronize()
print("ok")
