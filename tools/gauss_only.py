"""Runs Gauss5 x ITER on a ROWS x 16384 window (for ncu captures of the register-walk filter kernel)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import noize_job_b200 as nz
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
it = int(sys.argv[2]) if len(sys.argv) > 2 else 4
a = torch.rand(rows, 16384, device="cuda")
b = torch.empty_like(a)
for _ in range(2):
    nz.device.kernel_filter(a, b, 2, it)
torch.cuda.synchronize()
print("ok")
