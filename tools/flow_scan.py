import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import noize_job_b200 as nz
N = 16384
a = torch.empty(N, N, device="cuda"); b = torch.empty_like(a)
nz.device.fractal(a, 3, 0.4, octaves=13, noise_size=1700)
a = nz.device.kernel_filter(a, b, 2, 17).clone()
for it in (1, 2, 3, 4, 5):
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); nz.device.flowmap(a, b, None, it, 0.0, 0.005); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"flow iterations={it}: {best:.3f} ms")
