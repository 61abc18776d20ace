"""Gauss5 x17 (walk kernel) timing across grid widths / prefetch depths (profiling aid)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import noize_job_b200 as nz
H = 16384
for W in [int(x) for x in (sys.argv[1:] or ["16384", "16512", "12288"])]:
    a = torch.rand(H, W, device="cuda"); b = torch.empty_like(a)
    best = 1e9
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); nz.device.kernel_filter(a, b, 2, 17); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"W={W} gauss5 x17: {best:.3f} ms  {H * W / best / 1e3:.0f} Mcells/s  env PFR={os.environ.get('NZ_WALK_PFR')} ZC={os.environ.get('NZ_WALK_ZC')}")
