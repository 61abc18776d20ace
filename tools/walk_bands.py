"""Gauss5 x17 on row bands of the heights a 1/2/4/8-GPU split produces (profiling aid for the chunk heuristic)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import noize_job_b200 as nz
W = 16384
for rows in [int(x) for x in (sys.argv[1:] or [16384, 8226, 4164, 2116, 2082, 1024, 3000])]:
    a = torch.rand(rows, W, device="cuda"); b = torch.empty_like(a)
    best = 1e9
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); nz.device.kernel_filter(a, b, 2, 17); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"rows={rows:6d} gauss5 x17: {best:.3f} ms  {rows * W / best / 1e3:.0f} Mcells/s  ZC={os.environ.get('NZ_WALK_ZC')}")
