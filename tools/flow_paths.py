"""Flow map: the fused formulations against each other (bitwise) and their times at N^2.
usage: python tools/flow_paths.py [N] [paths...]   paths from wave, tile, reg"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import noize_job_b200 as nz

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
paths = sys.argv[2:] or ["wave", "reg"]


def run(h, path, iters):
    os.environ["NZ_FLOW_PATH"] = path
    return nz.device.flowmap(h, torch.empty_like(h), None, iters, 0.0, 0.005)


torch.manual_seed(1)
for rows, width, iters in [(700, 600, 5), (300, 1000, 4), (97, 236, 3), (40, 20, 2), (513, 472, 1), (33, 4, 5), (1100, 1304, 5), (64, 64, 5), (2048, 2048, 5)]:
    base = torch.rand(rows, width, device="cuda")
    a = torch.empty_like(base)
    h = nz.device.kernel_filter(base, a, 3, 2).clone() * 0.05
    ref = run(h.clone(), "wave", iters).clone()
    for p in paths:
        if p == "wave":
            continue
        got = run(h.clone(), p, iters).clone()
        torch.cuda.synchronize()
        bad = (got != ref) & ~(got.isnan() & ref.isnan())
        print(f"{rows}x{width} I={iters} {p}: {'bit-identical' if not bad.any() else f'{int(bad.sum())} cells differ, first at {bad.nonzero()[0].tolist()}, max|d| {float((got-ref).abs().max()):.3g}'}")

# terrain-like input: filtered noise, as in the bench chain
h = torch.empty(N, N, device="cuda"); t = torch.empty_like(h)
nz.device.fractal(h, 3, 0.4, octaves=13, noise_size=1700)
h = nz.device.kernel_filter(h, t, 2, 17).clone()
for iters in ([5] if len(sys.argv) > 2 and sys.argv[-1] != "all" else [1, 2, 3, 4, 5]):
    for p in paths:
        for _ in range(2):
            run(h, p, iters)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            run(h, p, iters)
        e1.record(); torch.cuda.synchronize()
        print(f"N={N} I={iters} {p}: {e0.elapsed_time(e1)/5:.3f} ms")
