"""Per-stage device timing of the C5 chain at a given resolution (device-resident, CUDA events).
    python tools/stage_times.py [N] [reps]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import noize_job_b200 as nz  # noqa: E402


def timeit(fn, reps):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    d = nz.device
    a = torch.empty(N, N, device="cuda")
    b = torch.empty_like(a)
    scratch = torch.empty(max(16, d.flowmap_scratch_bytes(N, N, 5)), dtype=torch.uint8, device="cuda")
    cells = N * N
    rows = []
    for nt in (3, 5, 4, 1, 0, 2, 6, 7):
        ms = timeit(lambda: d.fractal(a, nt, 0.4, octaves=13, noise_size=1700), reps)
        rows.append((f"fbm type {nt} x13", ms, None))
    d.fractal(a, 3, 0.4, octaves=13, noise_size=1700)
    ms = timeit(lambda: d.kernel_filter(a, b, 2, 17), reps)
    rows.append(("gauss5 x17", ms, 8 * cells))
    ms = timeit(lambda: d.kernel_filter(a, b, 3, 3), reps)
    rows.append(("gauss3 x3", ms, 8 * cells))
    ms = timeit(lambda: d.kernel_filter(a, b, 11, 1), reps)
    rows.append(("sobel3_2d", ms, 8 * cells))
    ms = timeit(lambda: d.flowmap(a, b, scratch, 5, 0.0, 0.005), reps)
    rows.append(("flowmap x5", ms, 8 * cells))
    ms = timeit(lambda: d.min_erosion(a, b, 5), reps)
    rows.append(("min erosion x5", ms, 8 * cells))
    # SURVEY section 8f rows (compulsory bytes: 8 B/cell per unary map, 12 B/cell per binary map,
    # 8 B/cell per thermal phase = 32 B/cell per thermal iteration)
    d.fractal(a, 3, 0.4, octaves=13, noise_size=1700)
    d.fractal(b, 1, 0.4, octaves=2, noise_size=1700)
    ms = timeit(lambda: d.thermal_erosion(a, 45.0, 0.5, 0.75, 1), reps)
    rows.append(("thermal x1 (4 phase launches)", ms, 32 * cells))
    ms = timeit(lambda: d.thermal_erosion(a, 45.0, 0.5, 0.75, 1, tmp=b), reps)
    rows.append(("thermal x1 (fused tile)", ms, 8 * cells))
    d.fractal(b, 1, 0.4, octaves=2, noise_size=1700)
    # subtractive-flow erosion, 3 cycles = 1 + 2 + 3 flow iterations on per-iteration kernels: 64 + 144 + 208 B/cell
    # (outflow step 20/36/40 B, water step 20/24 B, fused erosion epilogue 24 B)
    ms = timeit(lambda: d.subtractive_flow_erosion(a, 3, 0.001, 0.0, 0.005), max(2, reps // 4))
    rows.append(("subtractive flow x3", ms, 416 * cells))
    d.fractal(a, 3, 0.4, octaves=13, noise_size=1700)
    ms = timeit(lambda: d.constant(a, 0, 0.999), reps)
    rows.append(("constant mul", ms, 8 * cells))
    ms = timeit(lambda: d.reduce(a, b, 3), reps)
    rows.append(("reduce max", ms, 12 * cells))
    cv = torch.linspace(0, 1, 256, device="cuda") ** 2
    ms = timeit(lambda: d.curve(a, cv), reps)
    rows.append(("curve 256", ms, 8 * cells))
    ms = timeit(lambda: d.map_range(a), reps)
    rows.append(("map range", ms, 4 * cells))
    ms = timeit(lambda: d.normalize(a, 0.1, 0.7), reps)
    rows.append(("normalize", ms, 8 * cells))
    d.fractal(a, 3, 0.4, octaves=13, noise_size=1700)
    R = N - 8
    vtx = torch.empty((R + 1) * (R + 1), 12, device="cuda")
    idx = torch.empty(6 * R * R, dtype=torch.int32, device="cuda")
    ms = timeit(lambda: d.heightmap_mesh(1, vtx, idx, R, N, 4, 2000.0, 1984.375, a), reps)
    rows.append(("mesh", ms, 4 * cells + 48 * (R + 1) ** 2 + 24 * R * R))
    for name, ms, byts in rows:
        extra = f"  {byts / ms / 1e6:9.1f} GB/s (compulsory bytes)" if byts else f"  {1942 * cells / ms / 1e9:9.2f} TFLOP/s (simplex count)"
        print(f"{name:18s} {ms:9.3f} ms  {cells / ms / 1e3:10.1f} Mcells/s{extra}")


if __name__ == "__main__":
    main()
