"""Device-resident time of every BASELINE.json config on one GPU (CUDA events, best of `reps`), device layer of the C ABI.
    python tools/config_times.py [reps]
C4 is measured on the 32 tiles one GPU of eight owns (tiles are independent: no communication)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import noize_job_b200 as nz  # noqa: E402

d = nz.device
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5


def timeit(fn):
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


def bufs(n):
    return torch.empty(n, n, device="cuda"), torch.empty(n, n, device="cuda")


def mesh_bufs(n):
    R = n - 8
    return R, torch.empty((R + 1) * (R + 1), 12, device="cuda"), torch.empty(6 * R * R, dtype=torch.int32, device="cuda")


def chain(a, b, noise_type, n, xpos=0, zpos=0, flow=True, erosion=True, mesh=None):
    d.fractal(a, noise_type, 0.4, octaves=13, xpos=xpos, zpos=zpos, noise_size=1700)
    cur = d.kernel_filter(a, b, 2, 17)
    other = b if cur is a else a
    if flow:
        r = d.flowmap(cur, other, None, 5, 0.0, 0.005)
        if r is not cur:
            cur, other = other, cur
    if erosion:
        r = d.min_erosion(cur, other, 5)
        if r is not cur:
            cur, other = other, cur
    if mesh:
        R, v, i = mesh
        d.heightmap_mesh(1, v, i, R, n, 4, 2000.0, R * (500.0 / 256.0), cur)


rows = []
a, b = bufs(256)
rows.append(("C1  256^2 simplex fBm x13", timeit(lambda: d.fractal(a, 3, 0.4, octaves=13, noise_size=1700)), 256 * 256))
a, b = bufs(1024)
m = mesh_bufs(1024)
rows.append(("C2  1024^2 simplex -> Gauss5 x17 -> flow x5 -> erosion x5 -> mesh", timeit(lambda: chain(a, b, 3, 1024, mesh=m)), 1024 * 1024))
a, b = bufs(4096)
rows.append(("C3  4096^2 cellular -> Gauss5 x17 -> flow x5", timeit(lambda: chain(a, b, 5, 4096, erosion=False)), 4096 * 4096))
a, b = bufs(1024)
c = torch.empty_like(a)
m = mesh_bufs(1024)


def c4_tiles(ntiles=32):
    for t in range(ntiles):
        tx, tz = t % 16, t // 16
        d.fractal(a, 4, 0.4, octaves=13, xpos=1000 * tx, zpos=1000 * tz, noise_size=1700)
        cur = d.kernel_filter(a, b, 3, 3)
        c.copy_(cur)
        d.kernel_filter(c, b if cur is a else a, 11, 1)           # Sobel3_2D on a copy
        d.heightmap_mesh(1, m[1], m[2], m[0], 1024, 4, 2000.0, m[0] * (500.0 / 256.0), cur)


rows.append(("C4  32 tiles x 1024^2 (one GPU's share of 16x16): rotated simplex -> Gauss3 x3 -> Sobel3_2D -> mesh", timeit(c4_tiles), 32 * 1024 * 1024))

# the same 32 tiles on 4 streams (what 4 host worker threads of the stage API do: one stream per thread)
NS = 4
streams = [torch.cuda.Stream() for _ in range(NS)]
sb = [(torch.empty(1024, 1024, device="cuda"), torch.empty(1024, 1024, device="cuda"), torch.empty(1024, 1024, device="cuda"), mesh_bufs(1024))
      for _ in range(NS)]


def c4_tiles_streams(ntiles=32):
    cur_stream = torch.cuda.current_stream()
    for s in streams:
        s.wait_stream(cur_stream)
    for t in range(ntiles):
        tx, tz = t % 16, t // 16
        s = streams[t % NS]
        ta, tb, tc, tm = sb[t % NS]
        with torch.cuda.stream(s):
            d.fractal(ta, 4, 0.4, octaves=13, xpos=1000 * tx, zpos=1000 * tz, noise_size=1700, stream=s)
            cur = d.kernel_filter(ta, tb, 3, 3, stream=s)
            tc.copy_(cur)
            d.kernel_filter(tc, tb if cur is ta else ta, 11, 1, stream=s)
            d.heightmap_mesh(1, tm[1], tm[2], tm[0], 1024, 4, 2000.0, tm[0] * (500.0 / 256.0), cur, stream=s)
    for s in streams:
        cur_stream.wait_stream(s)


rows.append(("C4  same 32 tiles on 4 concurrent streams", timeit(c4_tiles_streams), 32 * 1024 * 1024))
for name, ms, cells in rows:
    print(f"{name:100s} {ms:9.3f} ms  {cells / ms / 1e3:10.1f} Mcells/s")
