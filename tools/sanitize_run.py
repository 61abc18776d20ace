"""Small invocation of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool racecheck python tools/sanitize_run.py
Sizes are chosen to hit strip / chunk / tile seams and both border and interior code paths while staying sanitizer-sized."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import noize_job_b200 as nz  # noqa: E402

d = nz.device
n = 768
a = torch.empty(n, n, device="cuda"); b = torch.empty_like(a)
for nt in range(8):
    d.fractal(a, nt, 0.4, octaves=3, noise_size=170)
os.environ["NZ_FBM_PATH"] = "pair"
for nt in (3, 5, 1):
    d.fractal(a, nt, 0.4, octaves=3, noise_size=170)
os.environ.pop("NZ_FBM_PATH")
for path in ("walk", "fused", "generic"):
    os.environ["NZ_SEP_PATH"] = path
    d.kernel_filter(a, b, 2, 5)
os.environ.pop("NZ_SEP_PATH")
d.kernel_filter(a, b, 11, 1)
for path, it in (("reg", 5), ("reg", 3), ("reg", 1), ("wave", 5), ("tile", 5), ("tile", 2), ("wave", 1)):
    os.environ["NZ_FLOW_PATH"] = path
    d.flowmap(a.clone(), b, None, it, 0.0, 0.005)
os.environ.pop("NZ_FLOW_PATH")
os.environ["NZ_FLOW_UNFUSED"] = "1"
d.flowmap(a.clone(), b, torch.empty(5 * n * n * 4, dtype=torch.uint8, device="cuda"), 2, 0.0, 0.005)
os.environ.pop("NZ_FLOW_UNFUSED")
d.subtractive_flow_erosion(a.clone(), 2, 0.1, 0.0, 0.005)
d.flowmap(a.clone() * 3e37, b, None, 5, 0.0, 0.005)   # leaves the register kernel's guard: rerun on the wavefront kernel
d.min_erosion(a, b, 5)
d.thermal_erosion(a, 45.0, 0.5, 0.75, 2)
d.constant(a, 0, 0.5); d.reduce(a, b, 3); d.curve(a, torch.linspace(0, 1, 64, device="cuda")); d.normalize(a, 0.1, 0.5)
print(d.map_range(a).cpu().numpy())
c = torch.empty(300, 300, device="cuda"); d.crop(a, c, 100)
R = n - 8
v = torch.empty((R + 1) * (R + 1), 12, device="cuda"); i = torch.empty(6 * R * R, dtype=torch.int32, device="cuda")
for mt in (0, 1):
    d.heightmap_mesh(mt, v, i, R, n, 4, 2000.0, 1500.0, b)
# a grid above 4M cells takes the interior launches of the register-walk flow map (group strips at I = 5, independent strips
# at I = 3) with the border launch beside them; the merged filter launch (border items + interior CTAs) on the same grid
w2, h2 = 2304, 2048
a2 = torch.rand(h2, w2, device="cuda") * 0.05; b2 = torch.empty_like(a2)
d.kernel_filter(a2, b2, 2, 7)          # Gauss5: T = 4 + 3
d.kernel_filter(a2, b2, 5, 2)          # a filter with a scale factor
d.flowmap(a2.clone(), b2, None, 5, 0.0, 0.005)
d.flowmap(a2.clone(), b2, None, 3, 0.0, 0.005)
# row bands inside the library: two bands on this GPU, one exchange per pass, and the same chain in recompute mode
from noize_job_b200 import bands
cfg = bands.ChainConfig(N=512, noise_size=170, filter_iterations=6, flow_iterations=3, erosion_iterations=4)
for mode in ("exchange", "recompute"):
    ch = bands.LibBandChain(cfg, devices=[0, 0], mode=mode)
    ch.run(); ch.sync() if hasattr(ch, "sync") else None
    ch.release()
# deferred element-wise run through the host layer
h = np.random.rand(256 * 256).astype(np.float32)
with nz.host.pipeline():
    nz.host.constant(h, None, 0, 0.9, 256)
    nz.host.normalize(h, None, np.array([0.0, 1.0, 1.0], np.float32), 256)
    nz.host.curve(h, None, np.linspace(0, 1, 32).astype(np.float32), 256)
    nz.host.constant(h, None, 1, 0.5, 256)
torch.cuda.synchronize()
print("sanitize_run ok")
