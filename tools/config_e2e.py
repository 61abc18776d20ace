"""BASELINE config C2 (README example #1 on one 1024^2 tile) END TO END through the stage API: host buffers in and out,
the way a Unity worker thread would drive it (one residency scope per tile: heightmap D2H 4 MB + mesh D2H 74 MB)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import noize_job_b200 as nz

res, R = 1024, 1016
data = torch.empty(res * res, dtype=torch.float32, pin_memory=True).numpy()
vtx = torch.empty((R + 1) * (R + 1), 12, dtype=torch.float32, pin_memory=True).numpy()
idx = torch.empty(6 * R * R, dtype=torch.int32, pin_memory=True).numpy().view("uint32")
gen = nz.BasePipeline([
    nz.NoiseStage(nz.FractalNoise.Simplex, hurst=0.4, octaves=13, noiseSize=1700),
    nz.KernelFilterStage(nz.KernelFilterType.Gauss5_S1, iterations=17),
    nz.FlowMapStage(iterations=5, normMin=0.0, normMax=0.005),
    nz.ErosionFilterStage(iterations=5),
])
meshp = nz.BasePipeline([nz.MeshTileStage(nz.MeshType.OvershootSquareGridHeightMap)])
mesh = nz.Mesh(); mesh.vertices, mesh.indices = vtx, idx


def one(tile):
    with nz.host.pipeline():
        gen.Run(nz.GeneratorData("t", data, res, 1000 * tile, 0))
        meshp.Run(nz.MeshStageData("t", data, R, res, 4, 1984.375, 2000.0, mesh=mesh))


for t in range(5):
    one(t)
n = 50
t0 = time.perf_counter()
for t in range(n):
    one(t)
dt = (time.perf_counter() - t0) / n
b = data.nbytes + vtx.nbytes + idx.nbytes
print(f"C2 through the stage API: {dt * 1e3:.3f} ms per 1024^2 tile ({res * res / dt / 1e6:.0f} Mcells/s), {b / 1e6:.1f} MB downloaded per tile = {b / dt / 1e9:.1f} GB/s")
