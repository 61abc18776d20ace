"""Dumps every stage output of README example #1 (GPU path) in the reference's PipelineState format, so that a Unity
run of the same pipeline (PipelineSerdeManager.WriteData of the tile after each stage) can be diffed byte for byte.

    python tools/dump_chain.py <out_dir> [resolution] [xpos] [zpos]

Writes <out_dir>/save__noize_b200/files.json and data/{noise,gauss5x17,erosion5,flow5}.data (raw little-endian f32).
Parameters are the inspector panel of docs~/3.jpg-6.jpg: Simplex, hurst 0.422, 13 octaves, noise size 1757, (0, 424).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import noize_job_b200 as nz  # noqa: E402
from noize_job_b200 import serde  # noqa: E402

out = sys.argv[1]
res = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
xpos = int(sys.argv[3]) if len(sys.argv) > 3 else 0
zpos = int(sys.argv[4]) if len(sys.argv) > 4 else 424
m = serde.PipelineSerdeManager(out, "noize_b200", nz.host.version())
data = np.zeros(res * res, np.float32)
stages = [("noise", nz.NoiseStage(nz.FractalNoise.Simplex, hurst=0.422, octaves=13, noiseSize=1757)),
          ("gauss5x17", nz.KernelFilterStage(nz.KernelFilterType.Gauss5_S1, iterations=17)),
          ("erosion5", nz.ErosionFilterStage(iterations=5)),
          ("flow5", nz.FlowMapStage(iterations=5, normMin=0.0, normMax=0.005))]
for name, stage in stages:
    nz.BasePipeline([stage]).Run(nz.GeneratorData(name, data, res, xpos, zpos))
    m.WriteData(data, name)
    print(f"{name:10s} min {data.min():.6g} max {data.max():.6g} -> {m.GetFQN(name)}")
