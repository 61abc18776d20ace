"""The register-walk flow map reruns a launch on the wavefront kernel when a lane leaves its guarded fast paths.  On a row band
whose window edge is not a grid edge the interior body runs over ghost rows whose values are garbage by design: this checks
that the garbage never raises the flag on the bench terrain, for 1, 2, 4 and 8 bands (all on one GPU: devices repeat)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import noize_job_b200 as nz
from noize_job_b200 import bands
cfg = bands.ChainConfig(N=16384)
for nb in (1, 2, 4, 8):
    before = nz.device.flow_walk_reruns()
    ch = bands.LibBandChain(cfg, devices=[0] * nb) if nb > 1 else bands.LibBandChain(cfg)
    for _ in range(4):
        ch.run()
    torch.cuda.synchronize()
    print(f"{nb} band(s): flow-walk reruns {nz.device.flow_walk_reruns() - before} in 4 passes", flush=True)
    ch.release()
