"""Do two C5 chains in flight on two streams of ONE GPU overlap (the mesh stage is HBM-bound, the noise stage FP32-bound)?
Times K passes of one chain against K passes alternating between two chains on separate streams."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import noize_job_b200 as nz
from noize_job_b200 import bands
K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = bands.ChainConfig(N=16384)
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
chains = []
for s in streams:
    with torch.cuda.stream(s):
        chains.append(bands.LibBandChain(cfg))
def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.default_stream())
    for s in streams:
        s.wait_event(e0)
    fn()
    for s in streams:
        torch.cuda.default_stream().wait_stream(s)
    e1.record(torch.cuda.default_stream())
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)
def one():
    for _ in range(K):
        chains[0].run()
def two():
    for k in range(K):
        chains[k & 1].run()
for _ in range(3):
    chains[0].run(); chains[1].run()
torch.cuda.synchronize()
for rep in range(3):
    a = timed(one) / K
    b = timed(two) / K
    print(f"one chain {a:.3f} ms/pass   two chains alternating {b:.3f} ms/pass   ({100 * (a - b) / a:+.1f} %)", flush=True)
