#!/usr/bin/env python
"""bench.py — headline benchmark of the heightmap hot path (BASELINE.json configs, default C5 = configs[4]).

Default workload (C5): ONE 16384^2 heightmap through simplex fBm (13 octaves, hurst 0.4, noiseSize 1700) -> Gauss5 x17 ->
FlowMap x5 (normMin 0, normMax 0.005) -> Value Erosion x5 -> Overshoot mesh (R = 16376), split into row bands
over N GPUs (one process per GPU, NCCL halo exchange between band neighbours).  A "step" is one pass of
that chain.  Metric: Mcells/s = 16384^2 / step time (whole job, all GPUs), plus per-stage Mcells/s and the
end-to-end ms per heightmap.  Scaling is STRONG (the heightmap is fixed, bands shrink with N).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode exchange|recompute]
                    [--config C1|C2|C3|C4|C5]
    torchrun --nproc-per-node N ... bench.py --gpus N ...

--config selects another BASELINE.json config (C1 256^2 noise, C2 1024^2 chain, C3 4096^2 cellular chain, C4 16x16 tiles
of 1024^2 sharded over the N ranks) with the same JSON shape; the default C5 line also carries a short measurement of
C1-C4 under "configs" (C4 sharded over the same N ranks), so one driver run shows all five.

Timing: CUDA events on the stream the kernels run on, barrier + synchronize on both sides, max over
ranks.  Every C5 field is 1 GiB (>> 126 MB L2), so nothing is L2-resident between stages or steps; the small configs
(C1-C4) write a 256 MiB buffer between timed iterations to flush L2.
The `e2e` leg runs the same chain through the reference-facing stage API with pinned HOST buffers
(D2H of the heightmap and the mesh inside the timed region); `e2e_host_input` does the same for the part of the chain
that HAS an input (a host heightmap -> Gauss5 x17 -> FlowMap -> Value Erosion -> mesh), so an H2D is timed as well.
With N > 1 a `band_check` compares this rank's owned rows and mesh slice of the exchange-mode chain against a
recompute-mode run (no communication) bit for bit on the device.
The `cpu_baseline` leg / `--impl reference` time the CPU restatement of the reference (oracle/, -O3 -march=native,
OpenMP over rows like IJobFor) on a bounded sample: the reference itself is C# for Unity/Burst and cannot run here.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_GRID = 16384
FLOP_PER_CELL_SIMPLEX13 = 1942      # 13*(139+10)+5, fma = 2 (ALGORITHMIC: the textbook evaluation)
# What fbm_simplex_pair_kernel actually executes per cell: the hash chain and the gradient fold come from shared-memory
# tables, so the FP32 pipe sees 13 octaves x (13.5 FFMA2 + 7.5 FMUL2 + 4 FADD2 per cell = 77 flop) + ~15
EXECUTED_FLOP_PER_CELL_SIMPLEX13 = 1016
CPU_SAMPLE_N = 4096                 # bounded CPU sample: the same chain on a 4096^2 grid (1/16 of the cells)
# algorithmic FLOP per cell of the other bases, 13 octaves: SURVEY 8d estimates (cellular2 ~330, psrnoise ~120 + 3 sincos per
# octave, + 10 for the fBm wrapper, + 5 per cell); labelled as estimates wherever they are printed
FLOP_PER_CELL_EST = {3: FLOP_PER_CELL_SIMPLEX13, 5: 13 * (330 + 10) + 5, 4: 13 * (120 + 10) + 5}
METRIC = {
    "C1": "Mcells/s (C1: one 256^2 tile, simplex fBm 13 oct)",
    "C2": "Mcells/s (C2 chain on one 1024^2 tile: simplex fBm 13 oct -> Gauss5 x17 -> FlowMap x5 -> Value Erosion x5 -> mesh)",
    "C3": "Mcells/s (C3 chain on a 4096^2 grid: cellular fBm 13 oct -> Gauss5 x17 -> FlowMap x5)",
    "C4": "Mcells/s (C4: 16x16 tiles of 1024^2, rotated-simplex fBm 13 oct -> Gauss3 x3 -> Sobel3_2D -> mesh, tiles sharded over the GPUs)",
    "C5": "Mcells/s (C5 chain: simplex fBm 13 oct -> Gauss5 x17 -> FlowMap x5 -> Value Erosion x5 -> mesh)",
}
WORKLOAD = {
    "C1": "BASELINE.json configs[0]: single 256x256 tile, simplex fBm hurst 0.4, 13 octaves, noiseSize 1700",
    "C2": "BASELINE.json configs[1]: README example #1 full chain on one 1024^2 tile",
    "C3": "BASELINE.json configs[2]: README example #2 on a 4096^2 grid (cellular fBm -> Gauss5 x17 -> FlowMap)",
    "C4": "BASELINE.json configs[3]: 16x16 tiles at 1024^2, rotated-simplex fBm + Gauss3/Sobel2D + mesh, tiles sharded across the GPUs",
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="exchange", choices=["exchange", "recompute"])
    ap.add_argument("--config", default="C5", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--engine", default="lib", choices=["lib", "python"],
                    help="C5 band chain: 'lib' = the in-library chain behind the C ABI (nz_band_*), 'python' = bands.py over torch.distributed")
    ap.add_argument("--n", type=int, default=N_GRID, help="grid resolution (default: the C5 workload, 16384)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the short C1-C4 measurements of the default line")
    ap.add_argument("--no-full-pass", action="store_true", help="reference arm: skip the one full-size 16384^2 pass")
    return ap.parse_args()


def c5_config(args, world):
    return {"workload": f"BASELINE.json configs[4]: one {args.n}^2 heightmap in {world} row band(s), full chain",
            "l2": "every field is 1 GiB (> 126 MB L2): inputs larger than L2, no flush needed",
            "parallelism": f"row bands x{world}"}


def small_config(name, world):
    return {"workload": WORKLOAD[name], "l2": "a 256 MiB buffer is written between timed iterations (L2 flush)",
            "parallelism": f"tiles sharded over {world} GPU(s)" if name == "C4" else "one GPU (a single tile does not shard)"}


# --------------------------------------------------------------------------------------------------
# clocks during the timed region (NVML polled from a thread)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.samples, self.reasons, self.power = [], set(), []
        self.max_mhz = None
        self._stop = threading.Event()
        self._on = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception as e:  # NVML missing: report that instead of inventing clocks
            self.nv = None
            self.err = str(e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            if self._on.is_set():
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                        else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for bit, name in self.REASONS.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(0.004)

    def start(self):
        self._on.set()

    def pause(self):
        self._on.clear()

    def report(self):
        self._stop.set()
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": self.err}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "power_w_max": max(self.power) if self.power else None}


# --------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port of the Burst jobs)
# --------------------------------------------------------------------------------------------------
def host_cores():
    """Cores this process may run on.  Launchers export OMP_NUM_THREADS=1 (torchrun does), which says nothing about
    the box: the reference arm uses every core of its affinity mask, like Unity's job system does."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_oracle():
    import oracle
    oracle.build(fast=True, force=True)     # -march=native must be built on the box that runs it
    o = oracle.get(fast=True)
    cores = o.set_num_threads(host_cores())
    return o, cores


def cpu_chain_once(o, n, noise_type=3, erosion=True, mesh=True):
    """The C5/C2 chain (or C3 without erosion and mesh) on an n x n grid on the host cores; seconds per stage."""
    t = {}
    t0 = time.perf_counter()
    g = o.fractal(n, n, noise_type, 0.4, octaves=13, noise_size=1700)
    t["noise"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    g = o.kernel_filter(g, 2, 17)
    t["filter"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    g = o.flowmap(g, 5, 0.0, 0.005)
    t["flow"] = time.perf_counter() - t0
    if erosion:
        t0 = time.perf_counter()
        g = o.min_erosion(g, 5)
        t["erosion"] = time.perf_counter() - t0
    if mesh:
        t0 = time.perf_counter()
        R = n - 8
        o.heightmap_mesh(1, g, R, 4, 2000.0, R * (500.0 / 256.0))
        t["mesh"] = time.perf_counter() - t0
    return t


def cpu_tile_c4_once(o, tx, tz):
    """One C4 tile on the host cores: rotated-simplex fBm -> Gauss3 x3 -> Sobel3_2D on a copy -> mesh."""
    t0 = time.perf_counter()
    g = o.fractal(1024, 1024, 4, 0.4, octaves=13, xpos=1000 * tx, zpos=1000 * tz, noise_size=1700)
    g = o.kernel_filter(g, 3, 3)
    o.kernel_filter(g, 11, 1)
    o.heightmap_mesh(1, g, 1016, 4, 2000.0, 1016 * (500.0 / 256.0))
    return {"tile": time.perf_counter() - t0}


def cpu_reference(config, steps, warmup, full_pass=False):
    """Times the oracle (fast flavour) on a bounded sample of `config`; returns the cpu_baseline numbers."""
    o, cores = cpu_oracle()
    cpu_chain_once(o, 256)
    if config == "C5":
        n, cells = CPU_SAMPLE_N, CPU_SAMPLE_N ** 2
        once = lambda: cpu_chain_once(o, n)
        sample = (f"same chain and parameters on a {n}x{n} grid (1/{(N_GRID // n) ** 2} of the cells) per step; "
                  "Mcells/s is size-normalised")
    elif config == "C1":
        cells = 256 * 256
        once = lambda: {"noise": _timed(lambda: o.fractal(256, 256, 3, 0.4, octaves=13, noise_size=1700))}
        sample = "the whole config (one 256^2 tile) per step"
    elif config == "C2":
        cells = 1024 * 1024
        once = lambda: cpu_chain_once(o, 1024)
        sample = "the whole config (one 1024^2 tile) per step"
    elif config == "C3":
        cells = 4096 * 4096
        once = lambda: cpu_chain_once(o, 4096, noise_type=5, erosion=False, mesh=False)
        sample = "the whole config (one 4096^2 grid) per step"
    else:
        cells = 8 * 1024 * 1024
        k = [0]

        def once():
            tot = 0.0
            for j in range(8):
                t = (k[0] + j) % 256
                tot += cpu_tile_c4_once(o, t % 16, t // 16)["tile"]
            k[0] += 8
            return {"tiles": tot}
        sample = "8 of the 256 tiles per step (tiles are independent and identical in cost); Mcells/s is size-normalised"
    for _ in range(max(0, warmup - 1)):
        once()
    per = [once() for _ in range(steps)]
    tot = [sum(p.values()) for p in per]
    mean = sum(tot) / len(tot)
    r = {"value": cells / mean / 1e6, "best": cells / min(tot) / 1e6, "ms_per_step": mean * 1e3, "cores": cores,
         "stages_mcells_s": {k: cells / min(p[k] for p in per) / 1e6 for k in per[0]}, "sample": sample + f", mean of {steps} passes"}
    if full_pass and config == "C5":
        r["full_pass"] = cpu_full_pass(o)
    return r


def _timed(fn):
    t0 = time.perf_counter()
    fn()
    return time.perf_counter() - t0


def cpu_full_pass(o):
    """ONE pass of the C5 chain at its full size (16384^2: ~20 s on 16 cores, ~35 GB of host memory with the mesh) so that
    the size-normalised sample above can be checked against the real configuration."""
    try:
        import psutil
        free = psutil.virtual_memory().available
    except Exception:
        free = 0
    if free < 64 * 2 ** 30:
        return {"skipped": f"needs ~35 GB of host memory, {free / 2 ** 30:.0f} GiB available"}
    try:
        t = cpu_chain_once(o, N_GRID)
    except MemoryError as e:
        return {"skipped": f"MemoryError: {e}"}
    s = sum(t.values())
    return {"n": N_GRID, "ms": round(s * 1e3, 1), "mcells_s": round(N_GRID * N_GRID / s / 1e6, 2),
            "stages_mcells_s": {k: round(N_GRID * N_GRID / v / 1e6, 2) for k, v in t.items()}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfgname = args.config
    r = cpu_reference(cfgname, args.steps, args.warmup, full_pass=not args.no_full_pass)
    cpu = {"value": r["value"], "unit": "Mcells/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
           "note": "C++ restatement of the Burst jobs (oracle/), OpenMP over rows with every core of the affinity mask "
                   "(OMP_NUM_THREADS from the launcher is ignored); Unity/Burst cannot run in this image"}
    if "full_pass" in r:
        cpu["full_pass"] = r["full_pass"]
    line = {
        "impl": "reference", "metric": METRIC[cfgname],
        "value": r["value"], "unit": "Mcells/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": c5_config(args, world) if cfgname == "C5" else small_config(cfgname, world),
        "cpu_baseline": cpu,
        "stages_mcells_s": r["stages_mcells_s"],
        "e2e": {"value": r["value"], "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print_line(line)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    return 6650.0, "fallback of B200_PROFILING.md"


def fma_peak_tflops(nz, torch):
    sink = torch.zeros(1 << 20, device="cuda")
    sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    best = 0.0
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        flops = nz.device.fma_peak(sink, sms * 16, 8192)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, flops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


class Env:
    """Process-group plumbing shared by every leg of our arm."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import noize_job_b200 as nz
        from noize_job_b200 import bands
        self.torch, self.dist, self.nz, self.bands = torch, dist, nz, bands
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}: launch with torchrun --nproc-per-node {args.gpus}")
        torch.cuda.set_device(self.local)
        numa = bands.bind_host_to_gpu_numa_node(self.local) if self.world > 1 else ""
        if numa:
            print(f"[rank {self.rank}] {numa}", file=sys.stderr)
        nz.host.init(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.sampler = ClockSampler(self.local) if self.rank == 0 else None
        self.flush_buf = None

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(values, device="cuda", dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def sum_over_ranks(self, values):
        t = self.torch.tensor(values, device="cuda", dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t]

    def flush_l2(self):
        """Write a buffer larger than the 126 MB L2 (small configs only: their fields would otherwise stay L2-resident)."""
        if self.flush_buf is None:
            self.flush_buf = self.torch.empty(64 * 2 ** 20, dtype=self.torch.float32, device="cuda")
        self.flush_buf.fill_(1.0)

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


# ---- the small configs (C1-C4), device-resident, one GPU each except C4 ------------------------------------------------
def small_chain(env, name):
    """Returns (fn, cells, buffers-to-keep-alive) for one device-resident pass of config C1/C2/C3 on this GPU."""
    torch, d = env.torch, env.nz.device
    if name == "C1":
        a = torch.empty(256, 256, device="cuda")
        return (lambda: d.fractal(a, 3, 0.4, octaves=13, noise_size=1700)), 256 * 256
    n = 1024 if name == "C2" else 4096
    a, b = torch.empty(n, n, device="cuda"), torch.empty(n, n, device="cuda")
    R = n - 8
    mesh = (torch.empty((R + 1) * (R + 1), 12, device="cuda"), torch.empty(6 * R * R, dtype=torch.int32, device="cuda")) if name == "C2" else None

    def fn():
        d.fractal(a, 3 if name == "C2" else 5, 0.4, octaves=13, noise_size=1700)
        cur = d.kernel_filter(a, b, 2, 17)
        other = b if cur is a else a
        r = d.flowmap(cur, other, None, 5, 0.0, 0.005)
        if r is not cur:
            cur, other = other, cur
        if name == "C2":
            r = d.min_erosion(cur, other, 5)
            if r is not cur:
                cur, other = other, cur
            d.heightmap_mesh(1, mesh[0], mesh[1], R, n, 4, 2000.0, R * (500.0 / 256.0), cur)
    return fn, n * n


def time_small(env, fn, steps, warmup, sync_all=False):
    """Mean ms of fn() over `steps` iterations, each bracketed by events, L2 flushed in between (untimed)."""
    torch = env.torch
    for _ in range(max(warmup, 3)):
        fn()
    env.barrier() if sync_all else torch.cuda.synchronize()
    tot = 0.0
    for _ in range(steps):
        env.flush_l2()
        torch.cuda.synchronize()          # the timed work may run on streams of its own: the flush must be over
        if sync_all:
            env.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / steps


def run_small_config(env, name, steps, warmup):
    """Device-resident measurement of config C1..C4.  C1-C3 are single tiles: rank 0's GPU runs them (the other ranks wait);
    C4 shards its 256 tiles over all ranks (no communication) and the time is the max over ranks."""
    nz, torch = env.nz, env.torch
    l0 = nz.host.kernel_launch_count()
    if name == "C4":
        from noize_job_b200 import tiles
        cfg = tiles.TileWorldConfig()
        slots = 4
        # the tile loop runs inside the library (nz_tile_world_*): one C call per pass, `slots` tiles in flight
        tw = tiles.LibTileWorld(cfg, env.rank, env.world, slots=slots)
        fn = tw.run
        ms = time_small(env, fn, steps, warmup, sync_all=env.world > 1)
        ms = env.max_over_ranks([ms])[0]
        cells = cfg.tiles_x * cfg.tiles_z * cfg.resolution ** 2
        extra = {"tiles": cfg.tiles_x * cfg.tiles_z, "tiles_per_gpu": len(tw.mine), "streams_per_gpu": slots,
                 "engine": "nz_tile_world (C ABI)"}
        tw.release()
    else:
        ms, cells, extra = 0.0, 0, {}
        if env.rank == 0:
            fn, cells = small_chain(env, name)
            ms = time_small(env, fn, steps, warmup)
        ms, cells = env.max_over_ranks([ms, cells])
        cells = int(cells)
    launches = nz.host.kernel_launch_count() - l0
    torch.cuda.empty_cache()
    return {"ms": round(ms, 4), "mcells_s": round(cells / ms / 1e3, 1), "n_gpus": env.world if name == "C4" else 1,
            "gpu_launches": int(launches), **extra}


def e2e_small_config(env, name, steps=3):
    """Config C1..C3 through the stage API with pinned host buffers on rank 0 (D2H of the tile and its mesh timed)."""
    nz, torch = env.nz, env.torch
    if env.rank != 0 or name == "C4":
        return None
    n = {"C1": 256, "C2": 1024, "C3": 4096}[name]
    R = n - 8
    data = torch.empty(n * n, dtype=torch.float32, pin_memory=True).numpy()
    stages = [nz.NoiseStage(nz.FractalNoise.Cellular if name == "C3" else nz.FractalNoise.Simplex, hurst=0.4, octaves=13, noiseSize=1700)]
    if name != "C1":
        stages += [nz.KernelFilterStage(nz.KernelFilterType.Gauss5_S1, iterations=17), nz.FlowMapStage(iterations=5, normMin=0.0, normMax=0.005)]
    if name == "C2":
        stages.append(nz.ErosionFilterStage(iterations=5))
    gen = nz.BasePipeline(stages)
    d2h = data.nbytes
    if name == "C2":
        mesh = nz.Mesh()
        mesh.vertices = torch.empty((R + 1) * (R + 1), 12, dtype=torch.float32, pin_memory=True).numpy()
        mesh.indices = torch.empty(6 * R * R, dtype=torch.int32, pin_memory=True).numpy().view("uint32")
        meshp = nz.BasePipeline([nz.MeshTileStage(nz.MeshType.OvershootSquareGridHeightMap)])
        d2h += mesh.vertices.nbytes + mesh.indices.nbytes

    if name == "C2":
        stages[-1].keepResident = True          # the mesh pipeline works on the same uuid and closes the scope

    def one():
        gen.Run(nz.GeneratorData("bench", data, n, 0, 0))
        if name == "C2":
            meshp.Run(nz.MeshStageData("bench", data, R, n, 4, R * (500.0 / 256.0), 2000.0, mesh=mesh))
    one()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / steps
    return {"value": round(n * n / dt / 1e6, 1), "unit": "Mcells/s", "ms_per_step": round(dt * 1e3, 4), "h2d_bytes_per_step": 0,
            "d2h_bytes_per_step": int(d2h), "steps": steps, "api": "stage API over the C-ABI host layer (pinned host buffers)"}


def run_ours_small(args, env):
    name = args.config
    res = run_small_config(env, name, args.steps, args.warmup)
    e2e = None if args.no_e2e else e2e_small_config(env, name)
    if env.rank != 0:
        return
    cpu = None
    if not args.no_cpu:
        r = cpu_reference(name, 1, 1)
        cpu = {"value": round(r["value"], 2), "unit": "Mcells/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
    line = {
        "metric": METRIC[name], "value": res["mcells_s"], "unit": "Mcells/s", "n_gpus": env.world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": res["ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": small_config(name, env.world), "detail": res,
        "roofline": None, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": res["gpu_launches"],
        "clocks": env.sampler.report() if env.sampler else None,
    }
    print_line(line)


# ---- C5 -----------------------------------------------------------------------------------------------------------------
def make_chain(args, env, cfg, mode=None):
    bands = env.bands
    mode = mode or args.mode
    if args.engine == "lib" and hasattr(bands, "LibBandChain"):
        return bands.LibBandChain(cfg, env.rank, env.world, env.dist if env.world > 1 else None, mode=mode)
    return bands.BandChain(cfg, bands.CudaEngine(), env.rank, env.world, env.dist if env.world > 1 else None, mode=mode)


def band_check(args, env, cfg, chain):
    """Exchange-mode result (owned heightmap rows + this rank's mesh slice) against a recompute-mode run of the plain
    Python band chain (no communication; bit-exact against one grid by tests/test_gpu_bands.py), on the device."""
    torch, bands = env.torch, env.bands
    if env.world == 1:
        return "single band: nothing to compare"
    chain.run()
    own = chain.owned().clone()
    vtx, idx = chain.vtx.clone(), chain.idx.clone()
    ref = bands.BandChain(cfg, bands.CudaEngine(), env.rank, env.world, None, mode="recompute")
    ref.run()
    torch.cuda.synchronize()
    same = bool(torch.equal(own, ref.owned())) and bool(torch.equal(vtx.view(torch.int32), ref.vtx.view(torch.int32))) \
        and bool(torch.equal(idx, ref.idx))
    del ref, own, vtx, idx
    torch.cuda.empty_cache()
    ok = env.sum_over_ranks([0.0 if same else 1.0])[0] == 0.0
    return "bit-exact" if ok else "DIFFERS"


def run_ours(args):
    env = Env(args)
    if args.config != "C5":
        run_ours_small(args, env)
        env.close()
        return
    torch, dist, nz, bands = env.torch, env.dist, env.nz, env.bands
    world, rank, sampler, barrier = env.world, env.rank, env.sampler, env.barrier
    N = args.n
    cfg = bands.ChainConfig(N=N)
    chain = make_chain(args, env, cfg)

    fma_peak = fma_peak_tflops(nz, torch) if rank == 0 else 0.0

    # ---- device-resident timing -------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        chain.run()
    barrier()
    check = band_check(args, env, cfg, chain)
    barrier()
    names = ["noise", "filter", "flow", "erosion", "mesh", "end"]
    launches0 = nz.host.kernel_launch_count()
    reruns0 = nz.device.flow_walk_reruns()
    if sampler:
        sampler.start()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    per_step = [chain.run_timed() for _ in range(args.steps)]
    t_end.record()
    barrier()
    if sampler:
        sampler.pause()
    launches = nz.host.kernel_launch_count() - launches0
    ms_step = t_start.elapsed_time(t_end) / args.steps
    stage_ms = [sum(p()[i] for p in per_step) / args.steps for i in range(len(names) - 1)]
    if os.environ.get("NZ_BENCH_VERBOSE"):
        print(f"[rank {rank}] ms_step {ms_step:.3f} stages " + " ".join(f"{n}={v:.3f}" for n, v in zip(names, stage_ms)), file=sys.stderr)
    own_noise_ms, own_mesh_ms = stage_ms[0], stage_ms[4]
    red = env.max_over_ranks([ms_step] + stage_ms)
    ms_step, stage_ms = red[0], red[1:]

    # facts about this rank's band that the report needs after the chain is gone
    env.own_rows, env.vz0, env.vz1 = chain.own, chain.vz0, chain.vz1
    # ghost-row bytes RECEIVED per step, summed over the ranks (an inner band receives from both neighbours)
    env.halo_bytes_per_step = int(env.sum_over_ranks([chain.bytes_exchanged // max(1, chain.runs)])[0])
    # launches of the register-walk flow map that left its guarded fast paths and were redone on the wavefront kernel
    # (exact either way; 0 on ordinary terrain, garbage ghost rows of a band included), summed over the ranks
    env.flow_walk_reruns = int(env.sum_over_ranks([nz.device.flow_walk_reruns() - reruns0])[0])
    env.engine_name = chain.name

    # ---- end to end through the public API with host buffers ------------------------------------------------
    e2e = e2e_in = None
    if not args.no_e2e:
        e2e, e2e_in = run_e2e(args, env, chain, cfg)

    # ---- the other BASELINE configs, briefly (C4 sharded over the same ranks) -----------------------------------
    configs = None
    if not args.no_configs:
        del chain
        torch.cuda.empty_cache()
        configs = {}
        for name in ("C1", "C2", "C3", "C4"):
            configs[name] = run_small_config(env, name, steps=5, warmup=3)
            configs[name]["workload"] = WORKLOAD[name]
        chain = None

    if rank != 0:
        env.close()
        return

    cells = N * N
    hbm_peak, hbm_src = measured_peaks()
    R = cfg.R
    stage_bytes = {"filter": 8 * cells, "flow": 8 * cells, "erosion": 8 * cells,
                   "mesh": 4 * cells + 48 * (R + 1) ** 2 + 24 * R * R}
    stages = {}
    # per-stage fractions are whole-job throughput against the peaks of all `world` GPUs
    hbm_peak_all, fma_peak_all = hbm_peak * world, fma_peak * world
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    traffic_file = json.load(open(tpath)) if os.path.exists(tpath) else {}
    stage_traffic = traffic_file.get("stage_dram_bytes_per_step", {})
    for name, ms in zip(names[:-1], stage_ms):
        s = {"ms": round(ms, 4), "mcells_s": round(cells / ms / 1e3, 1)}
        if name == "noise":
            ach = FLOP_PER_CELL_SIMPLEX13 * cells / ms / 1e9
            s.update(bound="fp32", achieved_tflops=round(ach, 2), peak_tflops=round(fma_peak_all, 2), frac=round(ach / fma_peak_all, 4))
        else:
            ach = stage_bytes[name] / ms / 1e6
            s.update(bound="hbm", compulsory_gbs=round(ach, 1), peak_gbs=hbm_peak_all, frac=round(ach / hbm_peak_all, 4))
            # once the iterations are fused the stage's roofline is max(t_HBM, t_FP32); for Gauss5 x17
            # (340 FLOP/cell) and FlowMap x5 (45 FLOP/cell/iteration) the FP32 time is the larger one
            flop_per_cell = {"filter": 2 * (2 * cfg.filter_radius + 1) * 2 * cfg.filter_iterations,
                             "flow": 45 * cfg.flow_iterations}.get(name)
            if flop_per_cell and fma_peak > 0:
                t_hbm = stage_bytes[name] / (hbm_peak_all * 1e6)            # ms
                t_fp32 = flop_per_cell * cells / (fma_peak_all * 1e9)        # ms
                if t_fp32 > t_hbm:
                    s.update(bound="fp32 (iterations fused; the HBM figures are kept for reference)", flop_per_cell=flop_per_cell,
                             achieved_tflops=round(flop_per_cell * cells / ms / 1e9, 2), peak_tflops=round(fma_peak_all, 2),
                             roofline_ms=round(t_fp32, 4), hbm_frac=s["frac"], frac=round(t_fp32 / ms, 4))
        # DRAM bytes the stage's kernels moved per step at N = 16384 on one GPU (ncu --set full, profiles/traffic.json)
        s["traffic"] = stage_traffic.get(name) if (N == N_GRID and world == 1) else None
        stages[name] = s
    # dominant kernel: the fBm evaluator (one launch per step on this rank's band)
    own_rows, vz0, vz1 = env.own_rows, env.vz0, env.vz1
    own_cells = own_rows * N
    ach = FLOP_PER_CELL_SIMPLEX13 * own_cells / own_noise_ms / 1e9
    ach_exec = EXECUTED_FLOP_PER_CELL_SIMPLEX13 * own_cells / own_noise_ms / 1e9
    roofline = {"kernel": "fbm_simplex_pair_kernel", "bound": "fp32", "achieved": round(ach, 3), "peak": round(fma_peak, 3),
                "unit": "TFLOP/s", "frac": round(ach / fma_peak, 4), "traffic": traffic_file.get("fbm_kernel_dram_bytes_per_launch"),
                "executed": round(ach_exec, 3), "executed_frac": round(ach_exec / fma_peak, 4),
                "executed_note": "achieved/frac count the ALGORITHMIC flops of the textbook simplex (SURVEY 8d); the kernel replaces "
                                 "the hash chain and gradient fold by table walks, so the FP32 pipe executes ~1016 flop/cell",
                "peak_source": "builder-measured: FFMA micro-benchmark run in this process (nz_dev_fma_peak), burst; MEASURED_PEAKS.json "
                               "has no FP32 entry (nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4); tensor cores unused: no stage is a contraction",
                "flop_per_cell": FLOP_PER_CELL_SIMPLEX13, "cells_per_launch": own_cells, "ms_per_launch": round(own_noise_ms, 4),
                "hbm_peak_gbs": hbm_peak, "hbm_peak_source": hbm_src}

    # the dominant HBM-bound kernel (mesh_kernel: one launch per step) in the same shape as `roofline`
    nvr = vz1 - vz0
    mesh_bytes = 4 * (nvr + 2) * N + 48 * nvr * (R + 1) + 24 * R * max(vz1 - max(vz0, 1), 0)
    roofline_hbm = {"kernel": "mesh_kernel", "bound": "hbm", "achieved": round(mesh_bytes / own_mesh_ms / 1e6, 1), "peak": hbm_peak,
                    "unit": "GB/s", "frac": round(mesh_bytes / own_mesh_ms / 1e6 / hbm_peak, 4),
                    "traffic": stage_traffic.get("mesh") if (N == N_GRID and world == 1) else None,
                    "bytes_per_launch": int(mesh_bytes), "ms_per_launch": round(own_mesh_ms, 4), "peak_source": hbm_src}

    cpu = None
    if not args.no_cpu and world == 1:
        r = cpu_reference("C5", 1, 1)
        cpu = {"value": round(r["value"], 2), "unit": "Mcells/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
               "stages_mcells_s": {k: round(v, 2) for k, v in r["stages_mcells_s"].items()}}

    conf = c5_config(args, world)
    conf["halo"] = "exchange over NCCL" if args.mode == "exchange" else "recompute"
    conf["engine"] = env.engine_name
    line = {
        "metric": METRIC["C5"],
        "value": round(cells / ms_step / 1e3, 1), "unit": "Mcells/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": conf,
        "stages": stages, "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu, "e2e": e2e,
        "e2e_host_input": e2e_in, "band_check": check, "configs": configs, "gpu_launches": int(launches),
        "halo_bytes_per_step": int(env.halo_bytes_per_step),
        "flow_walk_reruns": env.flow_walk_reruns,
        "parity": "GPU == oracle within the per-stage tolerances (tests/); oracle == Burst reference UNPINNED at bit level (no Unity here)",
        "clocks": sampler.report() if sampler else None,
    }
    print_line(line)
    env.close()


def run_e2e(args, env, chain, cfg):
    """Same chain, host buffers in and out.  1 GPU: the reference-facing stage API (NoiseStage ... MeshTileStage
    over the C ABI host layer).  N GPUs: the band chain + D2H of the owned band and mesh slice into pinned memory.
    Returns (e2e, e2e_host_input): the second leg starts from a HOST heightmap (Gauss5 x17 -> flow -> erosion -> mesh), so
    its timed region has an H2D as well; it runs on one GPU (a host-input chain on N GPUs is N uploads of a band)."""
    torch, dist, nz = env.torch, env.dist, env.nz
    world, sampler, barrier = env.world, env.sampler, env.barrier
    N, R = cfg.N, cfg.R
    steps = max(1, min(args.steps, 3))
    e2e_in = None
    if world == 1:
        chain.release()
        torch.cuda.empty_cache()
        data = torch.empty(N * N, dtype=torch.float32, pin_memory=True).numpy()
        vtx = torch.empty((R + 1) * (R + 1), 12, dtype=torch.float32, pin_memory=True).numpy()
        idx = torch.empty(6 * R * R, dtype=torch.int32, pin_memory=True).numpy().view("uint32")
        def tail():
            st = [nz.KernelFilterStage(nz.KernelFilterType.Gauss5_S1, iterations=cfg.filter_iterations),
                  nz.FlowMapStage(iterations=cfg.flow_iterations, normMin=cfg.norm_min, normMax=cfg.norm_max),
                  nz.ErosionFilterStage(iterations=cfg.erosion_iterations)]
            # residency is owned by the stage objects (GpuStage): the last generator stage brings the heightmap home but
            # keeps the tile in HBM for the mesh pipeline, which works on the same uuid and closes the scope
            st[-1].keepResident = True
            return st
        gen = nz.BasePipeline([nz.NoiseStage(nz.FractalNoise.Simplex, hurst=cfg.hurst, octaves=cfg.octaves, noiseSize=cfg.noise_size)] + tail())
        meshp = nz.BasePipeline([nz.MeshTileStage(nz.MeshType.OvershootSquareGridHeightMap)])
        mesh = nz.Mesh()
        mesh.vertices, mesh.indices = vtx, idx

        def one():
            gen.Run(nz.GeneratorData("bench", data, N, 0, 0))
            meshp.Run(nz.MeshStageData("bench", data, R, N, cfg.mesh_margin, cfg.tile_size, cfg.tile_height, mesh=mesh))
        d2h = data.nbytes + vtx.nbytes + idx.nbytes

        # second leg: the chain from a host-resident heightmap (what a Burst NoiseStage, a loaded .data file or a
        # texture hands to the first GPU stage).  The input is refreshed from a second pinned copy outside the timing.
        src = torch.empty(N * N, dtype=torch.float32, pin_memory=True).numpy()
        filt = nz.BasePipeline(tail())

        def one_in():
            filt.Run(nz.GeneratorData("bench-in", data, N, 0, 0))
            meshp.Run(nz.MeshStageData("bench-in", data, R, N, cfg.mesh_margin, cfg.tile_size, cfg.tile_height, mesh=mesh))
    else:
        own = torch.empty(chain.own, N, dtype=torch.float32, pin_memory=True)
        chain.run()
        hv = torch.empty_like(chain.vtx, device="cpu", pin_memory=True)
        hi = torch.empty_like(chain.idx, device="cpu", pin_memory=True)

        def one():
            chain.run()
            own.copy_(chain.owned(), non_blocking=True)
            hv.copy_(chain.vtx, non_blocking=True)
            hi.copy_(chain.idx, non_blocking=True)
            torch.cuda.synchronize()
        d2h = own.numel() * 4 + hv.numel() * 4 + hi.numel() * 4

    def timed(fn, before=None):
        fn()
        barrier()
        if sampler:
            sampler.start()
        tot = 0.0
        for _ in range(steps):
            if before:
                before()
            barrier()
            t0 = time.perf_counter()
            fn()
            barrier()
            tot += time.perf_counter() - t0
        if sampler:
            sampler.pause()
        return tot / steps

    dt = timed(one)
    dt = env.max_over_ranks([dt])[0]
    d2h_all = env.sum_over_ranks([float(d2h)])[0]
    e2e = {"value": round(N * N / dt / 1e6, 1), "unit": "Mcells/s", "ms_per_step": round(dt * 1e3, 3),
           "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(d2h_all), "steps": steps,
           "api": "stage API over the C-ABI host layer (pinned host buffers)" if world == 1 else f"{chain.name} + D2H into pinned host buffers",
           "note": "the chain starts from a noise generator, so there is no input to upload; the heightmap and the mesh are downloaded every step"}
    if world == 1:
        import numpy as np
        src[:] = data                                # a real heightmap (the chain's own output) as the host input
        dti = timed(one_in, before=lambda: np.copyto(data, src))
        e2e_in = {"value": round(N * N / dti / 1e6, 1), "unit": "Mcells/s", "ms_per_step": round(dti * 1e3, 3),
                  "h2d_bytes_per_step": int(data.nbytes), "d2h_bytes_per_step": int(d2h), "steps": steps,
                  "workload": "host heightmap 16384^2 -> Gauss5 x17 -> FlowMap x5 -> Value Erosion x5 -> mesh (the C5 chain without its generator)",
                  "api": "stage API over the C-ABI host layer (pinned host buffers)"}
    return e2e, e2e_in


def main():
    args = parse()
    # stdout carries exactly ONE JSON line: libraries that chat on fd 1 (NCCL prints its version there) go to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    out = os.fdopen(real_stdout, "w")
    global print_line

    def print_line(obj):
        out.write(json.dumps(obj) + "\n")
        out.flush()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
