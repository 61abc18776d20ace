#!/usr/bin/env python
"""bench.py — headline benchmark of the heightmap hot path (BASELINE.json config C5).

Workload: ONE 16384^2 heightmap through simplex fBm (13 octaves, hurst 0.4, noiseSize 1700) -> Gauss5 x17 ->
FlowMap x5 (normMin 0, normMax 0.005) -> Value Erosion x5 -> Overshoot mesh (R = 16376), split into row bands
over N GPUs (one process per GPU, NCCL halo exchange between band neighbours).  A "step" is one pass of
that chain.  Metric: Mcells/s = 16384^2 / step time (whole job, all GPUs), plus per-stage Mcells/s and the
end-to-end ms per heightmap.  Scaling is STRONG (the heightmap is fixed, bands shrink with N).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode exchange|recompute]
    torchrun --nproc-per-node N ... bench.py --gpus N ...

Timing: CUDA events on the stream the kernels run on, barrier + synchronize on both sides, max over
ranks.  Every field is 1 GiB (>> 126 MB L2), so nothing is L2-resident between stages or steps.
The `e2e` leg runs the same chain through the reference-facing stage API with pinned HOST buffers
(D2H of the heightmap and the mesh inside the timed region).  The `cpu_baseline` leg / `--impl reference`
time the CPU restatement of the reference (oracle/, -O3 -march=native, OpenMP over rows like IJobFor) on a
bounded sample: the reference itself is C# for Unity/Burst and cannot run here.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="exchange", choices=["exchange", "recompute"])
    ap.add_argument("--n", type=int, default=N_GRID, help="grid resolution (default: the C5 workload, 16384)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


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
def cpu_chain_once(o, n):
    """The C5 chain on an n x n grid on the host cores; returns seconds per stage."""
    t = {}
    t0 = time.perf_counter()
    g = o.fractal(n, n, 3, 0.4, octaves=13, noise_size=1700)
    t["noise"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    g = o.kernel_filter(g, 2, 17)
    t["filter"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    g = o.flowmap(g, 5, 0.0, 0.005)
    t["flow"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    g = o.min_erosion(g, 5)
    t["erosion"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    R = n - 8
    o.heightmap_mesh(1, g, R, 4, 2000.0, R * (500.0 / 256.0))
    t["mesh"] = time.perf_counter() - t0
    return t


def cpu_reference(steps, warmup, n=CPU_SAMPLE_N):
    import oracle
    oracle.build(fast=True, force=True)     # -march=native must be built on the box that runs it
    o = oracle.get(fast=True)
    cpu_chain_once(o, 256)
    for _ in range(max(0, warmup - 1)):
        cpu_chain_once(o, n)
    per = []
    for _ in range(steps):
        per.append(cpu_chain_once(o, n))
    tot = [sum(p.values()) for p in per]
    best = min(tot)
    mean = sum(tot) / len(tot)
    stages = {k: n * n / min(p[k] for p in per) / 1e6 for k in per[0]}
    return {"value": n * n / mean / 1e6, "best": n * n / best / 1e6, "ms_per_step": mean * 1e3, "cores": o.num_threads(),
            "stages_mcells_s": stages,
            "sample": f"same chain and parameters on a {n}x{n} grid (1/{(N_GRID // n) ** 2} of the cells), "
                      f"mean of {steps} passes; Mcells/s is size-normalised"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "Mcells/s (C5 chain: simplex fBm 13 oct -> Gauss5 x17 -> FlowMap x5 -> Value Erosion x5 -> mesh)",
        "value": r["value"], "unit": "Mcells/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE.json configs[4]: one 16384^2 heightmap, full chain; CPU arm runs a bounded "
                               f"{CPU_SAMPLE_N}^2 sample per step"},
        "cpu_baseline": {"value": r["value"], "unit": "Mcells/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
                         "note": "C++ restatement of the Burst jobs (oracle/), OpenMP over rows; Unity/Burst cannot run in this image"},
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
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    best = 0.0
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        flops = nz.device.fma_peak(sink, sms * 16, 8192)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, flops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def run_ours(args):
    import torch
    import torch.distributed as dist
    import noize_job_b200 as nz
    from noize_job_b200 import bands

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    numa = bands.bind_host_to_gpu_numa_node(local) if world > 1 else ""
    if numa:
        print(f"[rank {rank}] {numa}", file=sys.stderr)
    nz.host.init(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N = args.n
    cfg = bands.ChainConfig(N=N)
    chain = bands.BandChain(cfg, bands.CudaEngine(), rank, world, dist if world > 1 else None, mode=args.mode)
    sampler = ClockSampler(local) if rank == 0 else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fma_peak = fma_peak_tflops(nz, torch) if rank == 0 else 0.0

    # ---- device-resident timing -------------------------------------------------------------------------
    for _ in range(args.warmup):
        chain.run()
    barrier()
    names = ["noise", "filter", "flow", "erosion", "mesh", "end"]
    evs = [[torch.cuda.Event(enable_timing=True) for _ in names] for _ in range(args.steps)]
    step = [0]

    def mark(name):
        evs[step[0]][names.index(name)].record()

    launches0 = nz.host.kernel_launch_count()
    if sampler:
        sampler.start()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for k in range(args.steps):
        step[0] = k
        chain.run(mark)
    t_end.record()
    barrier()
    if sampler:
        sampler.pause()
    launches = nz.host.kernel_launch_count() - launches0
    ms_step = t_start.elapsed_time(t_end) / args.steps
    stage_ms = [sum(evs[k][i].elapsed_time(evs[k][i + 1]) for k in range(args.steps)) / args.steps for i in range(len(names) - 1)]
    if os.environ.get("NZ_BENCH_VERBOSE"):
        print(f"[rank {rank}] ms_step {ms_step:.3f} stages " + " ".join(f"{n}={v:.3f}" for n, v in zip(names, stage_ms)), file=sys.stderr)
    t = torch.tensor([ms_step] + stage_ms, device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step, stage_ms = float(t[0]), [float(v) for v in t[1:]]
    own_noise_ms = sum(evs[k][0].elapsed_time(evs[k][1]) for k in range(args.steps)) / args.steps

    # ---- end to end through the public API with host buffers ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, nz, torch, dist, chain, cfg, rank, world, barrier, sampler)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
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
    stage_traffic = json.load(open(tpath)).get("stage_dram_bytes_per_step", {}) if os.path.exists(tpath) else {}
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
    own_cells = chain.own * N
    ach = FLOP_PER_CELL_SIMPLEX13 * own_cells / own_noise_ms / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("fbm_kernel_dram_bytes_per_launch")
    ach_exec = EXECUTED_FLOP_PER_CELL_SIMPLEX13 * own_cells / own_noise_ms / 1e9
    roofline = {"kernel": "fbm_simplex_pair_kernel", "bound": "fp32", "achieved": round(ach, 3), "peak": round(fma_peak, 3),
                "unit": "TFLOP/s", "frac": round(ach / fma_peak, 4), "traffic": traffic,
                "executed": round(ach_exec, 3), "executed_frac": round(ach_exec / fma_peak, 4),
                "executed_note": "achieved/frac count the ALGORITHMIC flops of the textbook simplex (SURVEY 8d); the kernel replaces "
                                 "the hash chain and gradient fold by table walks, so the FP32 pipe executes ~1016 flop/cell",
                "peak_source": "FFMA micro-benchmark run in this process (nz_dev_fma_peak), burst; tensor cores unused: no stage is a contraction",
                "flop_per_cell": FLOP_PER_CELL_SIMPLEX13, "cells_per_launch": own_cells, "ms_per_launch": round(own_noise_ms, 4),
                "hbm_peak_gbs": hbm_peak, "hbm_peak_source": hbm_src}

    # the dominant HBM-bound kernel (mesh_kernel: one launch per step) in the same shape as `roofline`
    mesh_ms = sum(evs[k][4].elapsed_time(evs[k][5]) for k in range(args.steps)) / args.steps
    nvr = chain.vz1 - chain.vz0
    mesh_bytes = 4 * (nvr + 2) * N + 48 * nvr * (R + 1) + 24 * R * max(chain.vz1 - max(chain.vz0, 1), 0)
    roofline_hbm = {"kernel": "mesh_kernel", "bound": "hbm", "achieved": round(mesh_bytes / mesh_ms / 1e6, 1), "peak": hbm_peak,
                    "unit": "GB/s", "frac": round(mesh_bytes / mesh_ms / 1e6 / hbm_peak, 4),
                    "traffic": stage_traffic.get("mesh") if (N == N_GRID and world == 1) else None,
                    "bytes_per_launch": int(mesh_bytes), "ms_per_launch": round(mesh_ms, 4), "peak_source": hbm_src}

    cpu = None
    if not args.no_cpu and world == 1:
        r = cpu_reference(1, 1)
        cpu = {"value": round(r["value"], 2), "unit": "Mcells/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
               "stages_mcells_s": {k: round(v, 2) for k, v in r["stages_mcells_s"].items()}}

    line = {
        "metric": "Mcells/s (C5 chain: simplex fBm 13 oct -> Gauss5 x17 -> FlowMap x5 -> Value Erosion x5 -> mesh)",
        "value": round(cells / ms_step / 1e3, 1), "unit": "Mcells/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"BASELINE.json configs[4]: one {N}^2 heightmap in {world} row band(s), full chain, mode={args.mode}",
                   "l2": "every field is 1 GiB (> 126 MB L2): inputs larger than L2, no flush needed",
                   "parallelism": f"row bands x{world}, halo {'exchange over NCCL' if args.mode == 'exchange' else 'recompute'}"},
        "stages": stages, "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        "halo_bytes_per_step": int(chain.bytes_exchanged // max(1, args.steps + args.warmup)),
        "clocks": sampler.report() if sampler else None,
    }
    print_line(line)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, nz, torch, dist, chain, cfg, rank, world, barrier, sampler):
    """Same chain, host buffers in and out.  1 GPU: the reference-facing stage API (NoiseStage ... MeshTileStage
    over the C ABI host layer).  N GPUs: BandChain + D2H of the owned band and mesh slice into pinned memory."""
    N, R = cfg.N, cfg.R
    steps = max(1, min(args.steps, 3))
    if world == 1:
        del chain.buf_a, chain.buf_b, chain.vtx, chain.idx, chain.flow_scratch
        torch.cuda.empty_cache()
        data = torch.empty(N * N, dtype=torch.float32, pin_memory=True).numpy()
        vtx = torch.empty((R + 1) * (R + 1), 12, dtype=torch.float32, pin_memory=True).numpy()
        idx = torch.empty(6 * R * R, dtype=torch.int32, pin_memory=True).numpy().view("uint32")
        gen = nz.BasePipeline([
            nz.NoiseStage(nz.FractalNoise.Simplex, hurst=cfg.hurst, octaves=cfg.octaves, noiseSize=cfg.noise_size),
            nz.KernelFilterStage(nz.KernelFilterType.Gauss5_S1, iterations=cfg.filter_iterations),
            nz.FlowMapStage(iterations=cfg.flow_iterations, normMin=cfg.norm_min, normMax=cfg.norm_max),
            nz.ErosionFilterStage(iterations=cfg.erosion_iterations),
        ])
        meshp = nz.BasePipeline([nz.MeshTileStage(nz.MeshType.OvershootSquareGridHeightMap)])
        mesh = nz.Mesh()
        mesh.vertices, mesh.indices = vtx, idx

        def one():
            # one outer residency scope: the heightmap stays in HBM between the generator and the mesh pipeline
            with nz.host.pipeline():
                gen.Run(nz.GeneratorData("bench", data, N, 0, 0))
                meshp.Run(nz.MeshStageData("bench", data, R, N, cfg.mesh_margin, cfg.tile_size, cfg.tile_height, mesh=mesh))
        d2h = data.nbytes + vtx.nbytes + idx.nbytes
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
    one()
    barrier()
    if sampler:
        sampler.start()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    barrier()
    dt = (time.perf_counter() - t0) / steps
    if sampler:
        sampler.pause()
    t = torch.tensor([dt, float(d2h)], device="cuda", dtype=torch.float64)
    if world > 1:
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dt, d2h = float(mx[0]), float(sm[1])
    return {"value": round(N * N / dt / 1e6, 1), "unit": "Mcells/s", "ms_per_step": round(dt * 1e3, 3),
            "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(d2h), "steps": steps,
            "api": "stage API over the C-ABI host layer (pinned host buffers)" if world == 1 else "BandChain + D2H into pinned host buffers",
            "note": "the chain starts from a noise generator, so there is no input to upload; the heightmap and the mesh are downloaded every step"}


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
