"""Row-band partition of one large heightmap across the GPUs of a box (BASELINE config C5).

One process per GPU (torch.distributed).  Rank b owns rows [b*N/g, (b+1)*N/g) of the N x N grid.
Every stage kernel of libnoize_b200 works on a rectangular window, and clamp-to-edge at a window
border that is NOT a border of the full grid only corrupts r cells per stencil iteration, so a
band runs the ordinary single-GPU kernels on (own rows + ghost rows) and keeps its own rows:

    stage            ghost rows needed            source
    noise            0                            pure function of (x, z): no communication
    Gauss K x I      r*I above and below          neighbour's boundary rows (halo exchange) or recompute
    flow map x I     2*I+1 above and below        (outflow step 1 + water step 1 per iteration, +1 velocity)
    min erosion x I  I above only                 (trailing window)
    mesh             1 above and below            (normals use z-1, z+1)

Two modes:
    "exchange"   one torch.distributed send/recv (NCCL over NVLink) of the ghost rows before each stencil stage
    "recompute"  the noise stage also evaluates all ghost rows the rest of the chain will consume
                 (sum of the halos above); no communication at all.
Both give bit-identical results to the single-GPU chain on the owned rows.

The compute engine is injected: the product engine (CudaEngine) calls the device layer of the C ABI and
needs a GPU; the CPU tests drive the same orchestration over gloo with an oracle-backed engine.
"""
from dataclasses import dataclass, field

from . import device as _dev


@dataclass
class ChainConfig:
    """The C5 chain: simplex fBm -> Gauss5 x17 -> FlowMap x5 -> Value Erosion x5 -> mesh."""
    N: int = 16384
    noise_type: int = 3
    hurst: float = 0.4
    starting_amplitude: float = 1.0
    stepdown: float = 2.0
    detune_rate: float = 0.0
    octaves: int = 13
    noise_size: int = 1700
    xpos: int = 0
    zpos: int = 0
    filter_type: int = 2          # Gauss5_S1
    filter_radius: int = 2
    filter_iterations: int = 17
    flow_iterations: int = 5
    norm_min: float = 0.0
    norm_max: float = 0.005
    erosion_iterations: int = 5
    mesh_type: int = 1            # OvershootSquareGridHeightMap
    mesh_margin: int = 4          # mesh resolution R = N - 2*margin
    tile_height: float = 2000.0
    tile_size: float = field(default=0.0)

    def __post_init__(self):
        if self.tile_size == 0.0:
            self.tile_size = (self.N - 2 * self.mesh_margin) * (500.0 / 256.0)

    @property
    def R(self):
        return self.N - 2 * self.mesh_margin

    def halos(self):
        """(above, below) ghost rows each stage consumes."""
        g = self.filter_radius * self.filter_iterations
        f = 2 * self.flow_iterations + 1
        return {"filter": (g, g), "flow": (f, f), "erosion": (self.erosion_iterations, 0), "mesh": (1, 1)}


def band_rows(N, world, rank):
    return rank * N // world, (rank + 1) * N // world


class CudaEngine:
    """Product engine: the device layer of the C ABI on torch CUDA tensors.  Needs a GPU; no fallback."""
    name = "cuda"

    def __init__(self):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("CudaEngine needs a CUDA device (noize_b200 has no CPU path)")
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device())

    def empty(self, rows, width):
        return self.torch.empty(rows, width, dtype=self.torch.float32, device=self.device)

    def empty_bytes(self, n):
        return self.torch.empty(n, dtype=self.torch.uint8, device=self.device)

    def empty_mesh(self, nvert, nidx):
        t = self.torch
        return (t.empty(nvert, 12, dtype=t.float32, device=self.device), t.empty(nidx, dtype=t.int32, device=self.device))

    def fractal(self, dst, cfg, z_first):
        _dev.fractal(dst, cfg.noise_type, cfg.hurst, cfg.starting_amplitude, cfg.stepdown, cfg.detune_rate, cfg.octaves,
                     cfg.xpos, cfg.zpos, cfg.noise_size, z_first=z_first)

    def kernel_filter(self, data, tmp, cfg):
        return _dev.kernel_filter(data, tmp, cfg.filter_type, cfg.filter_iterations)

    def flowmap_scratch_bytes(self, width, rows, cfg):
        return _dev.flowmap_scratch_bytes(width, rows, cfg.flow_iterations)

    def flowmap(self, height, tmp, scratch, cfg):
        return _dev.flowmap(height, tmp, scratch, cfg.flow_iterations, cfg.norm_min, cfg.norm_max)

    def min_erosion(self, data, tmp, cfg):
        return _dev.min_erosion(data, tmp, cfg.erosion_iterations)

    def mesh(self, vtx, idx, heights, h_row_first, vz0, vz1, cfg):
        _dev.heightmap_mesh(cfg.mesh_type, vtx, idx, cfg.R, cfg.N, cfg.mesh_margin, cfg.tile_height, cfg.tile_size,
                            heights, h_row_first=h_row_first, vz_begin=vz0, vz_end=vz1)


class BandChain:
    """Runs ChainConfig on this rank's row band.  `dist` is torch.distributed (already initialised) or None
    for a single band."""

    def __init__(self, cfg, engine, rank=0, world=1, dist=None, mode="exchange", with_mesh=True):
        if mode not in ("exchange", "recompute"):
            raise ValueError("mode must be 'exchange' or 'recompute'")
        if world > 1 and dist is None and mode == "exchange":
            raise ValueError("exchange mode with world > 1 needs torch.distributed")
        self.cfg, self.eng, self.rank, self.world, self.dist, self.mode = cfg, engine, rank, world, dist, mode
        self.with_mesh = with_mesh
        N = cfg.N
        self.z0, self.z1 = band_rows(N, world, rank)
        self.own = self.z1 - self.z0
        h = cfg.halos()
        self.h = h
        order = ["filter", "flow", "erosion"] + (["mesh"] if with_mesh else [])
        if mode == "exchange":
            self.cap_above = max(h[s][0] for s in order)
            self.cap_below = max(h[s][1] for s in order)
        else:
            self.cap_above = sum(h[s][0] for s in order)
            self.cap_below = sum(h[s][1] for s in order)
        if world > 1 and self.own < max(self.cap_above, self.cap_below):
            raise ValueError(f"band of {self.own} rows is smaller than the ghost zone ({self.cap_above}/{self.cap_below})")
        # ghost rows that exist (the full grid's border has none: clamp-to-edge applies there)
        self.above = min(self.cap_above, self.z0)
        self.below = min(self.cap_below, N - self.z1)
        rows = self.above + self.own + self.below
        self.buf_a = engine.empty(rows, N)
        self.buf_b = engine.empty(rows, N)
        self.first_row = self.z0 - self.above          # global row of buffer row 0
        # vertex rows whose height row (vz + off) this band owns
        off = (N - cfg.R) // 2
        self.vz0 = max(self.z0 - off, 0) if rank > 0 else 0
        self.vz1 = min(self.z1 - off, cfg.R + 1) if rank < world - 1 else cfg.R + 1
        self.flow_scratch = None
        self.vtx = self.idx = None
        self.bytes_exchanged = 0
        self.runs = 0
        self.name = f"bands.BandChain over torch.distributed ({engine.name} engine)"

    # -- helpers ---------------------------------------------------------------------------------------
    def _win(self, buf, above, below):
        """View of `buf` covering own rows plus (above, below) ghost rows, clipped at the grid border."""
        a = min(above, self.z0)
        b = min(below, self.cfg.N - self.z1)
        lo = self.above - a
        return buf[lo: self.above + self.own + b], a, b

    def _exchange(self, buf, above, below):
        """Fill `above` ghost rows over and `below` ghost rows under the owned rows of `buf` from the
        neighbours' owned rows (and serve theirs).  Row counts are symmetric across ranks."""
        if self.world == 1 or (above == 0 and below == 0):
            return
        dist = self.dist
        ops = []
        top, bot = self.above, self.above + self.own   # owned rows are buf[top:bot]
        up, down = self.rank - 1, self.rank + 1
        if up >= 0:
            if below > 0:   # my first `below` rows are the upper neighbour's lower ghost rows
                ops.append(dist.P2POp(dist.isend, buf[top: top + below], up))
            if above > 0:
                ops.append(dist.P2POp(dist.irecv, buf[top - above: top], up))
        if down < self.world:
            if above > 0:   # my last `above` rows are the lower neighbour's upper ghost rows
                ops.append(dist.P2POp(dist.isend, buf[bot - above: bot], down))
            if below > 0:
                ops.append(dist.P2POp(dist.irecv, buf[bot: bot + below], down))
        for op in ops:
            self.bytes_exchanged += op.tensor.numel() * 4 if op.op is dist.irecv else 0
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    # -- the chain ---------------------------------------------------------------------------------------
    def run(self, stages=None):
        """Runs the chain; returns the buffer whose owned rows hold the final heightmap.  `stages`
        optionally receives (name) callbacks for per-stage timing: stages(name) is called before each."""
        cfg, eng = self.cfg, self.eng
        mark = stages if stages is not None else (lambda name: None)
        exch = self.mode == "exchange"
        order = ["filter", "flow", "erosion"] + (["mesh"] if self.with_mesh else [])
        # ghost rows still needed AFTER each stage (recompute mode shrinks the window stage by stage)
        rem_a = {s: sum(self.h[t][0] for t in order[i + 1:]) for i, s in enumerate(order)}
        rem_b = {s: sum(self.h[t][1] for t in order[i + 1:]) for i, s in enumerate(order)}

        cur, other = self.buf_a, self.buf_b
        mark("noise")
        if exch:
            w, a, _ = self._win(cur, 0, 0)
        else:
            w, a, _ = self._win(cur, self.cap_above, self.cap_below)
        eng.fractal(w, cfg, self.z0 - a)

        def stencil(name, fn):
            nonlocal cur, other
            mark(name)
            ha, hb = self.h[name]
            if exch:
                self._exchange(cur, ha, hb)
                above, below = ha, hb
            else:
                above, below = ha + rem_a[name], hb + rem_b[name]
            win, _, _ = self._win(cur, above, below)
            tmp, _, _ = self._win(other, above, below)
            res = fn(win, tmp)
            if res.data_ptr() == tmp.data_ptr():
                cur, other = other, cur

        stencil("filter", lambda win, tmp: eng.kernel_filter(win, tmp, cfg))

        def flow(win, tmp):
            need = eng.flowmap_scratch_bytes(win.shape[1], win.shape[0], cfg)
            if need and (self.flow_scratch is None or self.flow_scratch.numel() < need):
                self.flow_scratch = eng.empty_bytes(need)
            return eng.flowmap(win, tmp, self.flow_scratch if need else None, cfg)

        stencil("flow", flow)
        stencil("erosion", lambda win, tmp: eng.min_erosion(win, tmp, cfg))

        if self.with_mesh:
            mark("mesh")
            if exch:
                self._exchange(cur, 1, 1)
            win, a, _ = self._win(cur, 1, 1)
            R = cfg.R
            nv = (self.vz1 - self.vz0) * (R + 1)
            t0 = max(self.vz0, 1)
            ni = 6 * R * max(self.vz1 - t0, 0)
            if self.vtx is None:
                self.vtx, self.idx = eng.empty_mesh(nv, max(ni, 1))
            eng.mesh(self.vtx, self.idx, win, self.z0 - a, self.vz0, self.vz1, cfg)
        mark("end")
        self.result = cur
        self.runs += 1
        return cur

    STAGES = ("noise", "filter", "flow", "erosion", "mesh")

    def run_timed(self):
        """run() with a CUDA event before every stage (on the current torch stream); returns a callable that gives the
        per-stage milliseconds [noise, filter, flow, erosion, mesh] once the work has been synchronised."""
        import torch
        names = list(self.STAGES) + ["end"]
        evs = {n: torch.cuda.Event(enable_timing=True) for n in names}
        self.run(lambda name: evs[name].record())
        if not self.with_mesh:
            evs["mesh"] = evs["end"]
        return lambda: [evs[a].elapsed_time(evs[b]) for a, b in zip(names[:-1], names[1:])]

    def release(self):
        """Drop the device buffers (the object keeps its geometry)."""
        self.buf_a = self.buf_b = self.vtx = self.idx = self.flow_scratch = self.result = None

    def owned(self, buf=None):
        buf = self.result if buf is None else buf
        return buf[self.above: self.above + self.own]

def bind_host_to_gpu_numa_node(device_index):
    """Pin this process (and so the pages of every pinned host buffer it allocates afterwards: first touch) to the NUMA
    node the GPU hangs off.  With one process per GPU and 20 GB of mesh per step going device -> host, buffers that
    land on the other socket make every rank's download cross the inter-socket link.  Returns a short description of
    what was done ('' when the topology cannot be read: the process is left alone)."""
    import glob
    import os
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        paths = glob.glob(f"/sys/bus/pci/devices/{bus}/numa_node")
        if not paths:
            return ""
        node = int(open(paths[0]).read().strip())
        if node < 0:
            return ""
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return ""
        os.sched_setaffinity(0, cpus)
        return f"gpu {device_index} ({bus}) -> numa node {node}, {len(cpus)} cpus"
    except Exception:                      # unreadable sysfs, no such attribute: not an error for the caller
        return ""


class LibBandChain:
    """The same chain driven INSIDE libnoize_b200.so (nz_band_chain_*): this class only passes the configuration through the
    C ABI and wraps the device pointers it gets back.  One band per process (`dist` = torch.distributed, used ONCE to hand
    rank 0's NCCL unique id to the other ranks; the ghost rows then travel by ncclSend/ncclRecv issued by the library on
    the caller's stream), or, with `devices`, every band in this process (peer copies)."""

    STAGES = BandChain.STAGES

    def __init__(self, cfg, rank=0, world=1, dist=None, mode="exchange", with_mesh=True, devices=None):
        import ctypes as C
        import torch
        from . import lib as _l
        if mode not in ("exchange", "recompute"):
            raise ValueError("mode must be 'exchange' or 'recompute'")
        self.torch, self._l, self._C = torch, _l, C
        self.cfg, self.rank, self.world, self.mode, self.with_mesh = cfg, rank, world, mode, with_mesh
        lib = _l.load()
        c = _l.ChainConfig(cfg.N, cfg.noise_type, cfg.hurst, cfg.starting_amplitude, cfg.stepdown, cfg.detune_rate, cfg.octaves,
                           cfg.xpos, cfg.zpos, cfg.noise_size, cfg.filter_type, cfg.filter_iterations, cfg.flow_iterations,
                           cfg.norm_min, cfg.norm_max, cfg.erosion_iterations, cfg.mesh_type, cfg.R if with_mesh else 0,
                           cfg.mesh_margin, cfg.tile_height, cfg.tile_size)
        m = 0 if mode == "exchange" else 1
        self.comm = 0
        if devices is not None:
            devs = (C.c_int32 * len(devices))(*devices)
            self.world = len(devices)
            self.handle = int(lib.nz_band_chain_create_local(C.byref(c), devs, len(devices), m))
            self.name = f"nz_band_chain (C ABI), {len(devices)} bands in one process, peer copies"
        else:
            device = torch.cuda.current_device()
            if world > 1:
                # the communicator also tells the band its place (rank, world); recompute mode never transfers through it
                if dist is None:
                    raise ValueError("world > 1 needs torch.distributed to hand out the NCCL unique id")
                ident = torch.zeros(128, dtype=torch.uint8)
                if rank == 0:
                    buf = C.create_string_buffer(128)
                    _l.check(lib.nz_comm_unique_id(buf, 128))
                    ident = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
                ident = ident.cuda()
                dist.broadcast(ident, 0)
                raw = bytes(ident.cpu().numpy().tobytes())
                self.comm = int(lib.nz_comm_create(raw, world, rank, device))
                if self.comm < 0:
                    _l.check(self.comm)
            self.handle = int(lib.nz_band_chain_create(C.byref(c), self.comm, device, m, _l.stream_ptr()))
            self.name = ("nz_band_chain (C ABI), one band per process, ncclSend/ncclRecv issued by the library"
                         if world > 1 else "nz_band_chain (C ABI), single band")
        if self.handle < 0:
            _l.check(self.handle)
        self.nlocal = int(lib.nz_band_chain_local_bands(self.handle))
        self.runs = 0
        self._info = None
        i = self.info(0)
        self.z0, self.z1, self.own, self.vz0, self.vz1 = i.z0, i.z1, i.z1 - i.z0, i.vz0, i.vz1

    def info(self, local_band=0):
        i = self._l.BandInfo()
        self._l.check(self._l.load().nz_band_chain_info(self.handle, local_band, self._C.byref(i)))
        return i

    def run(self, stages=None):
        self._l.check(self._l.load().nz_band_chain_run(self.handle, 0))
        self.runs += 1
        self._info = None

    def run_timed(self):
        lib, C = self._l.load(), self._C
        self._l.check(lib.nz_band_chain_run(self.handle, 1))
        self.runs += 1
        self._info = None
        handle = self.handle

        def result():
            ms = (C.c_float * 5)()
            self._l.check(lib.nz_band_chain_stage_ms(handle, ms))
            return list(ms)
        return result

    def sync(self):
        self._l.check(self._l.load().nz_band_chain_sync(self.handle))

    @property
    def bytes_exchanged(self):
        return sum(self.info(k).halo_bytes_per_run for k in range(self.nlocal)) * max(1, self.runs)

    def _wrap(self, ptr, nbytes, dtype, device):
        """torch view of library-owned device memory (valid until the next run / destroy)."""
        torch = self.torch
        if nbytes == 0 or not ptr:
            return torch.empty(0, dtype=dtype, device=f"cuda:{device}")

        class _Mem:
            __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2, "strides": None}
        return torch.as_tensor(_Mem(), device=f"cuda:{device}").view(dtype)

    def owned(self, local_band=0):
        i = self.info(local_band)
        N = self.cfg.N
        return self._wrap(i.d_rows, (i.z1 - i.z0) * N * 4, self.torch.float32, i.device).view(i.z1 - i.z0, N)

    def mesh_slice(self, local_band=0):
        i = self.info(local_band)
        R = self.cfg.R
        nv = (i.vz1 - i.vz0) * (R + 1)
        ni = 6 * R * max(i.vz1 - max(i.vz0, 1), 0)
        return (self._wrap(i.d_vertices, nv * 48, self.torch.float32, i.device).view(nv, 12),
                self._wrap(i.d_indices, ni * 4, self.torch.int32, i.device))

    @property
    def vtx(self):
        return self.mesh_slice(0)[0]

    @property
    def idx(self):
        return self.mesh_slice(0)[1]

    def download(self, heights=None, vertices=None, indices=None):
        """Full-grid host arrays (numpy, C-contiguous); every local band writes its own slice."""
        p = lambda a: None if a is None else a.ctypes.data
        self._l.check(self._l.load().nz_band_chain_download(self.handle, p(heights), p(vertices), p(indices)))

    def release(self):
        if self.handle:
            self._l.check(self._l.load().nz_band_chain_destroy(self.handle))
            self.handle = 0
        if self.comm:
            self._l.check(self._l.load().nz_comm_destroy(self.comm))
            self.comm = 0

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def lib_band_geometry(N, world, rank, mesh_resolution=0):
    """nz_band_geometry: (z0, z1, vz0, vz1) of band `rank` of `world` as the LIBRARY partitions the grid (no GPU needed)."""
    import ctypes as C
    from . import lib as _l
    i = _l.BandInfo()
    _l.check(_l.load().nz_band_geometry(N, world, rank, mesh_resolution, C.byref(i)))
    return i.z0, i.z1, i.vz0, i.vz1
