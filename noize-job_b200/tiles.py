"""Tile sharding of a tiled world across the GPUs of a box (BASELINE config C4).

The reference generates a world tile by tile: MeshTileGenerator hands tile (tx, tz) to the generator pipeline with
xpos = tileResolution * tx, zpos = tileResolution * tz (Scripts/MeshTileGenerator.cs:184-192), ONE tile in flight
(:125-138).  Tiles are independent (the noise is a pure function of position and every filter clamps at the tile's own
border), so sharding needs no communication: rank r of `world` owns the tiles with (tz * tiles_x + tx) % world == r,
runs the ordinary single-tile chain on each, and keeps `streams` tiles in flight on separate CUDA streams so that
1024^2 tiles, which cannot fill 148 SMs alone, overlap (tools/config_times.py: 32 tiles 4.64 ms serial, 3.73 ms on 4 streams).

The compute engine is injected exactly as in bands.py: TileCudaEngine calls the device layer of the C ABI; the CPU
tests drive the same orchestration over gloo with an oracle-backed engine.
"""
from dataclasses import dataclass

from . import device as _dev


@dataclass
class TileWorldConfig:
    """C4: rotated-simplex fBm -> Gauss3 x3 (GaussHF.asset) -> Sobel3_2D on a copy (Sobel2D.asset) -> mesh."""
    tiles_x: int = 16
    tiles_z: int = 16
    resolution: int = 1024            # generator resolution of one tile
    tile_resolution: int = 1000       # world cells between tile origins: tiles overlap by resolution - tile_resolution
    noise_type: int = 4
    hurst: float = 0.4
    octaves: int = 13
    noise_size: int = 1700
    filter_type: int = 3              # Gauss3_S1
    filter_iterations: int = 3
    edge_filter_type: int = 11        # Sobel3_2D
    mesh_type: int = 1
    mesh_margin: int = 4
    tile_height: float = 2000.0

    @property
    def R(self):
        return self.resolution - 2 * self.mesh_margin

    @property
    def tile_size(self):
        return self.R * (500.0 / 256.0)

    def tiles(self):
        return [(tx, tz) for tz in range(self.tiles_z) for tx in range(self.tiles_x)]


def tile_owner(tx, tz, cfg, world):
    return (tz * cfg.tiles_x + tx) % world


class TileCudaEngine:
    """Product engine: device layer of the C ABI on torch CUDA tensors, one stream per slot.  No fallback."""
    name = "cuda"

    def __init__(self, nslots):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("TileCudaEngine needs a CUDA device (noize_b200 has no CPU path)")
        self.torch = torch
        self.streams = [torch.cuda.Stream() for _ in range(nslots)]

    def alloc(self, cfg):
        t = self.torch
        n, R = cfg.resolution, cfg.R
        return dict(a=t.empty(n, n, device="cuda"), b=t.empty(n, n, device="cuda"), edge=t.empty(n, n, device="cuda"),
                    vtx=t.empty((R + 1) * (R + 1), 12, device="cuda"), idx=t.empty(6 * R * R, dtype=t.int32, device="cuda"))

    def run_tile(self, slot, buf, cfg, tx, tz):
        """Enqueues the chain of one tile on the slot's stream; returns (heights, edges) tensors of the slot."""
        s = self.streams[slot]
        with self.torch.cuda.stream(s):
            _dev.fractal(buf["a"], cfg.noise_type, cfg.hurst, octaves=cfg.octaves, xpos=cfg.tile_resolution * tx,
                         zpos=cfg.tile_resolution * tz, noise_size=cfg.noise_size, stream=s)
            cur = _dev.kernel_filter(buf["a"], buf["b"], cfg.filter_type, cfg.filter_iterations, stream=s)
            other = buf["b"] if cur is buf["a"] else buf["a"]
            buf["edge"].copy_(cur)
            edges = _dev.kernel_filter(buf["edge"], other, cfg.edge_filter_type, 1, stream=s)
            _dev.heightmap_mesh(cfg.mesh_type, buf["vtx"], buf["idx"], cfg.R, cfg.resolution, cfg.mesh_margin, cfg.tile_height,
                                cfg.tile_size, cur, stream=s)
        return cur, edges

    def wait(self, slot):
        self.streams[slot].synchronize()


class TileWorld:
    """Runs this rank's tiles of a TileWorldConfig.  `consume(tx, tz, heights, edges, vtx, idx)` is called for every
    finished tile (e.g. to download it); tiles of different slots are in flight concurrently."""

    def __init__(self, cfg, engine, rank=0, world=1, slots=4):
        self.cfg, self.eng, self.rank, self.world, self.slots = cfg, engine, rank, world, slots
        self.mine = [(tx, tz) for tx, tz in cfg.tiles() if tile_owner(tx, tz, cfg, world) == rank]
        self.bufs = [engine.alloc(cfg) for _ in range(slots)]

    def run(self, consume=None):
        pending = [None] * self.slots                    # (tx, tz, heights, edges) of the tile occupying each slot
        done = 0
        for k, (tx, tz) in enumerate(self.mine):
            slot = k % self.slots
            if pending[slot] is not None:                # the slot's buffers are reused: its previous tile must be finished
                self._finish(slot, pending[slot], consume)
                done += 1
            h, e = self.eng.run_tile(slot, self.bufs[slot], self.cfg, tx, tz)
            pending[slot] = (tx, tz, h, e)
        for slot in range(self.slots):
            if pending[slot] is not None:
                self._finish(slot, pending[slot], consume)
                done += 1
        return done

    def _finish(self, slot, item, consume):
        self.eng.wait(slot)
        if consume is not None:
            tx, tz, h, e = item
            consume(tx, tz, h, e, self.bufs[slot]["vtx"], self.bufs[slot]["idx"])


class LibTileWorld:
    """The same sharding with the tile loop INSIDE the library (nz_tile_world_*): one C call runs this rank's tiles, `slots`
    of them in flight on separate streams.  The Python loop of TileWorld costs ~80 us of host time per tile (five ctypes
    calls and a torch copy), which is what a 1024^2 tile takes on the GPU: on 8 GPUs the Python form is host-bound."""

    def __init__(self, cfg, rank=0, world=1, slots=4, device=None):
        import ctypes as C
        import torch
        from . import lib as _l
        self._l, self._C, self.cfg, self.rank, self.world, self.slots = _l, C, cfg, rank, world, slots
        self.mine = [(tx, tz) for tx, tz in cfg.tiles() if tile_owner(tx, tz, cfg, world) == rank]
        c = _l.TileConfig(cfg.resolution, cfg.tile_resolution, cfg.noise_type, cfg.hurst, 1.0, 2.0, 0.0, cfg.octaves, cfg.noise_size,
                          cfg.filter_type, cfg.filter_iterations, cfg.edge_filter_type, 1, cfg.mesh_type, cfg.R, cfg.mesh_margin,
                          cfg.tile_height, cfg.tile_size)
        dev = torch.cuda.current_device() if device is None else device
        self.handle = int(_l.load().nz_tile_world_create(C.byref(c), dev, slots))
        if self.handle < 0:
            _l.check(self.handle)
        flat = [v for t in self.mine for v in t]
        self._tiles = (C.c_int32 * max(1, len(flat)))(*flat)

    def run(self, heights=None, edges=None, vertices=None, indices=None):
        """Runs this rank's tiles; optional HOST arrays (numpy, C-contiguous, one entry per owned tile) receive the outputs."""
        p = lambda a: None if a is None else a.ctypes.data
        self._l.check(self._l.load().nz_tile_world_run(self.handle, self._tiles, len(self.mine), p(heights), p(edges), p(vertices), p(indices)))
        return len(self.mine)

    def release(self):
        if self.handle:
            self._l.check(self._l.load().nz_tile_world_destroy(self.handle))
            self.handle = 0

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass
