"""Tile sharding (BASELINE config C4) on CPU: world_size 2 over gloo with an oracle-backed engine.  Under test is
noize_job_b200.tiles.TileWorld (ownership, slot reuse, tile -> noise-domain mapping), not the arithmetic."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class OracleTileEngine:
    name = "oracle"

    def __init__(self, nslots):
        import oracle
        self.o = oracle.get()

    def alloc(self, cfg):
        return dict(vtx=None, idx=None)

    def run_tile(self, slot, buf, cfg, tx, tz):
        n = cfg.resolution
        h = self.o.fractal(n, n, cfg.noise_type, cfg.hurst, octaves=cfg.octaves, xpos=cfg.tile_resolution * tx,
                           zpos=cfg.tile_resolution * tz, noise_size=cfg.noise_size)
        h = self.o.kernel_filter(h, cfg.filter_type, cfg.filter_iterations)
        e = self.o.kernel_filter(h, cfg.edge_filter_type, 1)
        buf["vtx"], buf["idx"] = self.o.heightmap_mesh(cfg.mesh_type, h, cfg.R, cfg.mesh_margin, cfg.tile_height, cfg.tile_size)
        return h, e

    def wait(self, slot):
        pass


def small_cfg():
    import noize_job_b200 as nz
    from noize_job_b200 import tiles
    return tiles.TileWorldConfig(tiles_x=3, tiles_z=2, resolution=40, tile_resolution=32, octaves=3, noise_size=60)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from noize_job_b200 import tiles
    cfg = small_cfg()
    got = {}
    tw = tiles.TileWorld(cfg, OracleTileEngine(2), rank, world, slots=2)
    n = tw.run(lambda tx, tz, h, e, v, i: got.__setitem__((tx, tz), (h.copy(), e.copy(), v.copy(), i.copy())))
    assert n == len(tw.mine) == len(got)
    # gather which tiles every rank produced: each tile exactly once
    owned = [None] * world
    dist.all_gather_object(owned, sorted(got))
    if rank == 0:
        out.put((owned, {k: v[0] for k, v in got.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_tiles_are_sharded_once_and_match_single_process():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    [p.start() for p in procs]
    owned, rank0 = out.get()
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    from noize_job_b200 import tiles
    cfg = small_cfg()
    every = sorted(t for part in owned for t in part)
    assert every == sorted(cfg.tiles())                                    # a partition of the world
    assert all(tiles.tile_owner(tx, tz, cfg, 2) == r for r, part in enumerate(owned) for tx, tz in part)
    # single process, one slot: same tiles, same bits; and the tile -> noise-domain mapping of MeshTileGenerator.cs:188-189
    single = {}
    tiles.TileWorld(cfg, OracleTileEngine(1), 0, 1, slots=1).run(lambda tx, tz, h, e, v, i: single.__setitem__((tx, tz), h.copy()))
    for k, h in rank0.items():
        assert np.array_equal(h, single[k])
    import oracle
    o = oracle.get()
    ref = o.kernel_filter(o.fractal(40, 40, 4, 0.4, octaves=3, xpos=32 * 2, zpos=32 * 1, noise_size=60), 3, 3)
    assert np.array_equal(single[(2, 1)], ref)
    # neighbouring tiles overlap by resolution - tile_resolution cells of the SAME noise (before filtering)
    a = o.fractal(40, 40, 4, 0.4, octaves=3, xpos=0, zpos=0, noise_size=60)
    b = o.fractal(40, 40, 4, 0.4, octaves=3, xpos=32, zpos=0, noise_size=60)
    assert np.array_equal(a[:, 32:], b[:, :8])
