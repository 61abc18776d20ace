"""Row bands on real GPUs: the banded chain must equal the single-band chain bit for bit.

* one GPU is enough for the `recompute` mode (no communication): the ranks of a virtual world run one after
  another on cuda:0;
* the `exchange` mode (NCCL halo exchange) runs when the box has >= 2 GPUs, one process per GPU.
"""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cfg(bands, N):
    return bands.ChainConfig(N=N, noise_size=170, filter_iterations=6, flow_iterations=3, erosion_iterations=4)


def _single(nz, bands, N):
    import torch
    chain = bands.BandChain(_cfg(bands, N), bands.CudaEngine())
    chain.run()
    torch.cuda.synchronize()
    return chain.owned().clone(), chain.vtx.clone(), chain.idx.clone()


@pytest.mark.parametrize("world", [2, 3])
def test_recompute_bands_equal_single_grid_bitwise(nz, world):
    import torch
    from noize_job_b200 import bands
    N = 512
    ref, rv, ri = _single(nz, bands, N)
    R = N - 8
    for rank in range(world):
        chain = bands.BandChain(_cfg(bands, N), bands.CudaEngine(), rank, world, None, mode="recompute")
        chain.run()
        torch.cuda.synchronize()
        assert torch.equal(chain.owned(), ref[chain.z0:chain.z1]), f"rank {rank}/{world}: heightmap band differs"
        assert torch.equal(chain.vtx, rv[chain.vz0 * (R + 1):chain.vz1 * (R + 1)])
        t0 = max(chain.vz0, 1)
        n = 6 * R * (chain.vz1 - t0)
        assert torch.equal(chain.idx[:n], ri[6 * R * (t0 - 1):6 * R * (chain.vz1 - 1)])


def test_single_band_chain_matches_oracle(nz, oracle):
    import torch
    from noize_job_b200 import bands
    N = 256
    cfg = _cfg(bands, N)
    got, _, _ = _single(nz, bands, N)
    ref = oracle.fractal(N, N, 3, 0.4, octaves=13, noise_size=170)
    ref = oracle.min_erosion(oracle.flowmap(oracle.kernel_filter(ref, 2, 6), 3, 0.0, 0.005), 4)
    assert np.abs(got.cpu().numpy() - ref).max() <= 5e-5 * max(1.0, np.abs(ref).max())


def _nccl_worker(rank, world, port, N, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    import noize_job_b200 as nz
    from noize_job_b200 import bands
    nz.host.init(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ref, rv, ri = _single(nz, bands, N)
        chain = bands.BandChain(_cfg(bands, N), bands.CudaEngine(), rank, world, dist, mode="exchange")
        chain.run()
        torch.cuda.synchronize()
        R = N - 8
        assert torch.equal(chain.owned(), ref[chain.z0:chain.z1]), f"rank {rank}: heightmap band differs"
        assert torch.equal(chain.vtx, rv[chain.vz0 * (R + 1):chain.vz1 * (R + 1)])
        assert chain.bytes_exchanged > 0
        open(os.path.join(out_dir, f"ok_{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_nccl_halo_exchange_bands_equal_single_grid_bitwise(tmp_path):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_nccl_worker, args=(world, port, 512, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok_{r}").exists() for r in range(world))


def test_tile_world_on_gpu_matches_oracle_tile_by_tile(nz, oracle):
    """BASELINE config C4 (reduced): tiles sharded over 2 virtual ranks, 3 tiles in flight per rank on separate streams."""
    from noize_job_b200 import tiles
    cfg = tiles.TileWorldConfig(tiles_x=3, tiles_z=3, resolution=256, tile_resolution=250, octaves=6, noise_size=400)
    seen = {}

    def consume(tx, tz, h, e, v, i):
        seen[(tx, tz)] = (h.cpu().numpy().copy(), e.cpu().numpy().copy(), v.cpu().numpy().copy(), i.cpu().numpy().view(np.uint32).copy())

    for rank in range(2):
        tw = tiles.TileWorld(cfg, tiles.TileCudaEngine(3), rank, 2, slots=3)
        assert tw.run(consume) == len(tw.mine)
    assert sorted(seen) == sorted(cfg.tiles())
    for (tx, tz), (h, e, v, i) in seen.items():
        ref = oracle.kernel_filter(oracle.fractal(256, 256, 4, 0.4, octaves=6, xpos=250 * tx, zpos=250 * tz, noise_size=400), 3, 3)
        assert np.abs(h - ref).max() <= 1e-6
        assert np.abs(e - oracle.kernel_filter(ref, 11, 1)).max() <= 2e-6
        rv, ri = oracle.heightmap_mesh(1, h, cfg.R, 4, cfg.tile_height, cfg.tile_size)
        assert np.array_equal(i, ri) and np.abs(v - rv).max() <= 1e-5 * cfg.tile_height


def test_library_tile_world_matches_oracle_tile_by_tile(nz, oracle):
    """nz_tile_world_*: the C4 tile loop inside the library (3 tiles in flight), outputs downloaded per tile."""
    from noize_job_b200 import tiles
    cfg = tiles.TileWorldConfig(tiles_x=3, tiles_z=2, resolution=256, tile_resolution=250, octaves=6, noise_size=400)
    R = cfg.R
    for rank in range(2):
        tw = tiles.LibTileWorld(cfg, rank, 2, slots=3)
        n = len(tw.mine)
        h = np.zeros((n, 256, 256), np.float32)
        e = np.zeros((n, 256, 256), np.float32)
        v = np.zeros((n, (R + 1) * (R + 1), 12), np.float32)
        i = np.zeros((n, 6 * R * R), np.uint32)
        assert tw.run(h, e, v, i) == n
        for k, (tx, tz) in enumerate(tw.mine):
            ref = oracle.kernel_filter(oracle.fractal(256, 256, 4, 0.4, octaves=6, xpos=250 * tx, zpos=250 * tz, noise_size=400), 3, 3)
            assert np.abs(h[k] - ref).max() <= 1e-6
            assert np.abs(e[k] - oracle.kernel_filter(ref, 11, 1)).max() <= 2e-6
            rv, ri = oracle.heightmap_mesh(1, h[k], R, 4, cfg.tile_height, cfg.tile_size)
            assert np.array_equal(i[k], ri) and np.abs(v[k] - rv).max() <= 1e-5 * cfg.tile_height
        tw.run()                                   # device-resident form (what the bench times)
        tw.release()
