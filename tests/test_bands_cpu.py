"""Multi-rank row-band orchestration on CPU: world_size 2 and 3 over gloo, compute injected from the oracle.

What is under test is noize_job_b200.bands.BandChain (ghost-row bookkeeping, halo exchange, vertex-row
partition), NOT the arithmetic: the OracleEngine below is test infrastructure standing in for the CUDA
engine, which needs a GPU.  The banded result must equal the single-grid result bit for bit.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class OracleEngine:
    """CPU stand-in for bands.CudaEngine (same interface), backed by oracle/."""
    name = "oracle"

    def __init__(self):
        import oracle
        self.o = oracle.get()

    def empty(self, rows, width):
        return torch.full((rows, width), float("nan"), dtype=torch.float32)

    def empty_bytes(self, n):
        return torch.empty(max(n, 1), dtype=torch.uint8)

    def empty_mesh(self, nvert, nidx):
        return torch.zeros(nvert, 12, dtype=torch.float32), torch.zeros(nidx, dtype=torch.int32)

    def fractal(self, dst, cfg, z_first):
        rows, width = dst.shape
        dst.copy_(torch.from_numpy(self.o.fractal(width, rows, cfg.noise_type, cfg.hurst, cfg.starting_amplitude,
                                                  cfg.stepdown, cfg.detune_rate, cfg.octaves, cfg.xpos, cfg.zpos,
                                                  cfg.noise_size, z_first=z_first)))

    def kernel_filter(self, data, tmp, cfg):
        tmp.copy_(torch.from_numpy(self.o.kernel_filter(data.numpy(), cfg.filter_type, cfg.filter_iterations)))
        return tmp  # exercises the "result landed in the other buffer" path

    def flowmap_scratch_bytes(self, width, rows, cfg):
        return 16

    def flowmap(self, height, tmp, scratch, cfg):
        tmp.copy_(torch.from_numpy(self.o.flowmap(height.numpy(), cfg.flow_iterations, cfg.norm_min, cfg.norm_max)))
        return tmp

    def min_erosion(self, data, tmp, cfg):
        data.copy_(torch.from_numpy(self.o.min_erosion(data.numpy(), cfg.erosion_iterations)))
        return data

    def mesh(self, vtx, idx, heights, h_row_first, vz0, vz1, cfg):
        # the oracle meshes whole grids: embed the resident rows in a NaN grid, mesh it, keep this band's rows
        full = np.full((cfg.N, cfg.N), np.nan, np.float32)
        full[h_row_first:h_row_first + heights.shape[0]] = heights.numpy()
        v, i = self.o.heightmap_mesh(cfg.mesh_type, full, cfg.R, cfg.mesh_margin, cfg.tile_height, cfg.tile_size)
        R = cfg.R
        vtx.copy_(torch.from_numpy(v[vz0 * (R + 1):vz1 * (R + 1)]))
        t0 = max(vz0, 1)
        sl = i[6 * R * (t0 - 1):6 * R * (vz1 - 1)].astype(np.int32)
        idx[:sl.size].copy_(torch.from_numpy(sl))


def _reference(cfg):
    import oracle
    o = oracle.get()
    g = o.fractal(cfg.N, cfg.N, cfg.noise_type, cfg.hurst, cfg.starting_amplitude, cfg.stepdown, cfg.detune_rate,
                  cfg.octaves, cfg.xpos, cfg.zpos, cfg.noise_size)
    g = o.kernel_filter(g, cfg.filter_type, cfg.filter_iterations)
    g = o.flowmap(g, cfg.flow_iterations, cfg.norm_min, cfg.norm_max)
    g = o.min_erosion(g, cfg.erosion_iterations)
    v, i = o.heightmap_mesh(cfg.mesh_type, g, cfg.R, cfg.mesh_margin, cfg.tile_height, cfg.tile_size)
    return g, v, i


def _worker(rank, world, port, mode, N, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import noize_job_b200 as nz
        from noize_job_b200 import bands
        cfg = bands.ChainConfig(N=N, noise_size=170, filter_iterations=6, flow_iterations=3, erosion_iterations=4)
        chain = bands.BandChain(cfg, OracleEngine(), rank, world, dist, mode=mode)
        chain.run()
        ref, rv, ri = _reference(cfg)
        own = chain.owned().numpy()
        assert not np.isnan(own).any()
        assert np.array_equal(own, ref[chain.z0:chain.z1]), f"rank {rank}: heightmap band differs"
        R = cfg.R
        assert np.array_equal(chain.vtx.numpy(), rv[chain.vz0 * (R + 1):chain.vz1 * (R + 1)]), f"rank {rank}: vertices differ"
        t0 = max(chain.vz0, 1)
        want = ri[6 * R * (t0 - 1):6 * R * (chain.vz1 - 1)].astype(np.int32)
        assert np.array_equal(chain.idx.numpy()[:want.size], want), f"rank {rank}: indices differ"
        if mode == "exchange" and world > 1:
            assert chain.bytes_exchanged > 0
        else:
            assert chain.bytes_exchanged == 0
        # vertex rows of all ranks tile [0, R+1) without gaps
        spans = [None] * world
        dist.all_gather_object(spans, (chain.vz0, chain.vz1))
        assert spans[0][0] == 0 and spans[-1][1] == R + 1
        assert all(spans[k][1] == spans[k + 1][0] for k in range(world - 1))
        open(os.path.join(out_dir, f"ok_{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,mode", [(2, "exchange"), (2, "recompute"), (3, "exchange")])
def test_banded_chain_equals_single_grid_bitwise(tmp_path, world, mode):
    N = 192 if world == 3 else 160
    mp.spawn(_worker, args=(world, _free_port(), mode, N, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok_{r}").exists() for r in range(world))


def test_single_band_needs_no_process_group():
    from noize_job_b200 import bands
    cfg = bands.ChainConfig(N=96, noise_size=170, filter_iterations=3, flow_iterations=2, erosion_iterations=2)
    chain = bands.BandChain(cfg, OracleEngine())
    chain.run()
    ref, rv, ri = _reference(cfg)
    assert np.array_equal(chain.owned().numpy(), ref)
    assert np.array_equal(chain.vtx.numpy(), rv) and np.array_equal(chain.idx.numpy(), ri.astype(np.int32))


def test_band_plan_validation():
    from noize_job_b200 import bands
    cfg = bands.ChainConfig(N=64)
    with pytest.raises(ValueError):
        bands.BandChain(cfg, OracleEngine(), rank=0, world=4, dist=dist, mode="exchange")   # 16-row bands < 34 ghost rows
    with pytest.raises(ValueError):
        bands.BandChain(cfg, OracleEngine(), mode="bogus")
    assert cfg.halos() == {"filter": (34, 34), "flow": (11, 11), "erosion": (5, 0), "mesh": (1, 1)}
    assert [bands.band_rows(16384, 8, r) for r in (0, 7)] == [(0, 2048), (14336, 16384)]


def test_cuda_engine_fails_loudly_without_a_gpu():
    from noize_job_b200 import bands
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU path"):
        bands.CudaEngine()


def test_library_partition_equals_python_partition():
    """The in-library band chain (nz_band_chain_*) and bands.BandChain must cut the grid identically: owned heightmap rows
    and owned vertex rows, for every rank, including uneven splits and the margin rows of the first / last band."""
    from noize_job_b200 import bands

    class NoEngine:
        name = "none"

        def empty(self, rows, width):
            return None

    for N, margin in ((16384, 4), (1024, 4), (1000, 12), (257, 1)):
        cfg = bands.ChainConfig(N=N, mesh_margin=margin, filter_iterations=1, flow_iterations=1, erosion_iterations=1)
        for world in (1, 2, 3, 7, 8):
            covered, vcovered = [], []
            for rank in range(world):
                z0, z1, vz0, vz1 = bands.lib_band_geometry(N, world, rank, cfg.R)
                py = bands.BandChain(cfg, NoEngine(), rank, world, None, mode="recompute")
                assert (z0, z1, vz0, vz1) == (py.z0, py.z1, py.vz0, py.vz1), (N, world, rank)
                covered.append((z0, z1))
                vcovered.append((vz0, vz1))
            assert covered[0][0] == 0 and covered[-1][1] == N and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
            assert vcovered[0][0] == 0 and vcovered[-1][1] == cfg.R + 1 and all(a[1] == b[0] for a, b in zip(vcovered, vcovered[1:]))
    import noize_job_b200 as nz
    import pytest
    with pytest.raises(nz.NzError):
        bands.lib_band_geometry(16, 32, 0)
