"""Row bands INSIDE the library (nz_band_chain_*, nz_set_bands): every banded result must equal the single-band result bit
for bit, through the C ABI.

* one GPU is enough for bands that share a device (devices = [0, 0, 0]: ghost rows move by device-to-device copies) — this is
  what the driver's one-GPU box runs;
* with >= 2 GPUs the same tests run on distinct devices (peer copies over NVLink), and the one-band-per-process form
  exchanges its ghost rows with ncclSend/ncclRecv issued by the library.
"""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cfg(N=1024, **kw):
    from noize_job_b200 import bands
    base = dict(N=N, noise_size=170, filter_iterations=6, flow_iterations=3, erosion_iterations=4)
    base.update(kw)
    return bands.ChainConfig(**base)


def _download(chain):
    cfg = chain.cfg
    R = cfg.R
    h = np.zeros((cfg.N, cfg.N), np.float32)
    v = np.zeros(((R + 1) * (R + 1), 12), np.float32)
    i = np.zeros(6 * R * R, np.uint32)
    chain.download(h, v, i)
    return h, v, i


def _run(devices=None, mode="exchange", **kw):
    from noize_job_b200 import bands
    chain = bands.LibBandChain(_cfg(**kw), mode=mode, devices=devices)
    chain.run()
    out = _download(chain)
    halo = chain.bytes_exchanged
    chain.release()
    return out, halo


def _devices(n):
    import torch
    g = torch.cuda.device_count()
    return [k % g for k in range(n)]


def test_single_band_lib_chain_equals_python_chain_bitwise(nz):
    import torch
    from noize_job_b200 import bands
    (h, v, i), _ = _run(None, N=512)
    chain = bands.BandChain(_cfg(512), bands.CudaEngine())
    chain.run()
    torch.cuda.synchronize()
    assert np.array_equal(h, chain.owned().cpu().numpy())
    assert np.array_equal(v.view(np.uint32), chain.vtx.cpu().numpy().view(np.uint32))
    assert np.array_equal(i, chain.idx.cpu().numpy().view(np.uint32))


@pytest.mark.parametrize("n_bands,mode", [(2, "exchange"), (3, "exchange"), (4, "exchange"), (3, "recompute")])
def test_bands_in_one_process_equal_single_band_bitwise(nz, n_bands, mode):
    (h1, v1, i1), _ = _run(None)
    (h, v, i), halo = _run(_devices(n_bands), mode)
    assert np.array_equal(h, h1), "heightmap differs from the single band"
    assert np.array_equal(v.view(np.uint32), v1.view(np.uint32)), "vertices differ"
    assert np.array_equal(i, i1), "indices differ"
    assert (halo > 0) == (mode == "exchange")


def test_band_chain_all_filter_kinds(nz):
    """Sobel3_2D (two-branch 3x3), a wide kernel (Gauss9: r = 4) and the square-grid mesh through the banded chain."""
    for kw in (dict(filter_type=11, filter_iterations=2, mesh_type=0), dict(filter_type=0, filter_iterations=3), dict(filter_type=8, filter_iterations=5, flow_iterations=0)):
        (h1, v1, i1), _ = _run(None, N=512, **kw)
        (h, v, i), _ = _run(_devices(3), N=512, **kw)
        assert np.array_equal(h, h1) and np.array_equal(v.view(np.uint32), v1.view(np.uint32)) and np.array_equal(i, i1), kw


def test_band_chain_rejects_bands_smaller_than_the_ghost_zone(nz):
    from noize_job_b200 import bands
    with pytest.raises(nz.NzError) as e:
        bands.LibBandChain(_cfg(N=64, filter_iterations=17), devices=_devices(4))
    assert e.value.code == nz.lib.NZ_E_INVALID and "ghost zone" in str(e.value)


def test_timed_run_reports_stage_times(nz):
    from noize_job_b200 import bands
    chain = bands.LibBandChain(_cfg(), devices=_devices(2))
    ms = chain.run_timed()()
    assert len(ms) == 5 and all(m >= 0.0 for m in ms) and sum(ms) > 0.0
    chain.release()


# ---- host layer on bands (nz_set_bands): what a single-process C# host calls ------------------------------------------
def _stage_chain(nz, N, thermal=False):
    """noise -> Gauss5 x6 -> [thermal erosion: no banded form, gathers] -> flow x3 -> erosion x4 -> x0.5 -> mesh, one scope."""
    R = N - 8
    data = np.zeros(N * N, np.float32)
    vtx = np.zeros(((R + 1) * (R + 1), 12), np.float32)
    idx = np.zeros(6 * R * R, np.uint32)
    with nz.host.pipeline():
        nz.host.fractal(data, N, 3, 0.4, 1.0, 2.0, 0.0, 13, 0, 0, 170)
        nz.host.kernel_filter(data, None, 2, N, 6)
        if thermal:
            nz.host.thermal_erosion(data, 45.0, 0.5, 0.75, 1, N)
        nz.host.flowmap(data, N, 3, 0.0, 0.005)
        nz.host.min_erosion(data, N, 4)
        nz.host.constant(data, None, 0, 0.5, N)
        nz.host.heightmap_mesh(1, vtx, idx, R, N, 4, 2000.0, R * (500.0 / 256.0), data)
    return data, vtx, idx


@pytest.mark.parametrize("n_bands,thermal", [(2, False), (3, False), (2, True)])
def test_host_layer_on_bands_equals_one_device_bitwise(nz, n_bands, thermal):
    import torch
    N = 4096                                   # NZ_BANDS_MIN_RESOLUTION: smaller grids are never banded
    dev0 = torch.cuda.current_device()
    try:
        nz.host.init([dev0])
        nz.host.set_bands(0)
        ref = _stage_chain(nz, N, thermal)
        l0 = nz.host.kernel_launch_count()
        _stage_chain(nz, N, thermal)
        per_chain = nz.host.kernel_launch_count() - l0
        nz.host.init(_devices(n_bands))
        nz.host.set_bands(n_bands)
        l0 = nz.host.kernel_launch_count()
        got = _stage_chain(nz, N, thermal)
        assert nz.host.kernel_launch_count() - l0 > per_chain, "the banded path launches every stage once per band"
        for a, b, name in zip(got, ref, ("heightmap", "vertices", "indices")):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"{name} differs between {n_bands} bands and one device"
        # outside a scope every call brings its result home, banded or not
        a = np.zeros(N * N, np.float32)
        nz.host.fractal(a, N, 3, 0.4, 1.0, 2.0, 0.0, 13, 0, 0, 170)
        nz.host.kernel_filter(a, None, 2, N, 6)
        nz.host.set_bands(0)
        b = np.zeros(N * N, np.float32)
        nz.host.fractal(b, N, 3, 0.4, 1.0, 2.0, 0.0, 13, 0, 0, 170)
        nz.host.kernel_filter(b, None, 2, N, 6)
        assert np.array_equal(a, b)
    finally:
        nz.host.set_bands(0)
        nz.host.init([dev0])


def test_set_bands_validates(nz):
    import torch
    nz.host.init([torch.cuda.current_device()])
    with pytest.raises(nz.NzError) as e:
        nz.host.set_bands(2)
    assert e.value.code == nz.lib.NZ_E_INVALID
    nz.host.set_bands(1)
    nz.host.set_bands(0)


def test_failed_call_leaves_no_stale_mirror(nz):
    """a failed allocation inside an unscoped call must not leave a device mirror that the retry would trust."""
    N = 256
    a = np.random.default_rng(1).random(N * N, dtype=np.float32)
    want = a.copy()
    nz.host.kernel_filter(want, None, 2, N, 3)
    for skip in (0, 1):                       # the mirror itself fails, then its ping-pong partner (after the upload)
        b = a.copy()
        nz.host.test_fail_allocs(skip, 1)
        with pytest.raises(nz.NzError) as e:
            nz.host.kernel_filter(b, None, 2, N, 3)
        assert e.value.code == nz.lib.NZ_E_NOMEM
        nz.host.test_fail_allocs(0, 0)
        assert np.array_equal(b, a), "a failed call must leave the host slice untouched"
        b[:] = a * 0.5 + 0.25                 # the host data CHANGES before the retry: a stale mirror would ignore that
        want2 = b.copy()
        nz.host.kernel_filter(b, None, 2, N, 3)
        nz.host.kernel_filter(want2, None, 2, N, 3)
        assert np.array_equal(b, want2)


def test_reinit_inside_a_scope_is_refused(nz):
    import torch
    dev = torch.cuda.current_device()
    nz.host.pipeline_begin()
    try:
        with pytest.raises(nz.NzError) as e:
            nz.host.init([dev])
        assert e.value.code == nz.lib.NZ_E_STATE
    finally:
        nz.host.pipeline_end()
    nz.host.init([dev])


def test_device_layer_keeps_the_callers_device(nz, oracle):
    """nz_dev_* must run on the caller's current device and never switch it."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    nz.host.init([0])
    with torch.cuda.device(1):
        t = torch.empty(64, 256, device="cuda:1")
        nz.device.fractal(t, 3, 0.4, octaves=5, noise_size=170)
        assert torch.cuda.current_device() == 1
        got = t.cpu().numpy()
        other = torch.empty(64, 256, device="cuda:0")
        with pytest.raises(nz.NzError) as e:        # a buffer of another device is refused, not launched on
            nz.device.fractal(other, 3, 0.4, octaves=5, noise_size=170)
        assert e.value.code == nz.lib.NZ_E_INVALID
    assert np.abs(got - oracle.fractal(256, 64, 3, 0.4, octaves=5, noise_size=170)).max() <= 1e-6


# ---- one band per process, NCCL inside the library -------------------------------------------------------------------
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
        one = bands.LibBandChain(_cfg(N))
        one.run()
        ref, (rv, ri) = one.owned().clone(), [t.clone() for t in one.mesh_slice()]
        one.release()
        for mode in ("exchange", "recompute"):
            chain = bands.LibBandChain(_cfg(N), rank, world, dist, mode=mode)
            for _ in range(2):
                chain.run()
            torch.cuda.synchronize()
            R = N - 8
            assert torch.equal(chain.owned(), ref[chain.z0:chain.z1]), f"rank {rank} {mode}: heightmap band differs"
            v, i = chain.mesh_slice()
            assert torch.equal(v.view(torch.int32), rv[chain.vz0 * (R + 1):chain.vz1 * (R + 1)].view(torch.int32)), f"rank {rank} {mode}: vertices differ"
            t0 = max(chain.vz0, 1)
            assert torch.equal(i, ri[6 * R * (t0 - 1):6 * R * (chain.vz1 - 1)]), f"rank {rank} {mode}: indices differ"
            assert (chain.bytes_exchanged > 0) == (mode == "exchange")
            assert nz.load().nz_comm_async_error(chain.comm) == 0
            chain.release()
        open(os.path.join(out_dir, f"ok_{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_library_nccl_halo_exchange_equals_single_band_bitwise(tmp_path):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_nccl_worker, args=(world, port, 1024, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok_{r}").exists() for r in range(world))


def test_large_bands_with_edge_free_interior_launches_equal_single_band_bitwise(nz):
    """Bands large enough for the register-walk kernels (filter >= 8M cells, flow map > 4M cells per window): a band whose
    window edge is not a grid edge lets the clamp-free interior launch run over it (grid_edges() hint in nz_common.cuh).
    Three bands of a 6144^2 grid at the C5 iteration counts: the middle band has no grid edge at all."""
    kw = dict(N=6144, noise_size=1700, filter_iterations=17, flow_iterations=5, erosion_iterations=5)
    (h1, v1, i1), _ = _run(None, **kw)
    (h, v, i), halo = _run(_devices(3), **kw)
    assert np.array_equal(h, h1), "heightmap differs from the single band"
    assert np.array_equal(v.view(np.uint32), v1.view(np.uint32)) and np.array_equal(i, i1)
    (h, v, i), _ = _run(_devices(2), "recompute", **kw)
    assert np.array_equal(h, h1) and np.array_equal(v.view(np.uint32), v1.view(np.uint32)) and np.array_equal(i, i1)
