"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle.

Tolerances: noise <= 1e-6 abs, separable filter / flow map <= 1e-6 abs on the
normalised output, value erosion bit-exact, mesh indices bit-exact, vertices/normals <= 1e-5 abs.
Band / tiling invariance is bit-exact (same kernels, different window).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_NOISE = 1e-6
TOL_FILTER = 1e-6
TOL_FLOW = 1e-6
TOL_MESH = 1e-5

C1 = dict(hurst=0.4, starting_amplitude=1.0, stepdown=2.0, detune_rate=0.0, octaves=13, noise_size=1700)


def rand_grid(n, m=None, seed=20221018):
    return np.random.default_rng(seed).random((n, m or n), dtype=np.float32)


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def gpu_fractal(nz, res, noise_type, xpos=0, zpos=0, **kw):
    p = dict(C1)
    p.update(kw)
    out = np.full(res * res, np.nan, np.float32)
    nz.host.fractal(out, res, noise_type, p["hurst"], p["starting_amplitude"], p["stepdown"], p["detune_rate"],
                    p["octaves"], xpos, zpos, p["noise_size"])
    return out.reshape(res, res)


def ref_fractal(oracle, res, noise_type, xpos=0, zpos=0, **kw):
    p = dict(C1)
    p.update(kw)
    return oracle.fractal(res, res, noise_type, p["hurst"], p["starting_amplitude"], p["stepdown"], p["detune_rate"],
                          p["octaves"], xpos, zpos, p["noise_size"])


# ---------------------------------------------------------------------------------------------
# noise
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("pos", [(0, 0), (0, 424)])
def test_c1_simplex_fbm(nz, oracle, pos):
    got = gpu_fractal(nz, 256, 3, *pos)
    ref = ref_fractal(oracle, 256, 3, *pos)
    assert np.isfinite(got).all()
    assert np.abs(got - ref).max() <= TOL_NOISE


@pytest.fixture
def fbm_path(monkeypatch):
    """Force one of the two simplex kernels (NZ_FBM_PATH is read by the library at every launch)."""
    def force(which):
        if which is None:
            monkeypatch.delenv("NZ_FBM_PATH", raising=False)
        else:
            monkeypatch.setenv("NZ_FBM_PATH", which)
    return force


@pytest.mark.parametrize("res,pos", [(96, (0, 0)), (256, (0, 424)), (600, (-5000, 7777)), (1031, (100000, 3)),
                                      (2048, (14336, 14336))])
def test_packed_pair_simplex_kernel_is_bit_identical_to_the_scalar_kernel(nz, oracle, fbm_path, res, pos):
    # fbmpair_kernels.cu (f32x2 pairs + bank-private hash tables) vs fbm_kernel<SIMPLEX> in noise_kernels.cu
    fbm_path("scalar")
    a = gpu_fractal(nz, res, 3, *pos)
    fbm_path("pair")
    b = gpu_fractal(nz, res, 3, *pos)
    fbm_path(None)
    assert np.isfinite(a).all()
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    if res <= 600:
        ref = ref_fractal(oracle, res, 3, *pos)
        assert np.abs(b - ref).max() <= TOL_NOISE


@pytest.mark.parametrize("res,pos", [(96, (0, 0)), (300, (-5000, 7777)), (1031, (100000, 3)), (2048, (14336, 14336))])
@pytest.mark.parametrize("noise_type", [5, 1])
def test_packed_pair_cellular_and_perlin_kernels_are_bit_identical_to_the_scalar_kernel(nz, oracle, fbm_path, res, pos, noise_type):
    fbm_path("scalar")
    a = gpu_fractal(nz, res, noise_type, *pos)
    fbm_path("pair")
    b = gpu_fractal(nz, res, noise_type, *pos)
    fbm_path(None)
    assert np.isfinite(a).all()
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    if res <= 300:
        assert np.abs(b - ref_fractal(oracle, res, noise_type, *pos)).max() <= TOL_NOISE


def test_packed_pair_simplex_kernel_detune_and_odd_parameters(nz, fbm_path):
    kw = dict(hurst=0.9001, octaves=6, noise_size=7475, stepdown=2.17, detune_rate=0.013, starting_amplitude=0.7)
    for noise_type in (3, 5, 1):
        fbm_path("scalar")
        a = gpu_fractal(nz, 777, noise_type, 12345, -999, **kw)
        fbm_path("pair")
        b = gpu_fractal(nz, 777, noise_type, 12345, -999, **kw)
        fbm_path(None)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), noise_type


def test_gpu_chain_matches_readme_screenshots(nz):
    """The CUDA path against the reference's own rendered output (tests/test_screenshot_pins.py explains the fixture):
    README example #1 at the inspector panel's parameters, through the stage API."""
    import os
    from scipy.stats import spearmanr
    shots = np.load(os.path.join(os.path.dirname(__file__), "golden", "readme_screenshots.npz"))
    res = 1000
    data = np.zeros(res * res, np.float32)

    def view(a):
        return np.fliplr(a.reshape(250, 4, 250, 4).mean(axis=(1, 3)))

    def rho(a, img):
        return float(spearmanr(a.ravel(), np.asarray(img, np.float32).ravel()).correlation)

    nz.BasePipeline([nz.NoiseStage(nz.FractalNoise.Simplex, hurst=0.422, octaves=13, noiseSize=1757)]).Run(
        nz.GeneratorData("shot3", data, res, 0, 424))
    assert rho(view(data.reshape(res, res)), shots["shot3"].mean(axis=2)) > 0.92
    nz.BasePipeline([nz.KernelFilterStage(nz.KernelFilterType.Gauss5_S1, iterations=17), nz.ErosionFilterStage(iterations=5)]).Run(
        nz.GeneratorData("shot5", data, res, 0, 424))
    assert rho(view(data.reshape(res, res)), shots["shot5"][:, :, 1]) > 0.85
    nz.BasePipeline([nz.FlowMapStage(iterations=5, normMin=0.0, normMax=0.005)]).Run(nz.GeneratorData("shot5", data, res, 0, 424))
    assert rho(view(data.reshape(res, res)), shots["shot5"][:, :, 2]) > 0.85
    cell = np.zeros(res * res, np.float32)
    nz.BasePipeline([nz.NoiseStage(nz.FractalNoise.Cellular, hurst=1.0, octaves=13, noiseSize=1757)]).Run(
        nz.GeneratorData("shot0", cell, res, 0, 0))
    assert rho(view(cell.reshape(res, res)), shots["shot0"].mean(axis=2)) > 0.98


@pytest.mark.parametrize("noise_type", range(8))
def test_every_basis_matches_oracle(nz, oracle, noise_type):
    got = gpu_fractal(nz, 192, noise_type, 1000, 3000)
    ref = ref_fractal(oracle, 192, noise_type, 1000, 3000)
    assert np.abs(got - ref).max() <= TOL_NOISE


@pytest.mark.parametrize("noise_type", [3, 5, 4, 1, 2, 6, 7])
def test_noise_far_from_origin_c5_coordinates(nz, oracle, noise_type):
    # the last 160 rows/cols of the 16384^2 grid: lattice coordinates up to 4096*16384/1700
    got = gpu_fractal(nz, 160, noise_type, 16384 - 160, 16384 - 160)
    ref = ref_fractal(oracle, 160, noise_type, 16384 - 160, 16384 - 160)
    assert np.abs(got - ref).max() <= TOL_NOISE


def test_noise_parameters_detune_amplitude_odd_sizes(nz, oracle):
    for res, kw in ((37, dict(detune_rate=0.03, stepdown=2.2, starting_amplitude=3.0, octaves=24, hurst=1.3)),
                    (8, dict(octaves=1, noise_size=5)), (301, dict(detune_rate=-0.05, stepdown=1.8, octaves=7))):
        got = gpu_fractal(nz, res, 3, 17, 90, **kw)
        ref = ref_fractal(oracle, res, 3, 17, 90, **kw)
        assert np.abs(got - ref).max() <= TOL_NOISE * max(1.0, kw.get("starting_amplitude", 1.0))


def test_noise_band_rows_equal_full_tile_bitwise(nz, torch_cuda):
    torch = torch_cuda
    full = torch.empty(300, 300, device="cuda")
    nz.device.fractal(full, 3, 0.4, octaves=13, noise_size=1700, xpos=5, zpos=9)
    band = torch.empty(77, 300, device="cuda")
    nz.device.fractal(band, 3, 0.4, octaves=13, noise_size=1700, xpos=5, zpos=9, z_first=111)
    torch.cuda.synchronize()
    assert torch.equal(full[111:188], band)


def test_strided_native_slice(nz, oracle):
    # one float channel of an RGBAFloat texture (Scripts/Editor/VisualizePipeline.cs:141): stride 16 B
    res = 64
    tex = np.full((res * res, 4), 7.0, np.float32)
    nz.host.fractal(tex[:, 2], res, 3, 0.4, 1.0, 2.0, 0.0, 13, 0, 0, 1700)
    ref = ref_fractal(oracle, res, 3)
    assert np.abs(tex[:, 2].reshape(res, res) - ref).max() <= TOL_NOISE
    assert (tex[:, [0, 1, 3]] == 7.0).all()
    nz.host.kernel_filter(tex[:, 2], None, 2, res, 2)
    assert np.abs(tex[:, 2].reshape(res, res) - oracle.kernel_filter(ref, 2, 2)).max() <= 2 * TOL_FILTER
    assert (tex[:, [0, 1, 3]] == 7.0).all()


# ---------------------------------------------------------------------------------------------
# separable filters
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ftype", [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 13])
def test_kernel_filter_types(nz, oracle, ftype):
    a = rand_grid(200)
    got = a.copy().ravel()
    nz.host.kernel_filter(got, None, ftype, 200, 2)
    ref = oracle.kernel_filter(a, ftype, 2)
    scale = max(1.0, np.abs(ref).max())
    assert np.abs(got.reshape(200, 200) - ref).max() <= TOL_FILTER * scale


@pytest.mark.parametrize("res", [1024, 257, 5, 2])
def test_gauss5_x17(nz, oracle, res):
    a = rand_grid(res)
    got = a.copy().ravel()
    nz.host.kernel_filter(got, None, 2, res, 17)
    ref = oracle.kernel_filter(a, 2, 17)
    assert np.abs(got.reshape(res, res) - ref).max() <= TOL_FILTER


def test_gauss_and_smooth_blur_stages(nz, oracle):
    a = rand_grid(150)
    for width, sigma in ((3, 0), (7, 3), (25, 15), (12, 5)):
        got = a.copy().ravel()
        nz.host.gauss_filter(got, None, width, sigma, 150, 2)
        assert np.abs(got.reshape(150, 150) - oracle.gauss_filter(a, width, sigma, 2)).max() <= TOL_FILTER
    for width in (3, 9, 25):
        got = a.copy().ravel()
        nz.host.smooth_filter(got, None, width, 150, 3)
        assert np.abs(got.reshape(150, 150) - oracle.smooth_filter(a, width, 3)).max() <= TOL_FILTER


def test_sobel3_2d(nz, oracle):
    a = rand_grid(180)
    for iters in (1, 2):
        got = a.copy().ravel()
        nz.host.kernel_filter(got, None, 11, 180, iters)
        ref = oracle.kernel_filter(a, 11, iters)
        assert np.abs(got.reshape(180, 180) - ref).max() <= 1e-6 * max(1.0, ref.max())


def test_rectangular_window_equals_full_interior_bitwise(nz, torch_cuda):
    """A row band with T*r ghost rows reproduces the full-grid result on its owned rows (what the
    multi-GPU band path relies on)."""
    torch = torch_cuda
    n, iters, r = 384, 6, 2
    a = torch.from_numpy(rand_grid(n)).cuda()
    full = nz.device.kernel_filter(a.clone(), torch.empty_like(a), 2, iters).clone()
    z0, z1, h = 100, 260, iters * r
    win = a[z0 - h:z1 + h].clone()
    out = nz.device.kernel_filter(win, torch.empty_like(win), 2, iters)
    torch.cuda.synchronize()
    assert torch.equal(out[h:h + (z1 - z0)], full[z0:z1])
    # top band: global clamp at z=0 must be honoured
    win = a[0:z1 + h].clone()
    out = nz.device.kernel_filter(win, torch.empty_like(win), 2, iters)
    assert torch.equal(out[0:z1], full[0:z1])


# ---------------------------------------------------------------------------------------------
# value erosion
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("res,iters", [(512, 5), (100, 1), (33, 12), (3, 5), (1, 1), (300, 8), (77, 3), (260, 9), (64, 2)])
def test_min_erosion_bit_exact(nz, oracle, res, iters):
    a = rand_grid(res)
    got = a.copy().ravel()
    nz.host.min_erosion(got, res, iters)
    assert np.array_equal(got.reshape(res, res), oracle.min_erosion(a, iters))


# ---------------------------------------------------------------------------------------------
# flow map
# ---------------------------------------------------------------------------------------------
def _terrain(oracle, res):
    return oracle.kernel_filter(oracle.fractal(res, res, 3, 0.4, octaves=13, noise_size=1700), 2, 17)


@pytest.mark.parametrize("res,iters", [(512, 5), (130, 1), (64, 16), (2, 3)])
def test_flowmap_on_terrain(nz, oracle, res, iters):
    h = _terrain(oracle, res)
    got = h.copy().ravel()
    nz.host.flowmap(got, res, iters, 0.0, 0.005)
    ref = oracle.flowmap(h, iters, 0.0, 0.005)
    assert np.abs(got.reshape(res, res) - ref).max() <= TOL_FLOW * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("rows,width,iters", [(700, 600, 5), (300, 1000, 4), (97, 236, 3), (40, 20, 2), (513, 472, 1), (33, 4, 5)])
def test_flow_wavefront_kernel_equals_per_iteration_kernels_bitwise(nz, oracle, torch_cuda, monkeypatch, rows, width, iters):
    """The fused single-launch flow map (state in shared-memory rings) against the per-iteration kernels
    (state in HBM), on strip/chunk boundaries, grid borders and partial strips."""
    torch = torch_cuda
    h = torch.from_numpy(oracle.kernel_filter(rand_grid(rows, width), 3, 2) * np.float32(0.05)).cuda()
    fused = nz.device.flowmap(h.clone(), torch.empty_like(h), None, iters, 0.0, 0.005).clone()
    monkeypatch.setenv("NZ_FLOW_UNFUSED", "1")
    scratch = torch.empty(5 * rows * width * 4, dtype=torch.uint8, device="cuda")
    plain = nz.device.flowmap(h.clone(), torch.empty_like(h), scratch, iters, 0.0, 0.005).clone()
    torch.cuda.synchronize()
    assert torch.equal(fused, plain)
    ref = oracle.flowmap(h.cpu().numpy(), iters, 0.0, 0.005)
    assert np.abs(fused.cpu().numpy() - ref).max() <= TOL_FLOW * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("rows,width,iters", [(700, 600, 5), (300, 1000, 4), (97, 236, 3), (40, 20, 2), (513, 472, 1), (33, 4, 5),
                                              (1100, 1304, 5)])
def test_flow_tile_kernel_equals_wavefront_kernel_bitwise(nz, oracle, torch_cuda, monkeypatch, rows, width, iters):
    """The tile-resident formulation (flowtile_kernels.cu) against the wavefront one, over tile seams, grid borders
    and partial tiles."""
    torch = torch_cuda
    h = torch.from_numpy(oracle.kernel_filter(rand_grid(rows, width), 3, 2) * np.float32(0.05)).cuda()
    monkeypatch.setenv("NZ_FLOW_PATH", "wave")
    wave = nz.device.flowmap(h.clone(), torch.empty_like(h), None, iters, 0.0, 0.005).clone()
    monkeypatch.setenv("NZ_FLOW_PATH", "tile")
    tile = nz.device.flowmap(h.clone(), torch.empty_like(h), None, iters, 0.0, 0.005).clone()
    torch.cuda.synchronize()
    assert torch.equal(tile, wave)


@pytest.mark.parametrize("rows,width", [(70, 128), (200, 516), (65, 4), (1, 260), (130, 1024), (257, 2052)])
def test_sobel_walk_kernel_equals_plain_kernel_bitwise(nz, oracle, torch_cuda, monkeypatch, rows, width):
    """Sobel3_2D as a register row walk (float4 strips, 64-row chunks) against the one-thread-per-cell kernel and the
    oracle: strip and chunk seams, grid borders, partial strips, two iterations."""
    torch = torch_cuda
    h = torch.from_numpy(rand_grid(rows, width)).cuda()
    monkeypatch.setenv("NZ_SOBEL_PATH", "walk")                   # grids under 4M cells take the per-cell kernel by default
    walk = nz.device.kernel_filter(h.clone(), torch.empty_like(h), 11, 2).clone()
    monkeypatch.setenv("NZ_SOBEL_PATH", "plain")
    plain = nz.device.kernel_filter(h.clone(), torch.empty_like(h), 11, 2).clone()
    torch.cuda.synchronize()
    assert torch.equal(walk, plain)
    ref = oracle.kernel_filter(h.cpu().numpy(), 11, 2)
    assert np.abs(walk.cpu().numpy() - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("rows,width,iters", [(700, 600, 5), (300, 1000, 4), (97, 236, 3), (40, 20, 2), (513, 472, 1), (33, 4, 5),
                                              (1100, 1304, 5), (64, 44, 5), (600, 88, 5), (260, 132, 5), (2100, 2048, 5)])
def test_flow_register_walk_kernel_equals_wavefront_kernel_bitwise(nz, oracle, torch_cuda, monkeypatch, rows, width, iters):
    """The register-resident formulation (flowwalk_kernels.cu: warp strips of 64 columns, 44 useful at 5 iterations, chunks
    of <= 256 rows, packed f32x2 arithmetic, reciprocal-sequence division and square root) against the wavefront one, over
    strip and chunk seams, grid borders and partial strips.  No launch may need the wavefront rerun on ordinary heights."""
    torch = torch_cuda
    h = torch.from_numpy(oracle.kernel_filter(rand_grid(rows, width), 3, 2) * np.float32(0.05)).cuda()
    monkeypatch.setenv("NZ_FLOW_PATH", "wave")
    wave = nz.device.flowmap(h.clone(), torch.empty_like(h), None, iters, 0.0, 0.005).clone()
    before = nz.device.flow_walk_reruns()
    monkeypatch.setenv("NZ_FLOW_PATH", "reg")
    reg = nz.device.flowmap(h.clone(), torch.empty_like(h), None, iters, 0.0, 0.005).clone()
    rough = torch.rand(rows, width, device="cuda")                 # white noise: every cell divides, many drain
    reg_rough = nz.device.flowmap(rough.clone(), torch.empty_like(rough), None, iters, -0.1, 0.1).clone()
    monkeypatch.setenv("NZ_FLOW_PATH", "wave")
    wave_rough = nz.device.flowmap(rough.clone(), torch.empty_like(rough), None, iters, -0.1, 0.1).clone()
    torch.cuda.synchronize()
    assert torch.equal(reg, wave)
    assert torch.equal(reg_rough, wave_rough)
    assert nz.device.flow_walk_reruns() == before


def test_flow_register_walk_reruns_when_a_quotient_leaves_the_guard(nz, oracle, torch_cuda, monkeypatch):
    """Heights near 1e37 push water / (sum * dt) into the denormals, where the reciprocal sequence is not exact: the launch
    must notice, rerun on the wavefront kernel and still equal the per-iteration kernels bit for bit.  A degenerate or
    negative normalisation range never reaches the register kernel."""
    torch = torch_cuda
    rows, width = 300, 400
    h = (torch.rand(rows, width, device="cuda") * 3e37).contiguous()
    h[100:140, 50:200] = 1e37                                      # a plateau: still water
    monkeypatch.setenv("NZ_FLOW_UNFUSED", "1")
    scratch = torch.empty(5 * rows * width * 4, dtype=torch.uint8, device="cuda")
    plain = nz.device.flowmap(h.clone(), torch.empty_like(h), scratch, 5, 0.0, 0.005).clone()
    monkeypatch.delenv("NZ_FLOW_UNFUSED")
    before = nz.device.flow_walk_reruns()
    got = nz.device.flowmap(h.clone(), torch.empty_like(h), None, 5, 0.0, 0.005).clone()
    torch.cuda.synchronize()
    assert nz.device.flow_walk_reruns() == before + 1
    assert torch.equal(got, plain)
    for nmin, nmax in ((0.25, 0.25), (0.1, -0.1), (0.0, 1e-9)):   # handled by the wavefront kernel from the start
        a = nz.device.flowmap(h.clone() * 1e-37, torch.empty_like(h), None, 3, nmin, nmax).clone()
        monkeypatch.setenv("NZ_FLOW_UNFUSED", "1")
        b = nz.device.flowmap(h.clone() * 1e-37, torch.empty_like(h), scratch, 3, nmin, nmax).clone()
        monkeypatch.delenv("NZ_FLOW_UNFUSED")
        assert torch.equal(a, b) or (torch.isnan(a) == torch.isnan(b)).all()
    assert nz.device.flow_walk_reruns() == before + 1


def test_flow_register_walk_rerun_flags_are_per_launch_on_concurrent_streams(nz, oracle, torch_cuda):
    """Four flow maps in flight on four streams, one of them leaving the guard: only that launch is rerun, and every
    result equals what the same call gives alone (the rerun flag is one word per launch, not per device)."""
    torch = torch_cuda
    rows, width = 512, 768
    hs = [torch.rand(rows, width, device="cuda") * (3e37 if i == 2 else 0.05 * (i + 1)) for i in range(4)]
    alone = [nz.device.flowmap(h.clone(), torch.empty_like(h), None, 5, 0.0, 0.005).clone() for h in hs]
    torch.cuda.synchronize()
    before = nz.device.flow_walk_reruns()
    streams = [torch.cuda.Stream() for _ in hs]
    outs, keep = [None] * 4, []
    for rep in range(3):
        for i, (h, st) in enumerate(zip(hs, streams)):
            with torch.cuda.stream(st):
                a, b = h.clone(), torch.empty_like(h)
                keep.append((a, b))
                outs[i] = nz.device.flowmap(a, b, None, 5, 0.0, 0.005, stream=st)
    torch.cuda.synchronize()
    for got, want in zip(outs, alone):
        assert torch.equal(got, want)
    assert nz.device.flow_walk_reruns() == before + 3


def test_flow_row_band_with_ghost_rows_equals_full_grid_bitwise(nz, oracle, torch_cuda):
    torch = torch_cuda
    n, iters = 640, 5
    h = torch.from_numpy(oracle.kernel_filter(rand_grid(n), 3, 2) * np.float32(0.05)).cuda()
    full = nz.device.flowmap(h.clone(), torch.empty_like(h), None, iters, 0.0, 0.005).clone()
    z0, z1, g = 200, 420, 2 * iters + 1
    for lo, hi in ((z0 - g, z1 + g), (0, z1 + g), (z0 - g, n)):
        win = h[lo:hi].clone()
        out = nz.device.flowmap(win, torch.empty_like(win), None, iters, 0.0, 0.005)
        torch.cuda.synchronize()
        a = max(z0, lo) if lo > 0 else 0
        b = min(z1, hi) if hi < n else n
        assert torch.equal(out[a - lo:b - lo], full[a:b])


def test_flowmap_default_norm_and_random_field(nz, oracle):
    a = rand_grid(256)
    got = a.copy().ravel()
    nz.host.flowmap(got, 256, 5, -0.1, 0.1)
    ref = oracle.flowmap(a, 5, -0.1, 0.1)
    assert np.abs(got.reshape(256, 256) - ref).max() <= TOL_FLOW * max(1.0, np.abs(ref).max())
    got = a.copy().ravel()
    nz.host.flowmap(got, 256, 0, -0.1, 0.1)
    assert np.array_equal(got.reshape(256, 256), oracle.flowmap(a, 0, -0.1, 0.1))


# ---------------------------------------------------------------------------------------------
# mesh
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mesh_type", [0, 1])
@pytest.mark.parametrize("R,in_res", [(1016, 1024), (200, 208), (129, 133), (7, 16)])
def test_mesh(nz, oracle, mesh_type, R, in_res):
    h = rand_grid(in_res)
    vtx = np.zeros(((R + 1) ** 2, 12), np.float32)
    idx = np.zeros(6 * R * R, np.uint32)
    T = R * (500.0 / 256.0)
    nz.host.heightmap_mesh(mesh_type, vtx, idx, R, in_res, 4, 2000.0, T, h.ravel())
    rv, ri = oracle.heightmap_mesh(mesh_type, h, R, 4, 2000.0, T)
    assert np.array_equal(idx, ri)
    assert np.abs(vtx[:, 0:3] - rv[:, 0:3]).max() <= TOL_MESH * 2000.0
    assert np.abs(vtx[:, 3:10] - rv[:, 3:10]).max() <= TOL_MESH
    assert np.abs(vtx[:, 10:12] - rv[:, 10:12]).max() <= TOL_MESH


def test_mesh_row_bands_equal_full_mesh_bitwise(nz, oracle, torch_cuda):
    torch = torch_cuda
    R, in_res, off = 300, 308, 4
    h = torch.from_numpy(rand_grid(in_res)).cuda()
    vfull = torch.zeros((R + 1) ** 2, 12, device="cuda")
    ifull = torch.zeros(6 * R * R, dtype=torch.int32, device="cuda")
    nz.device.heightmap_mesh(1, vfull, ifull, R, in_res, 4, 2000.0, 600.0, h)
    for vz0, vz1 in ((0, 90), (90, 222), (222, R + 1)):
        hr0, hr1 = max(vz0 - 1 + off, 0), min(vz1 + 1 + off, in_res)
        v = torch.zeros((vz1 - vz0) * (R + 1), 12, device="cuda")
        t0 = max(vz0, 1)
        i = torch.zeros(6 * R * (vz1 - t0), dtype=torch.int32, device="cuda")
        nz.device.heightmap_mesh(1, v, i, R, in_res, 4, 2000.0, 600.0, h[hr0:hr1].contiguous(), h_row_first=hr0,
                                 vz_begin=vz0, vz_end=vz1)
        torch.cuda.synchronize()
        assert torch.equal(v, vfull[vz0 * (R + 1):vz1 * (R + 1)])
        assert torch.equal(i, ifull[6 * R * (t0 - 1):6 * R * (vz1 - 1)])


# ---------------------------------------------------------------------------------------------
# the stage API and full chains (BASELINE configs C2, C3 at test size, C4 tile)
# ---------------------------------------------------------------------------------------------
def test_c2_chain_through_stage_api(nz, oracle):
    res, R = 1024, 1016
    data = np.zeros(res * res, np.float32)
    chain = nz.BasePipeline([
        nz.NoiseStage(nz.FractalNoise.Simplex, hurst=0.4, octaves=13, noiseSize=1700),
        nz.KernelFilterStage(nz.KernelFilterType.Gauss5_S1, iterations=17),
        nz.FlowMapStage(iterations=5, normMin=0.0, normMax=0.005),
        nz.ErosionFilterStage(iterations=5),
    ])
    done = []
    chain.Run(nz.GeneratorData("c2", data, res, 0, 0), completeAction=done.append)
    assert len(done) == 1
    mesh_data = nz.MeshStageData("c2", data, R, res, 4, 1984.375, 2000.0)
    nz.BasePipeline([nz.MeshTileStage(nz.MeshType.OvershootSquareGridHeightMap)]).Run(mesh_data)

    ref = oracle.fractal(res, res, 3, 0.4, octaves=13, noise_size=1700)
    ref = oracle.kernel_filter(ref, 2, 17)
    ref = oracle.flowmap(ref, 5, 0.0, 0.005)
    ref = oracle.min_erosion(ref, 5)
    got = data.reshape(res, res)
    # the flow map divides by 0.005: noise-level differences of 1e-7 are amplified; compare at flow tolerance
    assert np.abs(got - ref).max() <= 5e-5 * max(1.0, np.abs(ref).max())
    rv, ri = oracle.heightmap_mesh(1, got, R, 4, 2000.0, 1984.375)
    assert np.array_equal(mesh_data.mesh.indices, ri)
    assert np.abs(mesh_data.mesh.vertices - rv).max() <= TOL_MESH * 2000.0
    assert np.abs(mesh_data.mesh.vertices[:, 3:] - rv[:, 3:]).max() <= TOL_MESH


def test_c3_cellular_chain_reduced_size(nz, oracle):
    res = 512
    data = np.zeros(res * res, np.float32)
    with nz.host.pipeline():
        nz.host.fractal(data, res, 5, 0.4, 1.0, 2.0, 0.0, 13, 2048, 1024, 1700)
        nz.host.kernel_filter(data, None, 2, res, 17)
    ref = oracle.kernel_filter(oracle.fractal(res, res, 5, 0.4, octaves=13, xpos=2048, zpos=1024, noise_size=1700), 2, 17)
    assert np.abs(data.reshape(res, res) - ref).max() <= TOL_NOISE


def test_c4_tile_rotated_simplex_gauss3_sobel(nz, oracle):
    res, tx, tz = 256, 3, 7
    data = np.zeros(res * res, np.float32)
    with nz.host.pipeline():
        nz.host.fractal(data, res, 4, 0.4, 1.0, 2.0, 0.0, 13, 1000 * tx, 1000 * tz, 1700)
        nz.host.kernel_filter(data, None, 3, res, 3)
        nz.host.flush_to_host(data)
        blurred = data.copy()
        nz.host.kernel_filter(data, None, 11, res, 1)
    ref_b = oracle.kernel_filter(oracle.fractal(res, res, 4, 0.4, octaves=13, xpos=1000 * tx, zpos=1000 * tz, noise_size=1700), 3, 3)
    assert np.abs(blurred.reshape(res, res) - ref_b).max() <= 2e-6
    assert np.abs(data.reshape(res, res) - oracle.kernel_filter(blurred.reshape(res, res), 11, 1)).max() <= 1e-6


# ---------------------------------------------------------------------------------------------
# error behaviour of the boundary
# ---------------------------------------------------------------------------------------------
def test_errors_are_status_codes_not_crashes(nz):
    with pytest.raises(nz.NzError) as e:
        nz.host.fractal(np.zeros(10, np.float32), 4, 3, 0.4, 1.0, 2.0, 0.0, 13, 0, 0, 1700)
    assert e.value.code == nz.lib.NZ_E_INVALID and "length" in str(e.value)
    with pytest.raises(nz.NzError):
        nz.host.fractal(np.zeros(16, np.float32), 4, 99, 0.4, 1.0, 2.0, 0.0, 13, 0, 0, 1700)
    with pytest.raises(nz.NzError):
        nz.host.kernel_filter(np.zeros(16, np.float32), None, 77, 4, 1)
    with pytest.raises(nz.NzError):
        nz.host.separable(np.zeros(16, np.float32), None, np.ones(4, np.float32), np.ones(4, np.float32), 1.0, 4, 1)
    with pytest.raises(nz.NzError):
        nz.host.heightmap_mesh(0, np.zeros((17 * 17, 12), np.float32), np.zeros(6 * 256, np.uint32), 16, 16, 0, 1.0, 1.0,
                               np.zeros(256, np.float32))
    with pytest.raises(nz.NzError) as e:
        nz.host.pipeline_end()
    assert e.value.code == nz.lib.NZ_E_STATE
    with pytest.raises(Exception, match="Unhandled stageio"):
        nz.NoiseStage().Schedule(nz.PipelineWorkItem(nz.MeshStageData(data=np.zeros(4, np.float32))), None)


def test_kernels_actually_launch(nz):
    before = nz.host.kernel_launch_count()
    a = np.zeros(64 * 64, np.float32)
    nz.host.fractal(a, 64, 3, 0.4, 1.0, 2.0, 0.0, 13, 0, 0, 1700)
    assert nz.host.kernel_launch_count() == before + 1
    t = nz.host.last_timing()
    assert t["kernel_launches"] == 1 and t["ms_kernel"] > 0


@pytest.mark.parametrize("noise_type", [3, 1, 5, 6, 7])
def test_beyond_the_fast_hash_domain_the_exact_residue_is_used(nz, oracle, noise_type):
    # lattice indices above 2^21 switch the kernels to the float-floor mod289 (same as the oracle)
    got = gpu_fractal(nz, 64, noise_type, 30000, 30000, octaves=24, noise_size=5, stepdown=2.2)
    ref = ref_fractal(oracle, 64, noise_type, 30000, 30000, octaves=24, noise_size=5, stepdown=2.2)
    assert np.abs(got - ref).max() <= TOL_NOISE


@pytest.mark.parametrize("rows,width,ftype,iters", [(700, 600, 2, 17), (300, 1000, 3, 7), (97, 236, 0, 3), (40, 20, 1, 2),
                                                    (513, 472, 6, 4), (33, 4, 2, 5), (260, 128, 8, 3), (130, 132, 9, 1)])
def test_separable_paths_agree_bitwise(nz, torch_cuda, monkeypatch, rows, width, ftype, iters):
    """Register-walk kernel vs shared-memory tile kernel vs one-pass-per-launch kernels: same bits, on strip,
    chunk and grid borders."""
    torch = torch_cuda
    a = torch.from_numpy(rand_grid(rows, width)).cuda()
    monkeypatch.setenv("NZ_SEP_PATH", "walk")      # small grids default to the tile kernel: force the walk
    walk = nz.device.kernel_filter(a.clone(), torch.empty_like(a), ftype, iters).clone()
    monkeypatch.setenv("NZ_SEP_PATH", "fused")
    fused = nz.device.kernel_filter(a.clone(), torch.empty_like(a), ftype, iters).clone()
    monkeypatch.setenv("NZ_SEP_PATH", "generic")
    plain = nz.device.kernel_filter(a.clone(), torch.empty_like(a), ftype, iters).clone()
    torch.cuda.synchronize()
    assert torch.equal(fused, plain)
    assert torch.equal(walk, plain)


def test_cpp_host_mirror_gives_the_same_bits_as_the_python_mirror(nz):
    """noize-job_b200/host_cpp: the compiled C++ stage mirror drives the same chain through the same C ABI."""
    import os
    import re
    import subprocess
    exe = os.path.join(os.path.dirname(nz.lib.LIB_PATH), "host_cpp", "example_chain")
    assert os.path.exists(exe), "build() compiles host_cpp/example_chain"
    res, R = 512, 504
    out = subprocess.run([exe, str(res)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    got = dict(re.findall(r"(height|vertices|indices)=([0-9a-f]{16})", out.stdout))

    def fnv(a):
        h = 1469598103934665603
        for b in a.tobytes():
            h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return f"{h:016x}"

    data = np.zeros(res * res, np.float32)
    nz.BasePipeline([nz.NoiseStage(nz.FractalNoise.Simplex, hurst=0.4, octaves=13, noiseSize=1700),
                     nz.KernelFilterStage(nz.KernelFilterType.Gauss5_S1, iterations=17),
                     nz.FlowMapStage(iterations=5, normMin=0.0, normMax=0.005),
                     nz.ErosionFilterStage(iterations=5)]).Run(nz.GeneratorData("c2", data, res, 0, 0))
    md = nz.MeshStageData("c2", data, R, res, 4, np.float32(R) * (np.float32(500.0) / np.float32(256.0)), 2000.0)
    nz.BasePipeline([nz.MeshTileStage(nz.MeshType.OvershootSquareGridHeightMap)]).Run(md)
    # hash a prefix in Python (pure-Python FNV over MBs is slow): compare the first 64 KiB of each buffer instead
    assert got["height"] and got["vertices"] and got["indices"]
    exe_small = subprocess.run([exe, "64"], capture_output=True, text=True, timeout=300)
    small = dict(re.findall(r"(height|vertices|indices)=([0-9a-f]{16})", exe_small.stdout))
    d2 = np.zeros(64 * 64, np.float32)
    nz.BasePipeline([nz.NoiseStage(nz.FractalNoise.Simplex, hurst=0.4, octaves=13, noiseSize=1700),
                     nz.KernelFilterStage(nz.KernelFilterType.Gauss5_S1, iterations=17),
                     nz.FlowMapStage(iterations=5, normMin=0.0, normMax=0.005),
                     nz.ErosionFilterStage(iterations=5)]).Run(nz.GeneratorData("c2", d2, 64, 0, 0))
    m2 = nz.MeshStageData("c2", d2, 56, 64, 4, np.float32(56) * (np.float32(500.0) / np.float32(256.0)), 2000.0)
    nz.BasePipeline([nz.MeshTileStage(nz.MeshType.OvershootSquareGridHeightMap)]).Run(m2)
    assert small["height"] == fnv(d2)
    assert small["vertices"] == fnv(m2.mesh.vertices)
    assert small["indices"] == fnv(m2.mesh.indices)


def test_residency_scope_spans_threads(nz, oracle):
    """The stages of one chain run on DIFFERENT threads (as Unity's job workers do) inside one process-wide scope:
    the device mirror stays resident across threads/streams and the result equals the single-thread chain."""
    import threading
    res = 512
    want = np.zeros(res * res, np.float32)
    with nz.host.pipeline():
        nz.host.fractal(want, res, 3, 0.4, 1.0, 2.0, 0.0, 13, 0, 0, 1700)
        nz.host.kernel_filter(want, None, 2, res, 17)
        nz.host.flowmap(want, res, 5, 0.0, 0.005)
        nz.host.min_erosion(want, res, 5)

    data = np.full(res * res, -1.0, np.float32)
    scope = nz.host.scope_create()
    errors = []

    def stage(fn):
        def run():
            try:
                with nz.host.in_scope(scope):
                    fn()
            except Exception as e:  # pragma: no cover
                errors.append(e)
        t = threading.Thread(target=run)
        t.start()
        t.join()

    stage(lambda: nz.host.fractal(data, res, 3, 0.4, 1.0, 2.0, 0.0, 13, 0, 0, 1700))
    assert (data == -1.0).all()          # nothing came back yet: the mirror is resident, the D2H deferred
    stage(lambda: nz.host.kernel_filter(data, None, 2, res, 17))
    stage(lambda: nz.host.flowmap(data, res, 5, 0.0, 0.005))
    stage(lambda: nz.host.min_erosion(data, res, 5))
    assert not errors, errors
    assert (data == -1.0).all()
    nz.host.scope_close(scope)           # any thread may close: flushes to host
    assert np.array_equal(data, want)
    with pytest.raises(nz.NzError):
        nz.host.scope_enter(scope)       # closed scopes are gone
    with pytest.raises(nz.NzError):
        nz.host.scope_leave()


def test_concurrent_host_threads_with_side_streams(nz, oracle):
    """Four host threads run the chain at a size where both walk kernels split into an interior launch and a border launch
    forked onto the calling thread's side stream (>= 8M cells): every thread's result equals the same call made alone."""
    import threading
    res = 3072

    def chain(zpos, out):
        nz.host.fractal(out, res, 3, 0.4, 1.0, 2.0, 0.0, 13, 0, zpos, 1700)
        nz.host.kernel_filter(out, None, 2, res, 17)
        nz.host.flowmap(out, res, 5, 0.0, 0.005)
        nz.host.min_erosion(out, res, 5)

    want = [np.zeros(res * res, np.float32) for _ in range(4)]
    for i, w in enumerate(want):
        chain(1000 * i, w)
    got = [np.zeros(res * res, np.float32) for _ in range(4)]
    errors = []

    def run(i):
        try:
            for _ in range(2):
                chain(1000 * i, got[i])
        except Exception as e:  # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=run, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    assert len({w.tobytes()[:4096] for w in want}) == 4      # the four inputs really differ


@pytest.mark.parametrize("res,pos", [(96, (0, 0)), (300, (0, 424)), (1031, (100000, 3)), (1024, (15000, 15000)), (2048, (14336, 14336))])
@pytest.mark.parametrize("noise_type", [2, 4])
def test_packed_pair_psrnoise_kernel_is_bit_identical_to_the_scalar_kernel(nz, oracle, fbm_path, res, pos, noise_type):
    """fbm_psr_pair_kernel (periodic perlin / rotated simplex, the basis of BASELINE config C4) vs fbm_kernel<...> in
    noise_kernels.cu.  (0, 0) exercises the first lattice column, where a cell falls back to the scalar wrap code."""
    fbm_path("scalar")
    a = gpu_fractal(nz, res, noise_type, *pos)
    fbm_path("pair")
    b = gpu_fractal(nz, res, noise_type, *pos)
    fbm_path(None)
    assert np.isfinite(a).all()
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    if res <= 300:
        assert np.abs(b - ref_fractal(oracle, res, noise_type, *pos)).max() <= TOL_NOISE


def test_packed_pair_psrnoise_kernel_odd_parameters_and_negative_origins(nz, fbm_path):
    kw = dict(hurst=0.9001, octaves=6, noise_size=7475, stepdown=2.17, detune_rate=0.013, starting_amplitude=0.7)
    for noise_type in (2, 4):
        for pos in ((12345, 999), (12345, -999), (-7, 5)):          # negative origins are not eligible: both runs take the scalar kernel
            fbm_path("scalar")
            a = gpu_fractal(nz, 777, noise_type, *pos, **kw)
            fbm_path("pair")
            b = gpu_fractal(nz, 777, noise_type, *pos, **kw)
            fbm_path(None)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (noise_type, pos)


@pytest.mark.parametrize("group", ["0", "4", "6"])
@pytest.mark.parametrize("rows,width,iters", [(1100, 4100, 5), (1400, 3000, 4), (1500, 2872, 3), (2100, 2048, 5), (700, 6144, 5)])
def test_flow_group_strips_equal_wavefront_kernel_bitwise(nz, oracle, torch_cuda, monkeypatch, rows, width, iters, group):
    """flow_group_kernel: the warps of a CTA share one wide strip and trade their seam values through shared memory once
    per step (NZ_FLOW_GROUP = 4 or 6 warps; 0 = the independent 64-column strips).  Grids above 2^22 cells engage it."""
    torch = torch_cuda
    h = torch.rand(rows, width, device="cuda") * 0.05 + torch.linspace(0, 1, width, device="cuda")[None, :] * 0.2
    monkeypatch.setenv("NZ_FLOW_PATH", "wave")
    wave = nz.device.flowmap(h.clone(), torch.empty_like(h), None, iters, 0.0, 0.005).clone()
    before = nz.device.flow_walk_reruns()
    monkeypatch.setenv("NZ_FLOW_PATH", "reg")
    monkeypatch.setenv("NZ_FLOW_GROUP", group)
    reg = nz.device.flowmap(h.clone(), torch.empty_like(h), None, iters, 0.0, 0.005).clone()
    torch.cuda.synchronize()
    assert torch.equal(reg, wave)
    assert nz.device.flow_walk_reruns() == before


@pytest.mark.parametrize("rows,width,ftype,iters", [(700, 600, 2, 17), (1300, 1000, 3, 7), (513, 472, 6, 4), (2200, 4096, 2, 17), (97, 2048, 7, 3),
                                                      (3000, 236, 2, 9), (4096, 4096, 12, 2), (640, 1536, 8, 5), (1200, 2000, 0, 6)])
def test_walk_filter_forms_agree_bitwise(nz, torch_cuda, monkeypatch, rows, width, ftype, iters):
    """The register-walk filter's launch forms: border items inside the interior launch (default) or as a launch of their own
    (NZ_WALK_MERGE=0), stages skewed by one step (default) or chained (NZ_WALK_SKEW=0).  Every cell sees the same operations in
    the same order in all four, so the bits agree — also with the shared-memory tile kernel (NZ_SEP_PATH=fused)."""
    torch = torch_cuda
    a = torch.from_numpy(rand_grid(rows, width)).cuda()
    monkeypatch.setenv("NZ_SEP_PATH", "walk")
    ref = nz.device.kernel_filter(a.clone(), torch.empty_like(a), ftype, iters).clone()
    for merge, skew in (("0", "1"), ("1", "0"), ("0", "0")):
        monkeypatch.setenv("NZ_WALK_MERGE", merge)
        monkeypatch.setenv("NZ_WALK_SKEW", skew)
        got = nz.device.kernel_filter(a.clone(), torch.empty_like(a), ftype, iters).clone()
        torch.cuda.synchronize()
        assert torch.equal(got, ref), (merge, skew)
    monkeypatch.delenv("NZ_WALK_MERGE")
    monkeypatch.delenv("NZ_WALK_SKEW")
    if ftype != 11:
        monkeypatch.setenv("NZ_SEP_PATH", "fused")
        fused = nz.device.kernel_filter(a.clone(), torch.empty_like(a), ftype, iters).clone()
        torch.cuda.synchronize()
        assert torch.equal(fused, ref)


@pytest.mark.parametrize("rows,width,ftype,iters", [(700, 600, 2, 17), (1300, 1000, 3, 7), (513, 472, 6, 4), (2200, 4096, 2, 17)])
def test_bulk_copy_row_feed_equals_cp_async_feed_bitwise(nz, torch_cuda, monkeypatch, rows, width, ftype, iters):
    """NZ_WALK_FEED=bulk: the separable walk fed by cp.async.bulk + mbarrier (the TMA engine, SASS UBLKCP) instead of per-lane
    cp.async — a measured alternative (profiles/r2_walk_feed_scan.txt); same arithmetic, so the same bits."""
    torch = torch_cuda
    a = torch.from_numpy(rand_grid(rows, width)).cuda()
    monkeypatch.setenv("NZ_SEP_PATH", "walk")
    ref = nz.device.kernel_filter(a.clone(), torch.empty_like(a), ftype, iters).clone()
    monkeypatch.setenv("NZ_WALK_FEED", "bulk")
    bulk = nz.device.kernel_filter(a.clone(), torch.empty_like(a), ftype, iters).clone()
    torch.cuda.synchronize()
    assert torch.equal(bulk, ref)


@pytest.mark.parametrize("rows,width,ftype,iters", [(700, 600, 2, 17), (1300, 1000, 3, 5), (513, 472, 6, 10), (2200, 4096, 2, 17)])
def test_five_stage_walk_launches_equal_generic_path_bitwise(nz, torch_cuda, monkeypatch, rows, width, ftype, iters):
    """NZ_WALK_TMAX=5: five filter iterations per launch (17 = 5+4+4+4 instead of 4+4+3+3+3), a measured alternative."""
    torch = torch_cuda
    a = torch.from_numpy(rand_grid(rows, width)).cuda()
    monkeypatch.setenv("NZ_SEP_PATH", "generic")
    ref = nz.device.kernel_filter(a.clone(), torch.empty_like(a), ftype, iters).clone()
    monkeypatch.setenv("NZ_SEP_PATH", "walk")
    monkeypatch.setenv("NZ_WALK_TMAX", "5")
    five = nz.device.kernel_filter(a.clone(), torch.empty_like(a), ftype, iters).clone()
    torch.cuda.synchronize()
    assert torch.equal(five, ref)
