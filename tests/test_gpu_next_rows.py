"""GPU parity of the SURVEY section-8f rows (thermal erosion, element-wise stages) through the C ABI, against the
oracle.  Tolerances: thermal erosion, constant, reduce (sub/mul/max/min), crop, map range: BIT-EXACT (the kernels
perform the oracle's operations in the oracle's order); root-sum-squares, curve, normalize: bit-exact as well
(IEEE sqrt / division / fma), asserted as such."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def rnd(n, seed=20221018, lo=0.0, hi=1.0):
    return (np.random.default_rng(seed).random((n, n), dtype=np.float32) * np.float32(hi - lo) + np.float32(lo))


@pytest.mark.parametrize("res,iters,talus", [(64, 1, 45), (257, 2, 30), (1024, 3, 60), (130, 8, 10)])
def test_thermal_erosion_bit_exact(nz, oracle, res, iters, talus):
    g = rnd(res, res)
    got = g.copy().reshape(-1)
    nz.host.thermal_erosion(got, float(talus), 0.5, 0.75, iters, res)
    ref = oracle.thermal_erosion(g, talus, 0.5, 0.75, iters)
    assert np.array_equal(bits(got.reshape(res, res)), bits(ref))
    assert not np.array_equal(got.reshape(res, res), g)


def test_thermal_erosion_on_smooth_terrain_and_device_layer(nz, oracle):
    import torch
    res = 512
    g = oracle.kernel_filter(oracle.fractal(res, res, 3, 0.4, octaves=13, noise_size=170), 2, 3)
    t = torch.from_numpy(g).cuda()
    nz.device.thermal_erosion(t, 35.0, 0.4, 4.0, 5)
    ref = oracle.thermal_erosion(g, 35.0, 0.4, 4.0, 5)
    assert np.array_equal(bits(t.cpu().numpy()), bits(ref))


@pytest.mark.parametrize("res,iters,talus", [(64, 1, 45), (257, 2, 30), (1024, 3, 60), (130, 8, 10), (131, 3, 20), (6, 2, 45),
                                             (3, 1, 45), (200, 5, 5), (1500, 2, 25)])
def test_thermal_erosion_fused_tile_kernel_bit_exact(nz, oracle, res, iters, talus):
    """One launch per iteration (the four phases on a shared-memory tile with a 4-cell halo, ping-pong buffers) against
    the oracle and, through it, the four-launch path: tile seams, odd resolutions, grids smaller than a tile."""
    import torch
    g = rnd(res, res + iters)
    t, tmp = torch.from_numpy(g).cuda(), torch.full((res, res), float("nan"), device="cuda")
    out = nz.device.thermal_erosion(t, float(talus), 0.5, 0.75, iters, tmp=tmp)
    ref = oracle.thermal_erosion(g, talus, 0.5, 0.75, iters)
    assert np.array_equal(bits(out.cpu().numpy()), bits(ref))
    inplace = torch.from_numpy(g).cuda()
    nz.device.thermal_erosion(inplace, float(talus), 0.5, 0.75, iters)
    assert torch.equal(out, inplace)


@pytest.mark.parametrize("res", [33, 256])
def test_constant_reduce_bit_exact(nz, oracle, res):
    a, b = rnd(res, 1), rnd(res, 2, -0.5, 0.5)
    for op, val in ((0, 0.37), (1, 0.5)):
        got = a.copy().reshape(-1)
        nz.host.constant(got, None, op, val, res)
        assert np.array_equal(bits(got.reshape(res, res)), bits(oracle.constant(a, op, val))), op
    for op in range(5):
        got = a.copy().reshape(-1)
        nz.host.reduce(got, b.reshape(-1), None, op, res)
        assert np.array_equal(bits(got.reshape(res, res)), bits(oracle.reduce(a, b, op))), op


def test_curve_crop_range_normalize_bit_exact(nz, oracle):
    res = 200
    v = rnd(res, 5, -0.2, 1.2)
    curve = np.array([(i / 256) ** 2 for i in range(256)], np.float32)
    got = v.copy().reshape(-1)
    nz.host.curve(got, None, curve, res)
    assert np.array_equal(bits(got.reshape(res, res)), bits(oracle.curve(v, curve)))
    for out_res, off in ((120, 0), (120, 40), (64, 150)):
        out = np.full(out_res * out_res, np.nan, np.float32)
        nz.host.crop(v.reshape(-1), res, out, out_res, off)
        assert np.array_equal(bits(out.reshape(out_res, out_res)), bits(oracle.crop(v, out_res, off)))
    r = nz.host.map_range(v.reshape(-1))
    assert np.array_equal(bits(r), bits(oracle.map_range(v)))
    assert np.array_equal(bits(nz.host.map_range(v.reshape(-1), lim_min=-3.0)), bits(oracle.map_range(v, lim_min=-3.0)))
    got = v.copy().reshape(-1)
    nz.host.normalize(got, None, r, res)
    assert np.array_equal(bits(got.reshape(res, res)), bits(oracle.normalize(v, r)))


def test_strided_slices_and_bad_arguments(nz, oracle):
    res = 64
    tex = np.zeros((res * res, 4), np.float32)            # one channel of an RGBAFloat texture (stride 16)
    a = rnd(res, 9)
    tex[:, 2] = a.reshape(-1)
    nz.host.constant(tex[:, 2], None, 0, 0.25, res)
    assert np.array_equal(tex[:, 2].reshape(res, res), oracle.constant(a, 0, 0.25)) and not tex[:, [0, 1, 3]].any()
    with pytest.raises(nz.NzError):
        nz.host.constant(a.reshape(-1), None, 7, 0.5, res)
    with pytest.raises(nz.NzError):
        nz.host.reduce(a.reshape(-1), a.reshape(-1)[: res], None, 0, res)
    with pytest.raises(nz.NzError):
        nz.host.curve(a.reshape(-1), None, np.zeros(1, np.float32), res)


def test_stage_mirror_chain_with_reduce_and_curve(nz, oracle):
    """A demo-style chain (DynamicNoise.unity: filter -> Invert-like constant -> curve) and a ReducePipeline-style
    two-input reduce, all inside one residency scope; compared with the oracle stage by stage composition."""
    res = 256
    data = np.zeros(res * res, np.float32)
    right = np.zeros(res * res, np.float32)
    curve_fn = lambda t: t * t * (3 - 2 * t)
    with nz.host.pipeline():
        nz.BasePipeline([nz.NoiseStage(nz.FractalNoise.Simplex, hurst=0.5, octaves=6, noiseSize=300),
                         nz.KernelFilterStage(nz.KernelFilterType.Gauss3_S1, iterations=3),
                         nz.StageThermalErosion(iterations=2, talus=20, increment=0.5, meshHeightWidthRatio=0.75),
                         nz.CurveStage(curve_fn, samples=128),
                         nz.ConstantStage(nz.ConstantOperationType.MULTIPLY, 0.8)]).Run(nz.GeneratorData("l", data, res, 0, 0))
        nz.BasePipeline([nz.NoiseStage(nz.FractalNoise.Perlin, hurst=0.3, octaves=4, noiseSize=200)]).Run(
            nz.GeneratorData("r", right, res, 100, 50))
        rp = nz.BasePipeline([nz.ReduceStage(nz.ReductionType.MAX)])
        item = nz.ReduceData("lr", data, right, res, 0, 0)
        rp.Run(item)
    o = oracle
    left = o.fractal(res, res, 3, 0.5, octaves=6, noise_size=300)
    left = o.kernel_filter(left, 3, 3)
    left = o.thermal_erosion(left, 20, 0.5, 0.75, 2)
    s = np.float32(128)
    left = o.curve(left, np.array([curve_fn(float(np.float32(i) / s)) for i in range(128)], np.float32))
    left = o.constant(left, 0, 0.8)
    r = o.fractal(res, res, 1, 0.3, octaves=4, xpos=100, zpos=50, noise_size=200)
    want = o.reduce(left, r, 3)
    assert np.abs(data.reshape(res, res) - want).max() <= 2e-6
    assert np.abs(right.reshape(res, res) - r).max() <= 1e-6


def test_crop_stage_reference_behaviour_and_center(nz, oracle):
    g = rnd(96, 4)
    out = np.zeros(64 * 64, np.float32)
    nz.BasePipeline([nz.CropStage()]).Run(nz.DownsampleData("c", out, 64, 96, g.reshape(-1)))
    assert np.array_equal(out.reshape(64, 64), g[:64, :64])
    nz.BasePipeline([nz.CropStage("center")]).Run(nz.DownsampleData("c", out, 64, 96, g.reshape(-1)))
    assert np.array_equal(out.reshape(64, 64), g[16:80, 16:80])


# ---- subtractive-flow erosion -------------------------------------------------------------------
@pytest.mark.parametrize("res,cycles", [(64, 1), (130, 3), (512, 5), (33, 0)])
def test_subtractive_flow_erosion_bit_exact(nz, oracle, res, cycles):
    """Same operations in the same order as the oracle (IEEE add/mul/div/sqrt/fma, --fmad=false): bit-exact."""
    g = oracle.kernel_filter(oracle.fractal(res, res, 3, 0.4, octaves=13, noise_size=170), 2, 3)
    got = g.copy().reshape(-1)
    nz.host.subtractive_flow_erosion(got, res, cycles, 0.1, 0.0, 0.005)
    ref = oracle.subtractive_flow_erosion(g, cycles, 0.1, 0.0, 0.005)
    assert np.array_equal(bits(got.reshape(res, res)), bits(ref))
    assert (cycles == 0) == np.array_equal(got.reshape(res, res), g)


def test_subtractive_flow_erosion_stage_and_device_layer(nz, oracle):
    import torch
    res = 256
    g = oracle.kernel_filter(oracle.fractal(res, res, 5, 0.4, octaves=13, noise_size=170), 2, 3)
    data = g.copy().reshape(-1)
    stage = nz.ErosionStageSubtractiveFlow(erosiveIterations=4, erosiveFactor=0.05)   # default normalisation -0.1 / +0.1
    nz.BasePipeline([stage]).Run(nz.GeneratorData("sub", data, res, 0, 0))
    ref = oracle.subtractive_flow_erosion(g, 4, 0.05, -0.1, 0.1)
    assert np.array_equal(bits(data.reshape(res, res)), bits(ref))
    # rectangular window on the device layer, random (rough) field
    r = np.random.default_rng(3).random((96, 200), dtype=np.float32)
    t = torch.from_numpy(r).cuda()
    nz.device.subtractive_flow_erosion(t, 3, 0.1, 0.0, 0.5)
    assert np.array_equal(bits(t.cpu().numpy()), bits(oracle.subtractive_flow_erosion(r, 3, 0.1, 0.0, 0.5)))
    with pytest.raises(nz.NzError):
        nz.host.subtractive_flow_erosion(data, res, -1, 0.1, 0.0, 0.005)


@pytest.mark.parametrize("rows,width,cycles", [(700, 1000, 3), (300, 2048, 5), (257, 132, 2), (600, 600, 6)])
def test_subtractive_flow_fused_cycles_equal_per_iteration_kernels_bitwise(nz, oracle, monkeypatch, rows, width, cycles):
    """One launch per cycle on the SM-resident tile kernel (flows read and written once per cycle, erosion as the tile's
    epilogue) against the per-iteration HBM kernels, over many tiles, tile seams and grid borders; 6 cycles exceed the fused
    form's 5 iterations and take the per-iteration path either way."""
    import torch
    r = np.random.default_rng(11).random((rows, width), dtype=np.float32) * np.float32(0.2)
    monkeypatch.setenv("NZ_SUBFLOW_PATH", "unfused")
    a = torch.from_numpy(r).cuda()
    nz.device.subtractive_flow_erosion(a, cycles, 0.1, 0.0, 0.05)
    monkeypatch.delenv("NZ_SUBFLOW_PATH")
    b = torch.from_numpy(r).cuda()
    nz.device.subtractive_flow_erosion(b, cycles, 0.1, 0.0, 0.05)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    if rows * width <= 300000:
        assert np.array_equal(bits(b.cpu().numpy()), bits(oracle.subtractive_flow_erosion(r, cycles, 0.1, 0.0, 0.05)))


@pytest.mark.parametrize("res", [200, 1024])
def test_deferred_pointwise_run_is_one_pass_and_bit_exact(nz, oracle, res):
    """Inside a scope nz_constant / nz_normalize are deferred and applied as ONE pass with the next nz_curve (or when the
    values are needed): fewer launches, the same bits as the oracle's stage-by-stage composition."""
    v = rnd(res, 11, -0.3, 1.4)
    curve = np.array([np.sqrt(i / 63) for i in range(64)], np.float32)
    rng = oracle.map_range(v)
    o = oracle
    want = o.constant(o.curve(o.normalize(o.constant(v, 0, 0.9), rng), curve), 0, 0.5)
    want_bin = o.constant(want, 1, 0.25)

    got = v.copy().reshape(-1)
    before = nz.host.kernel_launch_count()
    with nz.host.pipeline():
        nz.host.constant(got, None, 0, 0.9, res)
        nz.host.normalize(got, None, rng, res)
        nz.host.curve(got, None, curve, res)              # run of three -> one launch
        nz.host.constant(got, None, 0, 0.5, res)          # deferred until ...
        mid = nz.host.kernel_launch_count() - before
        nz.host.flush_to_host(got)                        # ... the values are needed
        assert np.array_equal(bits(got.reshape(res, res)), bits(want))
        nz.host.constant(got, None, 1, 0.25, res)         # applied by the download at scope close
    assert mid == 1, mid
    assert nz.host.kernel_launch_count() - before == 3
    assert np.array_equal(bits(got.reshape(res, res)), bits(want_bin))

    # a stage that reads the values applies the run first; a stage that overwrites them drops it
    a = v.copy().reshape(-1)
    b = np.zeros(res * res, np.float32)
    with nz.host.pipeline():
        nz.host.constant(a, None, 0, 0.9, res)
        nz.host.kernel_filter(a, None, 3, res, 1)
        nz.host.constant(b, None, 0, 123.0, res)
        nz.host.fractal(b, res, 3, 0.4, 1.0, 2.0, 0.0, 4, 0, 0, 300)
    assert np.array_equal(bits(a.reshape(res, res)), bits(o.kernel_filter(o.constant(v, 0, 0.9), 3, 1)))
    assert np.abs(b.reshape(res, res) - o.fractal(res, res, 3, 0.4, octaves=4, noise_size=300)).max() <= 1e-6
    # more steps than one pass holds (PW_CHAIN_MAX = 8)
    c = v.copy().reshape(-1)
    wantc = v
    with nz.host.pipeline():
        for k in range(19):
            nz.host.constant(c, None, 0, 1.0 + 0.01 * k, res)
            wantc = o.constant(wantc, 0, 1.0 + 0.01 * k)
    assert np.array_equal(bits(c.reshape(res, res)), bits(wantc))
