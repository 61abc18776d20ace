"""Hypothesis-driven shape fuzzing of the device layer against the oracle (test plan iv): odd
and tiny grids, grids narrower than the kernel, rectangular windows, every separable width, random taps."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

pytestmark = pytest.mark.gpu
FUZZ = settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)


def grid(rows, width, seed):
    return np.random.default_rng(seed).random((rows, width), dtype=np.float32)


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@FUZZ
@given(rows=st.integers(1, 200), width=st.integers(1, 300), half=st.integers(1, 12), iters=st.integers(1, 6),
       seed=st.integers(0, 2**31 - 1), scale=st.sampled_from([1.0, 0.25]))
def test_separable_any_shape(nz, oracle, rows, width, half, iters, seed, scale):
    import torch
    g = grid(rows, width, seed)
    k = np.random.default_rng(seed + 1).random(2 * half + 1, dtype=np.float32)
    k /= k.sum()
    kz = k[::-1].copy() if seed & 1 else k
    a = dev(g)
    res = nz.device.separable(a, torch.empty_like(a), k, kz, scale, iters)
    ref = oracle.separable(g, k, kz, scale, iters)
    assert np.abs(res.cpu().numpy() - ref).max() <= 1e-6 * max(1.0, float(np.abs(ref).max()))


@FUZZ
@given(rows=st.integers(1, 150), width=st.integers(1, 260), iters=st.integers(0, 9), seed=st.integers(0, 2**31 - 1))
def test_min_erosion_any_shape_bit_exact(nz, oracle, rows, width, iters, seed):
    import torch
    g = grid(rows, width, seed)
    a = dev(g)
    res = nz.device.min_erosion(a, torch.empty_like(a), iters)
    assert np.array_equal(res.cpu().numpy(), oracle.min_erosion(g, iters))


@FUZZ
@given(rows=st.integers(1, 120), width=st.integers(1, 200), iters=st.integers(0, 7), seed=st.integers(0, 2**31 - 1))
def test_flowmap_any_shape(nz, oracle, rows, width, iters, seed):
    import torch
    g = grid(rows, width, seed) * np.float32(0.05)
    a = dev(g)
    need = nz.device.flowmap_scratch_bytes(width, rows, iters)
    scratch = torch.empty(max(need, 16), dtype=torch.uint8, device="cuda")
    res = nz.device.flowmap(a, torch.empty_like(a), scratch, iters, 0.0, 0.005)
    ref = oracle.flowmap(g, iters, 0.0, 0.005)
    assert np.abs(res.cpu().numpy() - ref).max() <= 1e-6 * max(1.0, float(np.abs(ref).max()))


@FUZZ
@given(rows=st.integers(1, 70), width=st.integers(1, 1200), noise_type=st.integers(0, 7), z_first=st.integers(-50, 5000),
       xpos=st.integers(-20000, 20000), zpos=st.integers(-20000, 20000), octaves=st.integers(1, 14))
def test_fractal_any_window(nz, oracle, rows, width, noise_type, z_first, xpos, zpos, octaves):
    import torch
    a = torch.empty(rows, width, device="cuda")
    nz.device.fractal(a, noise_type, 0.4, octaves=octaves, xpos=xpos, zpos=zpos, noise_size=1700, z_first=z_first)
    ref = oracle.fractal(width, rows, noise_type, 0.4, octaves=octaves, xpos=xpos, zpos=zpos, noise_size=1700, z_first=z_first)
    assert np.abs(a.cpu().numpy() - ref).max() <= 1e-6


@FUZZ
@given(res=st.integers(2, 300), iters=st.integers(0, 5), talus=st.integers(1, 89), seed=st.integers(0, 2**31 - 1))
def test_thermal_erosion_any_resolution_bit_exact(nz, oracle, res, iters, talus, seed):
    g = grid(res, res, seed)
    a = dev(g)
    nz.device.thermal_erosion(a, float(talus), 0.5, 0.75, iters)
    assert np.array_equal(a.cpu().numpy(), oracle.thermal_erosion(g, talus, 0.5, 0.75, iters))
