"""CPU-only known-answer tests that pin the oracle (oracle/noize_oracle.cpp).

The reference has no tests, so the oracle is pinned against
  * the literal tables extracted from the reference (tests/golden/kernel_tables.npz),
  * independent implementations (scipy.ndimage, closed forms),
  * structural properties of the published webgl-noise algorithms.
Parity of the noise basis functions against a real Unity/Burst run stays **unpinned**.
"""
import json
import os

import numpy as np
import pytest
import scipy.ndimage as ndi

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
RNG_SEED = 20221018  # SURVEY.md §8d micro-benchmark seed


def rand_grid(rows, width, seed=RNG_SEED):
    return np.random.default_rng(seed).random((rows, width), dtype=np.float32)


# ---------------------------------------------------------------------------------------------
# literal tables of the reference
# ---------------------------------------------------------------------------------------------
def test_gauss_tables_match_reference_literals(oracle):
    z = np.load(os.path.join(GOLDEN, "kernel_tables.npz"))
    blur = [k for k in z.files if k.startswith("blur/")]
    assert len(blur) == 16 * 12
    for k in blur:
        _, sigma, width = k.split("/")
        got = oracle.gauss_kernel(int(sigma), int(width))
        assert np.array_equal(got, z[k]), k


def test_kernel_filter_tables_match_reference_literals(oracle):
    z = np.load(os.path.join(GOLDEN, "kernel_tables.npz"))
    names = {0: "gauss9_s1", 1: "gauss7_s1", 2: "gauss5_s1", 3: "gauss3_s1", 4: "gauss9_s2", 5: "gauss7_s2",
             6: "gauss5_s2", 7: "gauss3_s2"}
    for f, n in names.items():
        kx, kz, fac = oracle.kernel_filter_table(f)
        assert np.array_equal(kx, z["kj/" + n]) and np.array_equal(kz, z["kj/" + n]) and fac == 1.0
    kx, kz, fac = oracle.kernel_filter_table(8)
    assert np.array_equal(kx, z["kj/smooth3"]) and np.float32(fac) == z["kj/smooth3Factor"][0]
    for f, (x, zz) in {9: ("sobel3_HX", "sobel3_HZ"), 10: ("sobel3_VX", "sobel3_VZ"),
                       12: ("prewitt3_HX", "prewitt3_HZ"), 13: ("prewitt3_VX", "prewitt3_VZ")}.items():
        kx, kz, fac = oracle.kernel_filter_table(f)
        assert np.array_equal(kx, z["kj/" + x]) and np.array_equal(kz, z["kj/" + zz]) and fac == 1.0
    with pytest.raises(ValueError):
        oracle.kernel_filter_table(11)


def test_limit_width(oracle):
    # BlurHelper.limitWidth, BlurKernels.cs:30-36
    assert [oracle.limit_width(w) for w in (1, 2, 3, 4, 24, 25, 26, 99)] == [3, 3, 3, 5, 25, 25, 25, 25]


def test_demo_assets_are_expressible():
    p = json.load(open(os.path.join(GOLDEN, "stage_params.json")))
    assert p["GaussHF.asset"] == {"filter": 3.0, "iterations": 3.0}      # Gauss3_S1 x3
    assert p["Sobel2D.asset"]["filter"] == 11.0                           # Sobel3_2D
    assert p["FlowMapStage.asset"]["normMax"] == 0.005
    assert p["MeshTileStage.asset"]["meshType"] == 1.0                    # Overshoot


# ---------------------------------------------------------------------------------------------
# hash arithmetic
# ---------------------------------------------------------------------------------------------
def test_mod289_float_equals_integer_mod_on_reachable_domain(oracle):
    import ctypes as C
    bad = C.c_int32(0)
    # lattice indices (non-negative tile origins) and permute's (34x+1)x with x <= 578
    assert oracle.lib.nzref_mod289_mismatches(0, 11_400_000, C.byref(bad)) == 0, bad.value
    # the float form is NOT integer mod for some negative multiples of 289 (documented quirk)
    n_neg = oracle.lib.nzref_mod289_mismatches(-200_000, -1, C.byref(bad))
    assert 0 < n_neg < 100
    assert oracle.lib.nzref_mod7_mismatches() == 0


def test_permute_is_a_permutation_polynomial_mod_289(oracle):
    vals = sorted(int(oracle.lib.nzref_permute(float(x))) for x in range(289))
    assert vals == list(range(289))


# ---------------------------------------------------------------------------------------------
# noise structure
# ---------------------------------------------------------------------------------------------
def _sample(fn, pts, *extra):
    return np.array([fn(float(x), float(y), *extra) for x, y in pts], np.float64)


def test_basis_ranges_and_continuity(oracle):
    pts = np.random.default_rng(1).uniform(-50, 50, (4000, 2)).astype(np.float32)
    L = oracle.lib
    for fn, extra, lim in ((L.nzref_snoise2, (), 1.01), (L.nzref_cnoise2, (), 1.01),
                           (L.nzref_psrnoise2, (1010.0, 102.0, 0.62), 1.01)):
        v = _sample(fn, pts, *extra)
        assert np.abs(v).max() <= lim and abs(v.mean()) < 0.05
        v2 = np.array([fn(float(x) + 1e-3, float(y), *extra) for x, y in pts])
        assert np.abs(v - v2).max() < 1e-2
    for fn in (L.nzref_snoise3, L.nzref_cnoise3):
        v = np.array([fn(float(x), float(y), 0.37 * float(x)) for x, y in pts])
        assert np.abs(v).max() <= 1.01 and abs(v.mean()) < 0.05
    f = np.array([oracle.cellular2(float(x), float(y)) for x, y in pts[:1500]])
    assert (f[:, 0] <= f[:, 1]).all() and f.min() >= 0 and f.max() < 1.5


def test_snoise_vanishes_on_lattice_points(oracle):
    # simplex corners map to v = i - (i.x+i.y)*C.x ; the surflet gradient.(0,0) is exactly 0 there
    # and the other two corners are at distance^2 >= 0.5 -> noise 0
    G2 = (3.0 - np.sqrt(3.0)) / 6.0
    for i, j in ((0, 0), (1, 0), (3, 5), (10, 2)):
        x, y = i - (i + j) * G2, j - (i + j) * G2
        assert abs(oracle.lib.nzref_snoise2(np.float32(x), np.float32(y))) < 2e-4


def test_cnoise_vanishes_on_integer_lattice(oracle):
    for i, j in ((0, 0), (1, 2), (17, 5)):
        assert abs(oracle.lib.nzref_cnoise2(float(i), float(j))) < 1e-6


def test_psrnoise_is_periodic_in_x(oracle):
    pts = np.random.default_rng(2).uniform(5, 60, (200, 2)).astype(np.float32)
    a = _sample(oracle.lib.nzref_psrnoise2, pts, 16.0, 16.0, 0.0)
    b = np.array([oracle.lib.nzref_psrnoise2(float(x) + 16.0, float(y), 16.0, 16.0, 0.0) for x, y in pts])
    assert np.abs(a - b).max() < 2e-4


def test_fractal_norm_value(oracle):
    # Fractal.cs:31-40: sum_{i<oct} G^i, G = exp2(-hurst)
    G = 2.0 ** -0.4
    assert abs(oracle.fractal_norm_value(0.4, 13) - sum(G ** i for i in range(13))) < 1e-5
    assert oracle.fractal_norm_value(0.0, 5) == 5.0


@pytest.mark.parametrize("noise_type", range(8))
def test_fractal_single_octave_is_the_basis(oracle, noise_type):
    g = oracle.fractal(16, 8, noise_type, 0.4, octaves=1, xpos=3, zpos=7, noise_size=10)
    for (z, x) in ((0, 0), (5, 11), (7, 15)):
        xi = (np.float32(x) + np.float32(3)) / np.float32(10)
        zi = (np.float32(z) + np.float32(7)) / np.float32(10)
        assert g[z, x] == np.float32(oracle.basis(noise_type, xi, zi))


def test_fractal_band_equals_slice_of_full_tile(oracle):
    full = oracle.fractal(96, 96, 3, 0.4, octaves=13, zpos=424, noise_size=1700)
    band = oracle.fractal(96, 20, 3, 0.4, octaves=13, zpos=424, noise_size=1700, z_first=40)
    assert np.array_equal(full[40:60], band)


def test_fractal_c1_statistics(oracle):
    g = oracle.fractal(256, 256, 3, 0.4, octaves=13, noise_size=1700)
    assert 0.15 < g.min() and g.max() < 0.9 and np.isfinite(g).all()


# ---------------------------------------------------------------------------------------------
# separable filter / min erosion vs scipy
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(64, 64), (33, 70), (5, 3), (1, 9)])
@pytest.mark.parametrize("ftype", [0, 2, 3, 6, 8])
def test_separable_matches_scipy_correlate(oracle, shape, ftype):
    a = rand_grid(*shape)
    kx, kz, fac = oracle.kernel_filter_table(ftype)
    got = oracle.kernel_filter(a, ftype, 3)
    ref = a.astype(np.float64)
    for _ in range(3):
        ref = ndi.correlate1d(ref, kx.astype(np.float64), axis=1, mode="nearest") * fac
        # Z operator pairs Kernel[k_off - k] with z + k (KernelOperators.cs:61-64): a correlation with the flipped kernel
        ref = ndi.correlate1d(ref, kz[::-1].astype(np.float64), axis=0, mode="nearest") * fac
    assert np.abs(got - ref).max() < 2e-6


def test_sobel_directional_and_2d(oracle):
    a = rand_grid(40, 52)
    A = oracle.kernel_filter(a, 9)
    B = oracle.kernel_filter(a, 10)
    a64 = a.astype(np.float64)
    refA = ndi.correlate1d(ndi.correlate1d(a64, [-1, 0, 1], axis=1, mode="nearest"), [1, 2, 1], axis=0, mode="nearest")
    refB = ndi.correlate1d(ndi.correlate1d(a64, [1, 2, 1], axis=1, mode="nearest"), [-1, 0, 1], axis=0, mode="nearest")
    assert np.abs(A - refA).max() < 1e-5 and np.abs(B - refB).max() < 1e-5
    S = oracle.kernel_filter(a, 11)
    assert np.abs(S - np.sqrt(refA ** 2 + refB ** 2)).max() < 1e-5


@pytest.mark.parametrize("iters", [1, 2, 5])
@pytest.mark.parametrize("shape", [(48, 48), (7, 31), (3, 2)])
def test_min_erosion_is_trailing_window_min(oracle, shape, iters):
    a = rand_grid(*shape)
    got = oracle.min_erosion(a, iters)
    n = iters + 1
    # N calls == min over the trailing (N+1)x(N+1) window [z-N..z] x [x-N..x], clamped
    pad = np.pad(a, ((iters, 0), (iters, 0)), mode="edge")
    ref = np.full_like(a, np.inf)
    for dz in range(n):
        for dx in range(n):
            ref = np.minimum(ref, pad[dz:dz + a.shape[0], dx:dx + a.shape[1]])
    assert np.array_equal(got, ref)
    sc = ndi.minimum_filter(a, size=n, mode="nearest", origin=((n - 1) // 2,) * 2)
    if shape[0] > n and shape[1] > n:
        assert np.array_equal(got, sc)


# ---------------------------------------------------------------------------------------------
# flow map
# ---------------------------------------------------------------------------------------------
def _flow_numpy(h, iters, nmin, nmax):
    """float64 re-derivation of FlowMapComponents.cs with explicit edge padding."""
    h = h.astype(np.float64)
    w = np.full_like(h, np.float64(np.float32(1e-4)))
    fW = np.zeros_like(h); fE = np.zeros_like(h); fS = np.zeros_like(h); fN = np.zeros_like(h)
    sh = lambda a, dz, dx: np.pad(a, 1, mode="edge")[1 + dz:1 + dz + a.shape[0], 1 + dx:1 + dx + a.shape[1]]
    ts = np.float64(np.float32(0.2))
    for _ in range(iters):
        H = h + w
        nW = np.maximum(0, fW + (H - sh(H, 0, -1))); nE = np.maximum(0, fE + (H - sh(H, 0, 1)))
        nS = np.maximum(0, fS + (H - sh(H, -1, 0))); nN = np.maximum(0, fN + (H - sh(H, 1, 0)))
        s = nW + nE + nS + nN
        K = np.where(s > 0, np.clip(w / np.where(s > 0, s * ts, 1), 0, 1), 0)
        fW, fE, fS, fN = nW * K, nE * K, nS * K, nN * K
        fin = sh(fE, 0, -1) + sh(fW, 0, 1) + sh(fN, -1, 0) + sh(fS, 1, 0)
        w = np.maximum(0, w + (fin - (fW + fE + fS + fN)) * ts)
    dl = sh(fE, 0, -1) - fW; dr = fE - sh(fW, 0, 1); dt = sh(fS, 1, 0) - fN; db = fS - sh(fN, -1, 0)
    v = np.sqrt(((dl + dr) * 0.5) ** 2 + ((dt + db) * 0.5) ** 2)
    return (v - np.float64(np.float32(nmin))) / (np.float64(np.float32(nmax)) - np.float64(np.float32(nmin)))


@pytest.mark.parametrize("shape,iters", [((64, 64), 5), ((20, 45), 3), ((9, 9), 1), ((2, 2), 2)])
def test_flowmap_matches_float64_rederivation(oracle, shape, iters):
    base = oracle.fractal(shape[1], shape[0], 3, 0.4, octaves=13, noise_size=170)
    smooth = oracle.kernel_filter(base, 2, 3)
    got = oracle.flowmap(smooth, iters, 0.0, 0.005)
    ref = _flow_numpy(smooth, iters, 0.0, 0.005)
    assert got.min() >= 0 and np.isfinite(got).all()
    assert np.abs(got - ref).max() < 2e-4  # values are O(1) after /0.005; float32 vs float64 of a branchy update


def test_flowmap_degenerate_range_zeroes_value(oracle):
    a = rand_grid(8, 8)
    out = oracle.flowmap(a, 1, 0.25, 0.25)   # range < 1e-12 -> v := 0 -> (0 - 0.25)/0 = -inf
    assert np.isneginf(out).all()


# ---------------------------------------------------------------------------------------------
# mesh
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mesh_type", [0, 1])
@pytest.mark.parametrize("R,in_res", [(8, 16), (29, 37), (12, 15)])
def test_mesh_indices_closed_form_and_vertices(oracle, mesh_type, R, in_res):
    h = rand_grid(in_res, in_res)
    H, T = 2000.0, 1984.375
    vtx, idx = oracle.heightmap_mesh(mesh_type, h, R, 3, H, T)
    # triangles: quad (x,z), x in 1..R, z in 1..R (SquareGridHeightMap.cs:82-105)
    zz, xx = np.meshgrid(np.arange(1, R + 1), np.arange(1, R + 1), indexing="ij")
    vi = (R + 1) * zz + xx
    tri = np.stack([vi - R - 2, vi - 1, vi - R - 1, vi - R - 1, vi - 1, vi], axis=-1).reshape(-1).astype(np.uint32)
    assert np.array_equal(idx, tri)
    assert idx.max() == (R + 1) ** 2 - 1
    off = (in_res - R) // 2
    v = vtx.reshape(R + 1, R + 1, 12)
    # position.y = h * Height at the centred crop
    assert np.allclose(v[:, :, 1], h[off:off + R + 1, off:off + R + 1] * np.float32(H), rtol=0, atol=0)
    # x/z positions
    xs = np.arange(R + 1, dtype=np.float32) * np.float32(T) / np.float32(R) - np.float32(0.5)
    assert np.array_equal(v[3, 1:, 0], xs[1:]) and v[3, 0, 0] == -(np.float32(0.5) * np.float32(T) / np.float32(R))
    assert np.array_equal(v[:, 2, 2], xs)
    # normals are unit, tangent = (-4a, 16, -4b), tangent.w = 0
    assert np.abs(np.linalg.norm(v[:, :, 3:6].astype(np.float64), axis=-1) - 1).max() < 1e-6
    assert (v[:, :, 7] == 16).all() and (v[:, :, 9] == 0).all()
    # interior normal from central differences
    z, x = R // 2, R // 2
    l, r = h[off + z, off + x - 1], h[off + z, off + x + 1]
    u, d = h[off + z - 1, off + x], h[off + z + 1, off + x]
    n = np.array([(l - r) / 2 * 8, 2 / H, (u - d) / 2 * 8], np.float64)
    assert np.abs(v[z, x, 3:6] - n / np.linalg.norm(n)).max() < 1e-6
    assert np.abs(v[z, x, 6] - (-4 * (r - l) / 2)) < 1e-6 and np.abs(v[z, x, 8] - (-4 * (u - d) / 2)) < 1e-6
    den = (R + 1.0) if mesh_type == 0 else (R - 0.5)
    assert np.abs(v[z, x, 10] - x / den) < 1e-6 and np.abs(v[z, x, 11] - z / den) < 1e-6


def test_mesh_rejects_grids_the_reference_would_overrun(oracle):
    h = rand_grid(16, 16)
    with pytest.raises(ValueError):
        oracle.heightmap_mesh(0, h, 16, 0, 10.0, 10.0)
    with pytest.raises(ValueError):
        oracle.heightmap_mesh(1, h, 16, 0, 10.0, 10.0)


def test_tile_geometry(oracle):
    # DynamicNoise.unity:530-534: tileResolution 256, tileSize 500, margin 3, generatorResolution 280
    total, mpix, tsize = oracle.tile_geometry(256, 500, 3)
    assert (total, mpix) == (258, 1) and abs(tsize - (500 + 2 * 500 / 256)) < 1e-4
    assert oracle.tile_geometry(1000, 1000, 5) == (1010, 5, 1010.0)


# ---------------------------------------------------------------------------------------------
# exact rewrites the CUDA noise kernel relies on (noize-job_b200/csrc/noise_kernels.cu)
# ---------------------------------------------------------------------------------------------
def test_round_to_nearest_equals_floor_plus_half_on_all_hash_values():
    f = np.float32
    p = np.arange(0, 290, dtype=np.float32)
    fr = p * f(0.024390243902439)
    gx = (f(2.0) * (fr - np.floor(fr)) - f(1.0)).astype(np.float32)
    magic = f(12582912.0)
    assert np.array_equal(np.floor((gx + f(0.5)).astype(np.float32)), ((gx + magic).astype(np.float32) - magic))


def test_centered_permute_is_congruent_to_the_canonical_one(oracle):
    f = np.float32
    magic, c = f(12582912.0), f(1.0) / f(289.0)
    for x in range(-300, 601):
        u = f(34 * x + 1) * f(x)                       # exact integer < 2^24
        q = f(np.float64(u) * np.float64(c) + np.float64(magic)) - magic   # fma: single rounding
        r = np.float64(u) - 289.0 * np.float64(q)     # fma(-289, q, u) is exact
        assert abs(r) <= 145 and (int(r) - int(u)) % 289 == 0
        if x >= 0:
            assert int(oracle.lib.nzref_permute(float(x))) == int(r) % 289
