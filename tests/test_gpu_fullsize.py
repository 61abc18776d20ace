"""BASELINE config C5 at FULL size (one 16384^2 heightmap, the bench workload) checked through size-independent
properties: every stage is local (halo 34 / 11 / 5 / 1 cells), so the full-grid GPU result restricted to a window
must equal the oracle run on that window plus its halo — in the interior, at the grid corner (clamp-to-edge), and
for the mesh's closed-form index stream.  Tolerances as in test_gpu_parity.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 16384
C5 = dict(hurst=0.4, octaves=13, noise_size=1700)
HALO = {"filter": 34, "flow": 11, "erosion": 5}


@pytest.fixture(scope="module")
def chain(nz):
    """Runs the C5 chain once on the device layer and keeps a CPU copy of the stage outputs on two windows."""
    import torch
    if torch.cuda.mem_get_info()[0] < 30e9:
        pytest.skip("needs ~25 GB of device memory")
    d = nz.device
    a = torch.empty(N, N, device="cuda")
    b = torch.empty_like(a)
    wins = {"interior": (9000, 5000), "corner": (0, 0), "far_corner": (N - 294, N - 294)}
    W = 294                                             # 192 + 2 * (34 + 11 + 5 + 1)
    out = {k: {} for k in wins}

    def grab(stage, t):
        for k, (z0, x0) in wins.items():
            out[k][stage] = t[z0:z0 + W, x0:x0 + W].cpu().numpy()

    d.fractal(a, 3, C5["hurst"], octaves=C5["octaves"], noise_size=C5["noise_size"])
    grab("noise", a)
    cur = d.kernel_filter(a, b, 2, 17)
    other = b if cur is a else a
    grab("filter", cur)
    res = d.flowmap(cur, other, None, 5, 0.0, 0.005)
    if res is not cur:
        cur, other = other, cur
    grab("flow", cur)
    res = d.min_erosion(cur, other, 5)
    if res is not cur:
        cur, other = other, cur
    grab("erosion", cur)
    R = N - 8
    vtx = torch.empty((R + 1) * (R + 1), 12, device="cuda")
    idx = torch.empty(6 * R * R, dtype=torch.int32, device="cuda")
    d.heightmap_mesh(1, vtx, idx, R, N, 4, 2000.0, R * (500.0 / 256.0), cur)
    torch.cuda.synchronize()
    rng = np.random.default_rng(5)
    tri_rows = np.unique(np.concatenate([[0, R - 1], rng.integers(0, R, 6)]))
    mesh = {"R": R, "tri": {int(z): idx[6 * R * z: 6 * R * (z + 1)].cpu().numpy().view(np.uint32) for z in tri_rows},
            "vrow": {int(z): vtx[(R + 1) * z: (R + 1) * (z + 1)].cpu().numpy() for z in (0, 7000, R)},
            "hrows": {int(z): cur[z + 4 - 1: z + 4 + 2].cpu().numpy() for z in (0, 7000, R)}}
    return wins, W, out, mesh


def oracle_window(oracle, z0, x0, W):
    """The chain on the W x W window whose top-left cell is (z0, x0) of the big grid (float positions are exact)."""
    st = {"noise": oracle.fractal(W, W, 3, C5["hurst"], octaves=C5["octaves"], xpos=x0, zpos=z0, noise_size=C5["noise_size"])}
    st["filter"] = oracle.kernel_filter(st["noise"], 2, 17)
    st["flow"] = oracle.flowmap(st["filter"], 5, 0.0, 0.005)
    st["erosion"] = oracle.min_erosion(st["flow"], 5)
    return st


@pytest.mark.parametrize("where", ["interior", "corner", "far_corner"])
def test_full_size_chain_equals_oracle_on_windows(chain, oracle, where):
    wins, W, out, _ = chain
    z0, x0 = wins[where]
    ref = oracle_window(oracle, z0, x0, W)
    got = out[where]
    assert np.abs(got["noise"] - ref["noise"]).max() <= 1e-6
    # a cut edge of the window (not a grid edge) invalidates `halo` cells per stage; a grid edge is exact (clamp)
    lo_z = lo_x = 0
    hi_z = hi_x = W
    tol = {"filter": 1e-6, "flow": 1e-6, "erosion": 1e-6}
    for stage in ("filter", "flow", "erosion"):
        h = HALO[stage]
        if z0 > 0: lo_z += h
        if x0 > 0: lo_x += h
        if z0 + W < N: hi_z -= h
        if x0 + W < N: hi_x -= h
        g, r = got[stage][lo_z:hi_z, lo_x:hi_x], ref[stage][lo_z:hi_z, lo_x:hi_x]
        assert g.shape[0] >= 192 and g.shape[1] >= 192
        assert np.abs(g - r).max() <= tol[stage] * max(1.0, float(np.abs(r).max())), stage


def test_full_size_mesh_indices_closed_form_and_vertex_rows(chain, oracle):
    _, _, _, mesh = chain
    R = mesh["R"]
    for z, tri in mesh["tri"].items():
        # SquareGrid / Overshoot triangles of quad row z (vertex rows z, z+1): vi = (R+1)*(z+1) + x + 1 is the quad's far
        # corner; (vi-R-2, vi-1, vi-R-1) and (vi-R-1, vi-1, vi) — OvershootSquareGridHeightMap.cs:88-97
        x = np.arange(R, dtype=np.int64)
        vi = (R + 1) * (z + 1) + x + 1
        want = np.stack([vi - R - 2, vi - 1, vi - R - 1, vi - R - 1, vi - 1, vi], 1).reshape(-1).astype(np.uint32)
        assert np.array_equal(tri, want), z
    f32 = np.float32
    for z, v in mesh["vrow"].items():
        # OvershootSquareGridHeightMap.SetVertexValues (:62-75) recomputed from the three height rows around z + off
        h = mesh["hrows"][z]
        assert h.shape[0] == 3
        xs = np.arange(R + 1) + 4
        t, l, r, u, d = h[1, xs], h[1, xs - 1], h[1, xs + 1], h[0, xs], h[2, xs]
        assert np.isfinite(v).all()
        assert np.array_equal(v[:, 1], t * f32(2000.0))                                  # position.y = height * tileHeight
        nx, ny, nz_ = (l - r) / f32(2) * f32(8), f32(2.0) / f32(2000.0), (u - d) / f32(2) * f32(8)
        inv = 1.0 / np.sqrt(nx.astype(np.float64) ** 2 + float(ny) ** 2 + nz_.astype(np.float64) ** 2)
        assert np.abs(v[:, 3] - nx * inv).max() < 1e-5 and np.abs(v[:, 5] - nz_ * inv).max() < 1e-5
        assert np.abs(np.linalg.norm(v[:, 3:6].astype(np.float64), axis=1) - 1).max() < 1e-5   # unit normals
        assert np.abs(v[:, 6] - (-f32(4.0) * ((r - l) / f32(2)))).max() < 1e-6 and np.all(v[:, 7] == 16.0) and np.all(v[:, 9] == 0.0)
        assert np.abs(v[:, 10] - np.arange(R + 1, dtype=np.float32) / f32(R - 0.5)).max() < 1e-6   # uv.x
        assert np.abs(v[:, 11] - f32(z) / f32(R - 0.5)).max() < 1e-6                               # uv.y
