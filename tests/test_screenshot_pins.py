"""The oracle against the reference's own OUTPUT: README screenshots (docs~/*.jpg) show the rendered heightmap next to
the inspector panel with every parameter of the run.  tests/golden/readme_screenshots.npz holds the cropped previews
(made by tests/golden/make_screenshot_pins.py); here the oracle is run with the panel's parameters and rank-correlated
with the render.  Image precision only (JPEG, unknown tone curve -> Spearman), but it pins what no unit test upstream
pins: which basis function `Simplex` / `Cellular` are, the fBm amplitude law, the (xpos, zpos, noiseSize) mapping,
the orientation, and the Gauss -> value-erosion -> flow-map chain.  Controls show the correlation is specific."""
import os

import numpy as np
import pytest
from scipy.stats import spearmanr

S = 250
SHOTS = np.load(os.path.join(os.path.dirname(__file__), "golden", "readme_screenshots.npz"))
PANEL = dict(hurst=0.422, octaves=13, xpos=0, zpos=424, noise_size=1757)      # docs~/3.jpg .. 6.jpg


def view(a):
    """1000^2 oracle grid -> the 250^2 preview: 4x4 box average, mirrored left-right (Unity plane UVs)."""
    n = a.shape[0] // S
    return np.fliplr(a[:S * n, :S * n].reshape(S, n, S, n).mean(axis=(1, 3)))


def rho(a, img):
    return float(spearmanr(a.ravel(), np.asarray(img, np.float32).ravel()).correlation)


@pytest.fixture(scope="module")
def chain(oracle):
    noise = oracle.fractal(1000, 1000, 3, **PANEL)
    g17 = oracle.kernel_filter(noise, 2, 17)
    er = oracle.min_erosion(g17, 5)
    return dict(noise=noise, g17=g17, er=er, flow=oracle.flowmap(er, 5, 0.0, 0.005))


def test_simplex_fbm_matches_readme_example_1_render(oracle, chain):
    gray = SHOTS["shot3"].mean(axis=2)
    r = rho(view(chain["noise"]), gray)
    assert r > 0.92, r
    # controls: other orientation, other tile position, other noise size, other basis
    assert rho(np.fliplr(view(chain["noise"])), gray) < 0.2
    assert rho(view(oracle.fractal(1000, 1000, 3, **dict(PANEL, zpos=0))), gray) < 0.6 * r
    assert rho(view(oracle.fractal(1000, 1000, 3, **dict(PANEL, noise_size=1000))), gray) < 0.6 * r
    for other in (1, 4, 5, 7):          # perlin, rotated simplex, cellular, domain-rotated simplex
        assert rho(view(oracle.fractal(1000, 1000, other, **PANEL)), gray) < 0.75 * r, other


def test_cellular_fbm_matches_readme_example_2_render(oracle):
    gray = SHOTS["shot0"].mean(axis=2)
    cell = oracle.fractal(1000, 1000, 5, 1.0, octaves=13, xpos=0, zpos=0, noise_size=1757)     # docs~/0.jpg panel
    r = rho(view(cell), gray)
    assert r > 0.98, r
    assert rho(np.fliplr(view(cell)), gray) < 0.8 * r
    assert rho(view(oracle.fractal(1000, 1000, 3, 1.0, octaves=13, xpos=0, zpos=0, noise_size=1757)), gray) < 0.8 * r


def test_gauss_erosion_flow_chain_matches_readme_example_1_render(oracle, chain):
    shot = SHOTS["shot5"].astype(np.float32)        # blue = flow map, red/green = the terrain underneath
    r_flow = rho(view(chain["flow"]), shot[:, :, 2])
    assert r_flow > 0.85, r_flow
    assert rho(view(chain["er"]), shot[:, :, 1]) > 0.85 and rho(view(chain["er"]), shot[:, :, 0]) > 0.8
    # controls: the flow map of the un-eroded terrain, or of the raw noise, fits clearly worse; terrain is not flow
    assert rho(view(oracle.flowmap(chain["g17"], 5, 0.0, 0.005)), shot[:, :, 2]) < r_flow - 0.1
    assert rho(view(oracle.flowmap(chain["noise"], 5, 0.0, 0.005)), shot[:, :, 2]) < r_flow - 0.1
    assert abs(rho(view(chain["er"]), shot[:, :, 2])) < 0.3


def test_gauss_stage_matches_readme_example_1_render(chain):
    gray = SHOTS["shot4"].mean(axis=2)
    assert rho(view(chain["g17"]), gray) > 0.89
