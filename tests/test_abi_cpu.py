"""CPU-only checks of the drop-in boundary: libnoize_b200.so loads, exports every symbol the header
declares, its host-side stage logic matches the reference's literal tables, and compute entry points
fail loudly (never fall back) when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def header_symbols():
    src = open(os.path.join(ROOT, "include", "noize_b200.h")).read()
    return re.findall(r"NZ_API\s+[\w\s\*]+?\b(nz_\w+)\s*\(", src)


def test_library_exports_every_declared_symbol(nz):
    names = header_symbols()
    assert len(names) >= 30 and len(set(names)) == len(names)
    lib = C.CDLL(nz.lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # and the Python binding binds exactly that set
    assert sorted(names) == sorted(nz.lib.SIGNATURES)


def test_header_enums_match_python_mirror(nz):
    src = open(os.path.join(ROOT, "include", "noize_b200.h")).read()
    for enum, prefix in ((nz.FractalNoise, "NZ_NOISE_"), (nz.KernelFilterType, "NZ_FILTER_")):
        vals = dict(re.findall(rf"({prefix}\w+)\s*=\s*(\d+)", src))
        vals = {k: int(v) for k, v in vals.items() if not k.endswith("__COUNT")}
        assert sorted(vals.values()) == [int(e) for e in enum]
    assert int(nz.MeshType.OvershootSquareGridHeightMap) == 1 and int(nz.GaussSigma.s8d00) == 15


def test_product_never_references_the_oracle():
    pkg = os.path.join(ROOT, "noize-job_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".cs")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "nzref_" not in txt and "import oracle" not in txt and "liboracle" not in txt, os.path.join(dirpath, f)


def test_gauss_tables_match_reference_literals(nz):
    z = np.load(os.path.join(GOLDEN, "kernel_tables.npz"))
    for k in (k for k in z.files if k.startswith("blur/")):
        _, sigma, width = k.split("/")
        assert np.array_equal(nz.host.gauss_kernel(int(sigma), int(width)), z[k]), k
    for f, n in enumerate(["gauss9_s1", "gauss7_s1", "gauss5_s1", "gauss3_s1", "gauss9_s2", "gauss7_s2", "gauss5_s2", "gauss3_s2"]):
        kx, kz, fac = nz.host.kernel_filter_table(f)
        assert np.array_equal(kx, z["kj/" + n]) and np.array_equal(kz, z["kj/" + n]) and fac == 1.0
    kx, kz, fac = nz.host.kernel_filter_table(nz.KernelFilterType.Smooth3)
    assert np.array_equal(kx, z["kj/smooth3"]) and np.float32(fac) == z["kj/smooth3Factor"][0]
    for f, (x, zz) in {9: ("sobel3_HX", "sobel3_HZ"), 10: ("sobel3_VX", "sobel3_VZ"), 12: ("prewitt3_HX", "prewitt3_HZ"),
                       13: ("prewitt3_VX", "prewitt3_VZ")}.items():
        kx, kz, fac = nz.host.kernel_filter_table(f)
        assert np.array_equal(kx, z["kj/" + x]) and np.array_equal(kz, z["kj/" + zz]) and fac == 1.0
    with pytest.raises(nz.NzError) as e:
        nz.host.kernel_filter_table(11)
    assert e.value.code == nz.lib.NZ_E_UNSUPPORTED


def test_host_logic_agrees_with_oracle(nz, oracle):
    for hurst, octv in ((0.4, 13), (0.9001, 6), (0.0, 1), (2.0, 24)):
        assert nz.host.fractal_norm_value(hurst, octv) == oracle.fractal_norm_value(hurst, octv)
    assert [nz.host.limit_width(w) for w in range(0, 30)] == [oracle.limit_width(w) for w in range(0, 30)]
    for args in ((256, 500, 3), (1000, 1000, 5), (1024, 2000, 7)):
        assert nz.host.tile_geometry(*args) == oracle.tile_geometry(*args)


@pytest.mark.skipif(torch.cuda.is_available(), reason="GPU present: covered by the gpu tests")
def test_compute_fails_loudly_without_a_device(nz):
    with pytest.raises(nz.NzError) as e:
        nz.host.fractal(np.zeros(16, np.float32), 4, 3, 0.4, 1.0, 2.0, 0.0, 13, 0, 0, 1700)
    assert e.value.code == nz.lib.NZ_E_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(nz.NzError):
        nz.host.pipeline_begin()


def test_argument_validation_precedes_device_use(nz):
    # bad arguments are rejected with NZ_E_INVALID even on a machine without a GPU
    for call in (lambda: nz.host.fractal(np.zeros(10, np.float32), 4, 3, 0.4, 1.0, 2.0, 0.0, 13, 0, 0, 1700),
                 lambda: nz.host.fractal(np.zeros(16, np.float32), 4, 3, 0.4, 1.0, 2.0, 0.0, 0, 0, 0, 1700),
                 lambda: nz.host.fractal(np.zeros(16, np.float32), 4, 8, 0.4, 1.0, 2.0, 0.0, 13, 0, 0, 1700),
                 lambda: nz.host.kernel_filter(np.zeros(16, np.float32), None, 14, 4, 1),
                 lambda: nz.host.min_erosion(np.zeros(15, np.float32), 4, 1)):
        with pytest.raises(nz.NzError) as e:
            call()
        assert e.value.code == nz.lib.NZ_E_INVALID
    with pytest.raises(TypeError):
        nz.host.fractal(np.zeros(16, np.float64), 4, 3, 0.4, 1.0, 2.0, 0.0, 13, 0, 0, 1700)


def test_stage_mirror_keeps_reference_defaults(nz):
    # Noise/NoiseStage.cs:37-54, Filter/KernelFilterStage.cs:17-19, Geologic/Stage/FlowMapStage.cs:18-23
    n = nz.NoiseStage()
    assert (n.hurst, n.startingAmplitude, n.octaves, n.stepdown, n.detuneRate, n.noiseSize) == (0.0, 1.0, 1, 2.0, 0.0, 1000)
    assert nz.KernelFilterStage().iterations == 1
    f = nz.FlowMapStage()
    assert (f.iterations, f.normMin, f.normMax) == (5, -0.1, 0.1)
    assert nz.MeshTileStage().meshType == nz.MeshType.SquareGridHeightMap
    d = nz.GeneratorData()
    assert (d.resolution, d.xpos, d.zpos) == (512, 0, 0)
    m = nz.MeshStageData()
    assert (m.resolution, m.inputResolution, m.marginPix, m.tileSize, m.tileHeight) == (512, 512, 5, 512.0, 512.0)
    with pytest.raises(Exception, match="No stages"):
        nz.BasePipeline([])


def test_cpp_host_mirror_is_built_and_fails_loudly_without_a_gpu(nz):
    import subprocess
    exe = os.path.join(os.path.dirname(nz.lib.LIB_PATH), "host_cpp", "example_chain")
    assert os.path.exists(exe), "build() compiles host_cpp/example_chain from noize_stages.hpp"
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    out = subprocess.run([exe, "64"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 1 and "no CPU fallback" in out.stderr


def test_csharp_bindings_name_exported_symbols_and_mirror_the_structs(nz):
    """The C# interop files cannot be compiled here (no Unity / mono), so keep them in step with the header mechanically:
    every function behind a [DllImport] is a symbol the library exports, and the blittable structs have as many 32-bit /
    pointer fields as their C counterparts."""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = nz.load()
    names = set()
    for path in glob.glob(os.path.join(root, "noize-job_b200", "unity", "Interop", "*.cs")):
        src = open(path).read()
        names |= set(re.findall(r"\[DllImport\(LIB\)\]\s*public static extern [^;(]*?\b(nz_[a-z0-9_]+)\s*\(", src))
    assert len(names) >= 45, sorted(names)
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    header = open(os.path.join(root, "include", "noize_b200.h")).read()

    def c_fields(struct_name):
        end = header.index("} " + struct_name + ";")
        body = header[header.rindex("typedef struct {", 0, end) + len("typedef struct {"):end]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        return sum(len(decl.split(",")) for decl in body.split(";") if decl.strip())

    def cs_fields(struct_name):
        src = open(os.path.join(root, "noize-job_b200", "unity", "Interop", "NoizeB200Worlds.cs")).read()
        body = re.search(r"struct " + struct_name + r" \{(.*?)\n    \}", src, re.S).group(1)
        body = re.sub(r"//.*", "", body)
        return sum(len(decl.split(",")) for decl in body.split(";") if decl.strip())

    for c, cs in (("nz_tile_config", "NzTileConfig"), ("nz_chain_config", "NzChainConfig"), ("nz_band_info", "NzBandInfo")):
        assert c_fields(c) == cs_fields(cs), (c, c_fields(c), cs_fields(cs))
