"""ctypes binding of the CPU oracle (oracle/noize_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference legs).
The product package (noize-job_b200/) must never import this module.

All arrays are numpy float32, row-major (rows, width); functions return NEW arrays and never
modify their inputs, unlike the in-place reference jobs they restate.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")


def build(fast=False, force=False):
    """Compile the oracle with oracle/Makefile (parity flavour, or the -march=native timing flavour)."""
    name = "liboracle_fast.so" if fast else "liboracle.so"
    path = os.path.join(_BUILD, name)
    src = os.path.join(_HERE, "noize_oracle.cpp")
    if force or not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        if force and os.path.exists(path):
            os.remove(path)
        subprocess.run(["make", "-C", _HERE, "fast" if fast else "all"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    return path


def _bind(lib):
    i32, f32 = C.c_int32, C.c_float
    lib.nzref_fractal_norm_value.restype = f32
    lib.nzref_fractal_norm_value.argtypes = [f32, i32]
    lib.nzref_fractal.restype = i32
    lib.nzref_fractal.argtypes = [_f32p, i32, i32, i32, i32, f32, f32, f32, f32, i32, i32, i32, i32]
    for n, nargs in (("nzref_basis", None), ("nzref_snoise2", 2), ("nzref_cnoise2", 2), ("nzref_psrnoise2", 5),
                     ("nzref_snoise3", 3), ("nzref_cnoise3", 3), ("nzref_mod289", 1), ("nzref_permute", 1)):
        fn = getattr(lib, n)
        fn.restype = f32
        fn.argtypes = [i32, f32, f32] if nargs is None else [f32] * nargs
    lib.nzref_cellular2.restype = None
    lib.nzref_cellular2.argtypes = [f32, f32, _f32p]
    lib.nzref_separable.restype = i32
    lib.nzref_separable.argtypes = [_f32p, _f32p, i32, i32, i32, _f32p, _f32p, f32, i32]
    lib.nzref_kernel_filter.restype = i32
    lib.nzref_kernel_filter.argtypes = [_f32p, _f32p, i32, i32, i32, i32]
    lib.nzref_kernel_filter_table.restype = i32
    lib.nzref_kernel_filter_table.argtypes = [i32, _f32p, _f32p, C.POINTER(i32), C.POINTER(f32)]
    lib.nzref_gauss_kernel.restype = i32
    lib.nzref_gauss_kernel.argtypes = [i32, i32, _f32p, C.POINTER(i32)]
    lib.nzref_limit_width.restype = i32
    lib.nzref_limit_width.argtypes = [i32]
    lib.nzref_min_erosion.restype = i32
    lib.nzref_min_erosion.argtypes = [_f32p, _f32p, i32, i32, i32]
    lib.nzref_flowmap.restype = i32
    lib.nzref_flowmap.argtypes = [_f32p, i32, i32, i32, f32, f32]
    lib.nzref_subtractive_flow_erosion.restype = i32
    lib.nzref_subtractive_flow_erosion.argtypes = [_f32p, i32, i32, i32, f32, f32, f32]
    lib.nzref_heightmap_mesh.restype = i32
    lib.nzref_heightmap_mesh.argtypes = [i32, _f32p, _u32p, i32, i32, i32, f32, f32, _f32p]
    lib.nzref_tile_geometry.restype = i32
    lib.nzref_tile_geometry.argtypes = [i32, i32, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(f32)]
    i64 = C.c_int64
    lib.nzref_thermal_max_diff.restype = f32
    lib.nzref_thermal_max_diff.argtypes = [f32, f32, i32]
    lib.nzref_thermal_erosion.restype = i32
    lib.nzref_thermal_erosion.argtypes = [_f32p, i32, f32, f32, f32, i32]
    lib.nzref_constant.restype = i32
    lib.nzref_constant.argtypes = [_f32p, i64, i32, f32]
    lib.nzref_reduce.restype = i32
    lib.nzref_reduce.argtypes = [_f32p, _f32p, i64, i32]
    lib.nzref_curve.restype = i32
    lib.nzref_curve.argtypes = [_f32p, i64, _f32p, i32]
    lib.nzref_crop.restype = i32
    lib.nzref_crop.argtypes = [_f32p, i32, _f32p, i32, i32]
    lib.nzref_map_range.restype = i32
    lib.nzref_map_range.argtypes = [_f32p, i64, f32, f32, _f32p]
    lib.nzref_normalize.restype = i32
    lib.nzref_normalize.argtypes = [_f32p, i64, _f32p]
    lib.nzref_num_threads.restype = i32
    lib.nzref_set_num_threads.restype = i32
    lib.nzref_set_num_threads.argtypes = [i32]
    lib.nzref_mod289_mismatches.restype = C.c_int64
    lib.nzref_mod289_mismatches.argtypes = [i32, i32, C.POINTER(i32)]
    lib.nzref_mod7_mismatches.restype = i32
    return lib


class Oracle:
    def __init__(self, fast=False):
        self.lib = _bind(C.CDLL(build(fast=fast)))
        self.fast = fast

    # -- noise -------------------------------------------------------------------------------
    def fractal_norm_value(self, hurst, octaves):
        return float(self.lib.nzref_fractal_norm_value(hurst, octaves))

    def fractal(self, width, rows, noise_type, hurst, starting_amplitude=1.0, stepdown=2.0, detune_rate=0.0,
                octaves=1, xpos=0, zpos=0, noise_size=1000, z_first=0):
        out = np.empty((rows, width), np.float32)
        rc = self.lib.nzref_fractal(out, width, rows, z_first, int(noise_type), hurst, starting_amplitude, stepdown,
                                    detune_rate, octaves, xpos, zpos, noise_size)
        assert rc == 0, rc
        return out

    def basis(self, noise_type, x, z):
        return float(self.lib.nzref_basis(int(noise_type), x, z))

    def cellular2(self, x, y):
        f = np.zeros(2, np.float32)
        self.lib.nzref_cellular2(x, y, f)
        return f

    # -- filters -----------------------------------------------------------------------------
    @staticmethod
    def _grid(a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        assert a.ndim == 2
        return a.copy()

    def separable(self, grid, kx, kz, factor=1.0, iterations=1):
        d = self._grid(grid)
        kx = np.ascontiguousarray(kx, np.float32)
        kz = np.ascontiguousarray(kz, np.float32)
        assert kx.size == kz.size and kx.size % 2 == 1
        rc = self.lib.nzref_separable(d, np.empty_like(d), d.shape[1], d.shape[0], kx.size, kx, kz, factor, iterations)
        assert rc == 0, rc
        return d

    def kernel_filter(self, grid, filter_type, iterations=1):
        d = self._grid(grid)
        rc = self.lib.nzref_kernel_filter(d, np.empty_like(d), d.shape[1], d.shape[0], int(filter_type), iterations)
        assert rc == 0, rc
        return d

    def kernel_filter_table(self, filter_type):
        kx, kz = np.zeros(9, np.float32), np.zeros(9, np.float32)
        ks, f = C.c_int32(), C.c_float()
        rc = self.lib.nzref_kernel_filter_table(int(filter_type), kx, kz, C.byref(ks), C.byref(f))
        if rc != 0:
            raise ValueError(rc)
        return kx[:ks.value].copy(), kz[:ks.value].copy(), float(f.value)

    def gauss_kernel(self, sigma, width):
        out = np.zeros(32, np.float32)
        w = C.c_int32()
        rc = self.lib.nzref_gauss_kernel(int(sigma), width, out, C.byref(w))
        assert rc == 0, rc
        return out[:w.value].copy()

    def limit_width(self, w):
        return int(self.lib.nzref_limit_width(w))

    def gauss_filter(self, grid, width, sigma, iterations=1):
        k = self.gauss_kernel(sigma, width)
        return self.separable(grid, k, k, 1.0, iterations)

    def smooth_filter(self, grid, width, iterations=1):
        w = self.limit_width(width)
        k = np.full(w, np.float32(1.0) / np.float32(w), np.float32)
        return self.separable(grid, k, k, 1.0, iterations)

    def min_erosion(self, grid, iterations=1):
        d = self._grid(grid)
        rc = self.lib.nzref_min_erosion(d, np.empty_like(d), d.shape[1], d.shape[0], iterations)
        assert rc == 0, rc
        return d

    def flowmap(self, grid, iterations=5, norm_min=-0.1, norm_max=0.1):
        d = self._grid(grid)
        rc = self.lib.nzref_flowmap(d, d.shape[1], d.shape[0], iterations, norm_min, norm_max)
        assert rc == 0, rc
        return d

    # -- section 8f rows: thermal erosion, subtractive-flow erosion and the element-wise stages --
    def subtractive_flow_erosion(self, grid, erosive_iterations=5, erosive_factor=0.1, norm_min=-0.1, norm_max=0.1):
        d = self._grid(grid)
        rc = self.lib.nzref_subtractive_flow_erosion(d, d.shape[1], d.shape[0], erosive_iterations, erosive_factor, norm_min, norm_max)
        assert rc == 0, rc
        return d

    def thermal_max_diff(self, talus, height_ratio, resolution):
        return float(self.lib.nzref_thermal_max_diff(talus, height_ratio, resolution))

    def thermal_erosion(self, grid, talus=45.0, increment=0.5, mesh_height_width_ratio=0.75, iterations=1):
        d = self._grid(grid)
        assert d.shape[0] == d.shape[1]
        rc = self.lib.nzref_thermal_erosion(d, d.shape[0], talus, increment, mesh_height_width_ratio, iterations)
        assert rc == 0, rc
        return d

    def constant(self, grid, op, value):
        d = self._grid(grid)
        rc = self.lib.nzref_constant(d, d.size, int(op), value)
        assert rc == 0, rc
        return d

    def reduce(self, left, right, op):
        d = self._grid(left)
        r = np.ascontiguousarray(right, np.float32)
        assert r.shape == d.shape
        rc = self.lib.nzref_reduce(d, r, d.size, int(op))
        assert rc == 0, rc
        return d

    def curve(self, grid, curve):
        d = self._grid(grid)
        c = np.ascontiguousarray(curve, np.float32)
        rc = self.lib.nzref_curve(d, d.size, c, c.size)
        assert rc == 0, rc
        return d

    def crop(self, grid, out_res, offset=0):
        g = np.ascontiguousarray(grid, np.float32)
        assert g.shape[0] == g.shape[1]
        out = np.empty((out_res, out_res), np.float32)
        rc = self.lib.nzref_crop(g, g.shape[0], out, out_res, offset)
        assert rc == 0, rc
        return out

    def map_range(self, grid, lim_min=np.inf, lim_max=-np.inf):
        g = np.ascontiguousarray(grid, np.float32)
        res = np.zeros(3, np.float32)
        rc = self.lib.nzref_map_range(g.reshape(-1), g.size, lim_min, lim_max, res)
        assert rc == 0, rc
        return res

    def normalize(self, grid, args3):
        d = self._grid(grid)
        rc = self.lib.nzref_normalize(d, d.size, np.ascontiguousarray(args3, np.float32))
        assert rc == 0, rc
        return d

    # -- mesh --------------------------------------------------------------------------------
    def heightmap_mesh(self, mesh_type, heights, resolution, margin_pix, tile_height, tile_size):
        h = np.ascontiguousarray(heights, np.float32)
        in_res = h.shape[0]
        assert h.shape == (in_res, in_res)
        R = resolution
        vtx = np.zeros(((R + 1) * (R + 1), 12), np.float32)
        idx = np.zeros(6 * R * R, np.uint32)
        rc = self.lib.nzref_heightmap_mesh(int(mesh_type), vtx, idx, R, in_res, margin_pix, tile_height, tile_size, h)
        if rc != 0:
            raise ValueError(rc)
        return vtx, idx

    def tile_geometry(self, tile_resolution, tile_size, margin):
        a, b, c = C.c_int32(), C.c_int32(), C.c_float()
        rc = self.lib.nzref_tile_geometry(tile_resolution, tile_size, margin, C.byref(a), C.byref(b), C.byref(c))
        assert rc == 0, rc
        return a.value, b.value, float(c.value)

    def num_threads(self):
        return int(self.lib.nzref_num_threads())

    def set_num_threads(self, n):
        """OpenMP worker count for every later call (overrides OMP_NUM_THREADS); returns the count now in force."""
        return int(self.lib.nzref_set_num_threads(int(n)))


_cache = {}


def get(fast=False):
    if fast not in _cache:
        _cache[fast] = Oracle(fast=fast)
    return _cache[fast]
