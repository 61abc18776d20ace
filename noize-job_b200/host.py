"""Host layer of the C ABI (nz_*): HOST buffers in, HOST buffers out — the call a reference stage makes.

Each function mirrors one static job delegate of the reference (argument order kept) and mutates the
numpy array it is given in place, exactly like the Burst jobs mutate their NativeSlice<float>.
"""
import ctypes as C
import threading
from contextlib import contextmanager

import numpy as np

from . import lib as _l


def version():
    return _l.load().nz_version().decode()


def init(device=0):
    """nz_init: `device` is one ordinal or a list of them (devices[0] runs the host layer; see set_bands)."""
    devs = [int(device)] if isinstance(device, int) else [int(d) for d in device]
    arr = (C.c_int32 * len(devs))(*devs)
    _l.check(_l.load().nz_init(arr, len(devs)))


def set_bands(n_bands):
    """nz_set_bands: host-layer stage calls on grids of >= 4096 rows run on n_bands row bands, one per device of init()."""
    _l.check(_l.load().nz_set_bands(int(n_bands)))


def test_fail_allocs(skip, count):
    """Fault injection (tests): after `skip` more device allocations the next `count` fail with NZ_E_NOMEM."""
    _l.check(_l.load().nz_test_fail_allocs(int(skip), int(count)))


def kernel_launch_count():
    return int(_l.load().nz_kernel_launch_count())


def last_timing():
    t = _l.Timing()
    _l.check(_l.load().nz_last_timing(C.byref(t)))
    return {"ms_h2d": t.ms_h2d, "ms_kernel": t.ms_kernel, "ms_d2h": t.ms_d2h, "kernel_launches": t.kernel_launches}


# ---- host-side stage logic (no GPU needed) ------------------------------------------------------
def fractal_norm_value(hurst, octaves):
    """FractalJob.CalcFractalNormValue, Noise/Fractal/Fractal.cs:31-40."""
    return float(_l.load().nz_fractal_norm_value(hurst, octaves))


def limit_width(width):
    """BlurHelper.limitWidth, Filter/Kernel/Blur/BlurKernels.cs:30-36."""
    return int(_l.load().nz_limit_width(width))


def gauss_kernel(sigma, width):
    """GaussianKernel.GetKernel, BlurKernels.cs:42-58."""
    out = np.zeros(32, np.float32)
    w = C.c_int32()
    _l.check(_l.load().nz_gauss_kernel(int(sigma), width, out.ctypes.data_as(_l._pf32), C.byref(w)))
    return out[:w.value].copy()


def kernel_filter_table(filter_type):
    """SeparableKernelFilter.Schedule's table switch, Filter/Kernel/KernelJob.cs:217-292."""
    kx, kz = np.zeros(9, np.float32), np.zeros(9, np.float32)
    ks, f = C.c_int32(), C.c_float()
    _l.check(_l.load().nz_kernel_filter_table(int(filter_type), kx.ctypes.data_as(_l._pf32), kz.ctypes.data_as(_l._pf32),
                                              C.byref(ks), C.byref(f)))
    return kx[:ks.value].copy(), kz[:ks.value].copy(), float(f.value)


def tile_geometry(tile_resolution, tile_size, margin):
    """MeshTileGenerator.calcTotalResolution/calcMarginVerts/RequestMesh, Scripts/MeshTileGenerator.cs:166-206."""
    a, b, c = C.c_int32(), C.c_int32(), C.c_float()
    _l.check(_l.load().nz_tile_geometry(tile_resolution, tile_size, margin, C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, float(c.value)


# ---- stage delegates ---------------------------------------------------------------------------
def fractal(dst, resolution, noise_type, hurst, starting_amplitude, stepdown, detune_rate, octaves, xpos, zpos,
            noise_size):
    """FractalJobDelegate, Fractal.cs:76-88."""
    _l.check(_l.load().nz_fractal(_l.as_slice(dst), resolution, int(noise_type), hurst, starting_amplitude, stepdown,
                                  detune_rate, octaves, xpos, zpos, noise_size))


def kernel_filter(src, tmp, filter_type, resolution, iterations=1):
    """SeperableKernelFilterDelegate (KernelJob.cs:308-314) x KernelFilterStage.iterations."""
    _l.check(_l.load().nz_kernel_filter(_l.as_slice(src), _l.as_slice(tmp), int(filter_type), resolution, iterations))


def separable(src, tmp, kx, kz, factor, resolution, iterations=1):
    """SeparableKernelFilter.ScheduleSeries, KernelJob.cs:165-185."""
    kx, pkx = _l.fptr(kx)
    kz, pkz = _l.fptr(kz)
    if kx.size != kz.size:
        raise ValueError("kx and kz must have the same size")
    _l.check(_l.load().nz_separable(_l.as_slice(src), _l.as_slice(tmp), kx.size, pkx, pkz, factor, resolution, iterations))


def gauss_filter(src, tmp, width, sigma, resolution, iterations=1):
    """GaussFilter.GaussFilterDelegate, Filter/Kernel/Blur/BlurJob.cs:23-30."""
    _l.check(_l.load().nz_gauss_filter(_l.as_slice(src), _l.as_slice(tmp), width, int(sigma), resolution, iterations))


def smooth_filter(src, tmp, width, resolution, iterations=1):
    """SmoothFilter.SmoothFilterDelegate, BlurJob.cs:46-52."""
    _l.check(_l.load().nz_smooth_filter(_l.as_slice(src), _l.as_slice(tmp), width, resolution, iterations))


def min_erosion(src, resolution, iterations=1):
    """ErosionKernelJobDelegate, KernelJob.cs:350."""
    _l.check(_l.load().nz_min_erosion(_l.as_slice(src), resolution, iterations))


def flowmap(height, resolution, iterations=5, norm_min=-0.1, norm_max=0.1):
    """FlowMapStage.ScheduleAll, Geologic/Stage/FlowMapStage.cs:124-195."""
    _l.check(_l.load().nz_flowmap(_l.as_slice(height), resolution, iterations, norm_min, norm_max))


def heightmap_mesh(mesh_type, vertices, indices, resolution, input_resolution, margin_pix, tile_height, tile_size,
                   heights):
    """HeightMapMeshJobScheduleDelegate, Mesh/Job/HeightMapMeshJob.cs:55-65.  `vertices` is a writable
    ((R+1)^2, 12) float32 array (48-byte Stream0 records), `indices` a writable 6*R^2 uint32 array."""
    R = resolution
    if vertices.dtype != np.float32 or vertices.size != (R + 1) * (R + 1) * 12 or not vertices.flags["C_CONTIGUOUS"]:
        raise ValueError("vertices must be a contiguous float32 array of (R+1)^2 x 12")
    if indices.dtype != np.uint32 or indices.size != 6 * R * R or not indices.flags["C_CONTIGUOUS"]:
        raise ValueError("indices must be a contiguous uint32 array of 6*R^2")
    _l.check(_l.load().nz_heightmap_mesh(int(mesh_type), vertices.ctypes.data, indices.ctypes.data, R, input_resolution,
                                         margin_pix, tile_height, tile_size, _l.as_slice(heights)))


# ---- SURVEY.md section 8f rows ---------------------------------------------------------------------
def thermal_erosion(src, talus, increment_ratio, mesh_height_width_ratio, iterations, resolution):
    """ThermalErosionFilterDelegate, Filter/Kernel/Blur/ThermalErosionFilter.cs:138-146."""
    _l.check(_l.load().nz_thermal_erosion(_l.as_slice(src), talus, increment_ratio, mesh_height_width_ratio, iterations,
                                          resolution))


def subtractive_flow_erosion(height, resolution, erosive_iterations, erosive_factor, norm_min, norm_max):
    """ErosionStageSubtractiveFlow.ScheduleAll, Geologic/Stage/ErosionStageSubtractiveFlow.cs:224-230 (cycle :138-222)."""
    _l.check(_l.load().nz_subtractive_flow_erosion(_l.as_slice(height), resolution, erosive_iterations, erosive_factor,
                                                   norm_min, norm_max))


def constant(src, tmp, operation, constant_value, resolution):
    """ConstantJobScheduleDelegate, Filter/ConstantJob.cs:49-55."""
    _l.check(_l.load().nz_constant(_l.as_slice(src), _l.as_slice(tmp), int(operation), constant_value, resolution))


def reduce(left, right, tmp, operation, resolution):
    """ReductionJobScheduleDelegate, Filter/ReductionJob.cs:55-61: left = op(left, right)."""
    _l.check(_l.load().nz_reduce(_l.as_slice(left), _l.as_slice(right), _l.as_slice(tmp), int(operation), resolution))


def curve(src, tmp, curve_samples, resolution):
    """CurveJobScheduleDelegate, Filter/Curve/CurveJob.cs:91-97."""
    _l.check(_l.load().nz_curve(_l.as_slice(src), _l.as_slice(tmp), _l.as_slice(curve_samples), resolution))


def crop(input, input_resolution, output, output_resolution, offset=0):
    """CropJobDelegate, Filter/Sample/CropJob.cs:63-69 (offset 0 == the reference, which never assigns it)."""
    _l.check(_l.load().nz_crop(_l.as_slice(input), input_resolution, _l.as_slice(output), output_resolution, offset))


def map_range(map_, lim_min=float("inf"), lim_max=float("-inf")):
    """GetMapRangeJob, Filter/NormalizeJob.cs:18-53: returns [min, max, range]."""
    res = np.zeros(3, np.float32)
    _l.check(_l.load().nz_map_range(_l.as_slice(map_), res.ctypes.data_as(_l._pf32), lim_min, lim_max))
    return res


def normalize(src, tmp, args3, resolution):
    """MapNormalizeValuesDelegate (NormalizeJob.cs:94-100) with NormalizeMap (FlowMapComponents.cs:150-166)."""
    a, pa = _l.fptr(args3)
    if a.size != 3:
        raise ValueError("args must be [min, max, range]")
    _l.check(_l.load().nz_normalize(_l.as_slice(src), _l.as_slice(tmp), pa, resolution))


# ---- residency -----------------------------------------------------------------------------------
_scope = threading.local()


def thread_in_scope():
    """True while the calling thread is inside a residency scope it opened through this module."""
    return getattr(_scope, "depth", 0) > 0 or getattr(_scope, "entered", 0) > 0


def pipeline_begin():
    """Open a residency scope on this thread.  Scopes nest: only the outermost one talks to the library, so a
    caller can keep tiles resident across several pipelines (e.g. the generator pipeline and the mesh pipeline)."""
    depth = getattr(_scope, "depth", 0)
    if depth == 0:
        _l.check(_l.load().nz_pipeline_begin())
    _scope.depth = depth + 1


def pipeline_end():
    depth = getattr(_scope, "depth", 0)
    if depth <= 1:
        _scope.depth = 0
        _l.check(_l.load().nz_pipeline_end())   # NZ_E_STATE when nothing is open
    else:
        _scope.depth = depth - 1


@contextmanager
def pipeline():
    """Keep device mirrors resident across chained stage calls; one D2H per dirty slice at exit."""
    pipeline_begin()
    try:
        yield
    finally:
        pipeline_end()


def scope_create():
    """Process-wide residency scope for chains whose stages run on different threads (see nz_scope_create)."""
    sid = int(_l.load().nz_scope_create())
    if sid < 0:
        _l.check(sid)
    return sid


def scope_enter(scope):
    _l.check(_l.load().nz_scope_enter(scope))


def scope_leave():
    _l.check(_l.load().nz_scope_leave())


def scope_close(scope):
    _l.check(_l.load().nz_scope_close(scope))


@contextmanager
def in_scope(scope):
    """Bracket one stage call made on the current thread inside `scope`."""
    scope_enter(scope)
    _scope.entered = getattr(_scope, "entered", 0) + 1
    try:
        yield
    finally:
        _scope.entered -= 1
        scope_leave()


def flush_to_host(arr):
    _l.check(_l.load().nz_flush_to_host(arr.ctypes.data))


def pin(arr):
    _l.check(_l.load().nz_pin(arr.ctypes.data, arr.nbytes))


def unpin(arr):
    _l.check(_l.load().nz_unpin(arr.ctypes.data))


# ---- named device-resident buffers (PipelineStateManager on the GPU) -------------------------------------
def context_write(name, src):
    """named buffer := src (device-to-device when src is resident in the caller's scope)."""
    _l.check(_l.load().nz_context_write(name.encode(), _l.as_slice(src)))


def context_read(name, dst):
    """dst := named buffer (device-to-device into dst's mirror)."""
    _l.check(_l.load().nz_context_read(name.encode(), _l.as_slice(dst)))


def context_exists(name):
    """Element count of the named buffer, or -1."""
    n = C.c_int32(-1)
    _l.check(_l.load().nz_context_exists(name.encode(), C.byref(n)))
    return int(n.value)


def context_download(name):
    n = context_exists(name)
    if n < 0:
        raise KeyError(name)
    out = np.empty(n, np.float32)
    _l.check(_l.load().nz_context_download(name.encode(), out.ctypes.data, n))
    return out


def context_upload(name, data):
    a = np.ascontiguousarray(data, np.float32).reshape(-1)
    _l.check(_l.load().nz_context_upload(name.encode(), a.ctypes.data, a.size))


def context_release(name):
    _l.check(_l.load().nz_context_release(name.encode()))
