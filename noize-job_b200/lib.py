"""ctypes binding of libnoize_b200.so — the C ABI declared in include/noize_b200.h.

This module is plumbing: it loads the CUDA library, converts numpy arrays / device pointers to the
C argument types and turns negative status codes into exceptions.  There is no Python or CPU
implementation of any stage here; if the shared library is missing, importing `load()` fails loudly.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnoize_b200.so")
CSRC = os.path.join(_HERE, "csrc")

NZ_OK, NZ_E_INVALID, NZ_E_CUDA, NZ_E_NOMEM, NZ_E_STATE, NZ_E_UNSUPPORTED = 0, -1, -2, -3, -4, -5
_CODE_NAMES = {-1: "NZ_E_INVALID", -2: "NZ_E_CUDA", -3: "NZ_E_NOMEM", -4: "NZ_E_STATE", -5: "NZ_E_UNSUPPORTED"}


class NzError(RuntimeError):
    """A C-ABI call returned a negative status (the C# wrapper throws at schedule time likewise)."""

    def __init__(self, code, message):
        super().__init__(f"{_CODE_NAMES.get(code, code)}: {message}")
        self.code = code


class Slice(C.Structure):
    """nz_slice_f32 == Unity's NativeSlice<float>: (byte* ptr, int stride, int length)."""
    _fields_ = [("ptr", C.c_void_p), ("stride_bytes", C.c_int32), ("length", C.c_int32)]


class ChainConfig(C.Structure):
    """nz_chain_config: the BASELINE C5 chain on one resolution^2 heightmap."""
    _fields_ = [("resolution", C.c_int32),
                ("noise_type", C.c_int32), ("hurst", C.c_float), ("starting_amplitude", C.c_float), ("stepdown", C.c_float),
                ("detune_rate", C.c_float), ("octaves", C.c_int32), ("xpos", C.c_int32), ("zpos", C.c_int32), ("noise_size", C.c_int32),
                ("filter_type", C.c_int32), ("filter_iterations", C.c_int32),
                ("flow_iterations", C.c_int32), ("norm_min", C.c_float), ("norm_max", C.c_float),
                ("erosion_iterations", C.c_int32),
                ("mesh_type", C.c_int32), ("mesh_resolution", C.c_int32), ("mesh_margin_pix", C.c_int32),
                ("tile_height", C.c_float), ("tile_size", C.c_float)]


class BandInfo(C.Structure):
    """nz_band_info"""
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("device", C.c_int32), ("z0", C.c_int32), ("z1", C.c_int32),
                ("vz0", C.c_int32), ("vz1", C.c_int32), ("d_rows", C.c_void_p), ("d_vertices", C.c_void_p),
                ("d_indices", C.c_void_p), ("halo_bytes_per_run", C.c_int64)]


class TileConfig(C.Structure):
    """nz_tile_config: the per-tile chain of a tiled world (BASELINE config C4)."""
    _fields_ = [("resolution", C.c_int32), ("tile_resolution", C.c_int32),
                ("noise_type", C.c_int32), ("hurst", C.c_float), ("starting_amplitude", C.c_float), ("stepdown", C.c_float),
                ("detune_rate", C.c_float), ("octaves", C.c_int32), ("noise_size", C.c_int32),
                ("filter_type", C.c_int32), ("filter_iterations", C.c_int32),
                ("edge_filter_type", C.c_int32), ("edge_filter_iterations", C.c_int32),
                ("mesh_type", C.c_int32), ("mesh_resolution", C.c_int32), ("mesh_margin_pix", C.c_int32),
                ("tile_height", C.c_float), ("tile_size", C.c_float)]


class Timing(C.Structure):
    _fields_ = [("ms_h2d", C.c_float), ("ms_kernel", C.c_float), ("ms_d2h", C.c_float), ("kernel_launches", C.c_int32)]


# every symbol include/noize_b200.h declares: name -> (restype, argtypes)
_i32, _f32, _vp, _sz = C.c_int32, C.c_float, C.c_void_p, C.c_size_t
_pi32, _pf32 = C.POINTER(C.c_int32), C.POINTER(C.c_float)
SIGNATURES = {
    "nz_init": (_i32, [_pi32, _i32]),
    "nz_shutdown": (_i32, []),
    "nz_last_error": (C.c_char_p, []),
    "nz_version": (C.c_char_p, []),
    "nz_kernel_launch_count": (C.c_int64, []),
    "nz_last_timing": (_i32, [C.POINTER(Timing)]),
    "nz_test_fail_allocs": (_i32, [_i32, _i32]),
    "nz_set_bands": (_i32, [_i32]),
    "nz_context_write": (_i32, [C.c_char_p, Slice]),
    "nz_context_read": (_i32, [C.c_char_p, Slice]),
    "nz_context_exists": (_i32, [C.c_char_p, _pi32]),
    "nz_context_download": (_i32, [C.c_char_p, _vp, _i32]),
    "nz_context_upload": (_i32, [C.c_char_p, _vp, _i32]),
    "nz_context_release": (_i32, [C.c_char_p]),
    "nz_band_geometry": (_i32, [_i32, _i32, _i32, _i32, C.POINTER(BandInfo)]),
    "nz_comm_unique_id": (_i32, [_vp, _i32]),
    "nz_comm_create": (C.c_int64, [_vp, _i32, _i32, _i32]),
    "nz_comm_async_error": (_i32, [C.c_int64]),
    "nz_comm_destroy": (_i32, [C.c_int64]),
    "nz_band_chain_create": (C.c_int64, [C.POINTER(ChainConfig), C.c_int64, _i32, _i32, _vp]),
    "nz_band_chain_create_local": (C.c_int64, [C.POINTER(ChainConfig), _pi32, _i32, _i32]),
    "nz_band_chain_run": (_i32, [C.c_int64, _i32]),
    "nz_band_chain_sync": (_i32, [C.c_int64]),
    "nz_band_chain_stage_ms": (_i32, [C.c_int64, _pf32]),
    "nz_band_chain_local_bands": (_i32, [C.c_int64]),
    "nz_band_chain_info": (_i32, [C.c_int64, _i32, C.POINTER(BandInfo)]),
    "nz_band_chain_download": (_i32, [C.c_int64, _vp, _vp, _vp]),
    "nz_band_chain_destroy": (_i32, [C.c_int64]),
    "nz_tile_world_create": (C.c_int64, [C.POINTER(TileConfig), _i32, _i32]),
    "nz_tile_world_run": (_i32, [C.c_int64, _pi32, _i32, _vp, _vp, _vp, _vp]),
    "nz_tile_world_slot": (_i32, [C.c_int64, _i32, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "nz_tile_world_destroy": (_i32, [C.c_int64]),
    "nz_fractal_norm_value": (_f32, [_f32, _i32]),
    "nz_gauss_kernel": (_i32, [_i32, _i32, _pf32, _pi32]),
    "nz_limit_width": (_i32, [_i32]),
    "nz_kernel_filter_table": (_i32, [_i32, _pf32, _pf32, _pi32, _pf32]),
    "nz_tile_geometry": (_i32, [_i32, _i32, _i32, _pi32, _pi32, _pf32]),
    "nz_fractal": (_i32, [Slice, _i32, _i32, _f32, _f32, _f32, _f32, _i32, _i32, _i32, _i32]),
    "nz_kernel_filter": (_i32, [Slice, Slice, _i32, _i32, _i32]),
    "nz_separable": (_i32, [Slice, Slice, _i32, _pf32, _pf32, _f32, _i32, _i32]),
    "nz_gauss_filter": (_i32, [Slice, Slice, _i32, _i32, _i32, _i32]),
    "nz_smooth_filter": (_i32, [Slice, Slice, _i32, _i32, _i32]),
    "nz_min_erosion": (_i32, [Slice, _i32, _i32]),
    "nz_flowmap": (_i32, [Slice, _i32, _i32, _f32, _f32]),
    "nz_heightmap_mesh": (_i32, [_i32, _vp, _vp, _i32, _i32, _i32, _f32, _f32, Slice]),
    "nz_thermal_erosion": (_i32, [Slice, _f32, _f32, _f32, _i32, _i32]),
    "nz_subtractive_flow_erosion": (_i32, [Slice, _i32, _i32, _f32, _f32, _f32]),
    "nz_constant": (_i32, [Slice, Slice, _i32, _f32, _i32]),
    "nz_reduce": (_i32, [Slice, Slice, Slice, _i32, _i32]),
    "nz_curve": (_i32, [Slice, Slice, Slice, _i32]),
    "nz_crop": (_i32, [Slice, _i32, Slice, _i32, _i32]),
    "nz_map_range": (_i32, [Slice, _pf32, _f32, _f32]),
    "nz_normalize": (_i32, [Slice, Slice, _pf32, _i32]),
    "nz_pipeline_begin": (_i32, []),
    "nz_pipeline_end": (_i32, []),
    "nz_scope_create": (C.c_int64, []),
    "nz_scope_enter": (_i32, [C.c_int64]),
    "nz_scope_leave": (_i32, []),
    "nz_scope_close": (_i32, [C.c_int64]),
    "nz_flush_to_host": (_i32, [_vp]),
    "nz_pin": (_i32, [_vp, _sz]),
    "nz_unpin": (_i32, [_vp]),
    "nz_dev_fractal": (_i32, [_vp, _i32, _i32, _i32, _i32, _f32, _f32, _f32, _f32, _i32, _i32, _i32, _i32, _vp]),
    "nz_dev_separable": (_i32, [_vp, _vp, _i32, _i32, _i32, _pf32, _pf32, _f32, _i32, C.POINTER(_vp), _vp]),
    "nz_dev_kernel_filter": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, C.POINTER(_vp), _vp]),
    "nz_dev_min_erosion": (_i32, [_vp, _vp, _i32, _i32, _i32, C.POINTER(_vp), _vp]),
    "nz_dev_flowmap_scratch_bytes": (_sz, [_i32, _i32, _i32]),
    "nz_dev_flowmap": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _f32, _f32, C.POINTER(_vp), _vp]),
    "nz_dev_flow_walk_reruns": (_i32, [C.POINTER(C.c_uint64)]),
    "nz_dev_heightmap_mesh": (_i32, [_i32, _vp, _vp, _i32, _i32, _i32, _f32, _f32, _vp, _i32, _i32, _i32, _i32, _vp]),
    "nz_dev_thermal_erosion": (_i32, [_vp, _vp, _i32, _f32, _f32, _f32, _i32, C.POINTER(_vp), _vp]),
    "nz_dev_subtractive_flow_scratch_bytes": (_sz, [_i32, _i32]),
    "nz_dev_subtractive_flow_erosion": (_i32, [_vp, _vp, _i32, _i32, _i32, _f32, _f32, _f32, _vp]),
    "nz_dev_constant": (_i32, [_vp, _sz, _i32, _f32, _vp]),
    "nz_dev_reduce": (_i32, [_vp, _vp, _sz, _i32, _vp]),
    "nz_dev_curve": (_i32, [_vp, _sz, _vp, _i32, _vp]),
    "nz_dev_crop": (_i32, [_vp, _i32, _vp, _i32, _i32, _vp]),
    "nz_dev_map_range_scratch_bytes": (_sz, []),
    "nz_dev_map_range": (_i32, [_vp, _sz, _f32, _f32, _vp, _vp, _vp]),
    "nz_dev_normalize": (_i32, [_vp, _sz, _f32, _f32, _vp]),
    "nz_dev_fma_peak": (_i32, [_vp, _i32, _i32, C.POINTER(C.c_double), _vp]),
}

_lib = None


def build(verbose=False):
    """Compile libnoize_b200.so for sm_100a with csrc/Makefile (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("building libnoize_b200.so failed")
    return LIB_PATH


def load():
    """Load the CUDA library.  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(noize_b200 has no CPU or PyTorch fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc < 0:
        raise NzError(rc, load().nz_last_error().decode("utf-8", "replace"))
    return rc


def as_slice(a):
    """numpy 1-D (possibly strided) or C-contiguous 2-D float32 array -> nz_slice_f32 viewing its memory."""
    if a is None:
        return Slice(None, 0, 0)
    if not isinstance(a, np.ndarray) or a.dtype != np.float32:
        raise TypeError("expected a numpy float32 array")
    if a.ndim == 2:
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("2-D grids must be C-contiguous")
        return Slice(a.ctypes.data, 4, a.size)
    if a.ndim != 1:
        raise ValueError("expected a 1-D slice or a 2-D grid")
    return Slice(a.ctypes.data, a.strides[0] if a.size > 1 else 4, a.size)


def fptr(a):
    a = np.ascontiguousarray(a, np.float32)
    return a, a.ctypes.data_as(_pf32)


def stream_ptr(stream=None):
    """cudaStream_t of a torch stream (or the current torch stream)."""
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)
