"""Device layer of the C ABI (nz_dev_*): caller-owned DEVICE buffers on a caller stream.

torch is used only as the device allocator / stream provider: every function takes torch CUDA tensors,
passes their raw pointers to libnoize_b200.so and returns.  Grids are (rows, width) float32, contiguous.
Functions that ping-pong return the tensor (one of the two passed in) that holds the result.
"""
import ctypes as C

from . import lib as _l


def _grid(t, name="grid"):
    import torch
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.is_contiguous()):
        raise TypeError(f"{name} must be a contiguous 2-D float32 CUDA tensor")
    return t


def _pick(result_ptr, a, b):
    return a if result_ptr.value == a.data_ptr() else b


def fractal(dst, noise_type, hurst, starting_amplitude=1.0, stepdown=2.0, detune_rate=0.0, octaves=1, xpos=0, zpos=0,
            noise_size=1000, z_first=0, stream=None):
    """Rows [z_first, z_first+rows) of the tile at (xpos,zpos); see nz_dev_fractal."""
    _grid(dst, "dst")
    rows, width = dst.shape
    _l.check(_l.load().nz_dev_fractal(dst.data_ptr(), width, rows, z_first, int(noise_type), hurst, starting_amplitude,
                                      stepdown, detune_rate, octaves, xpos, zpos, noise_size, _l.stream_ptr(stream)))
    return dst


def separable(data, tmp, kx, kz, factor=1.0, iterations=1, stream=None):
    _grid(data, "data"); _grid(tmp, "tmp")
    assert data.shape == tmp.shape
    kx, pkx = _l.fptr(kx)
    kz, pkz = _l.fptr(kz)
    res = C.c_void_p()
    _l.check(_l.load().nz_dev_separable(data.data_ptr(), tmp.data_ptr(), data.shape[1], data.shape[0], kx.size, pkx, pkz,
                                        factor, iterations, C.byref(res), _l.stream_ptr(stream)))
    return _pick(res, data, tmp)


def kernel_filter(data, tmp, filter_type, iterations=1, stream=None):
    _grid(data, "data"); _grid(tmp, "tmp")
    assert data.shape == tmp.shape
    res = C.c_void_p()
    _l.check(_l.load().nz_dev_kernel_filter(data.data_ptr(), tmp.data_ptr(), data.shape[1], data.shape[0], int(filter_type),
                                            iterations, C.byref(res), _l.stream_ptr(stream)))
    return _pick(res, data, tmp)


def min_erosion(data, tmp, iterations=1, stream=None):
    _grid(data, "data"); _grid(tmp, "tmp")
    assert data.shape == tmp.shape
    res = C.c_void_p()
    _l.check(_l.load().nz_dev_min_erosion(data.data_ptr(), tmp.data_ptr(), data.shape[1], data.shape[0], iterations,
                                          C.byref(res), _l.stream_ptr(stream)))
    return _pick(res, data, tmp)


def flow_walk_reruns():
    """Launches of the register-resident flow kernel that were rerun on the wavefront kernel (guarded fast paths left)."""
    n = C.c_uint64(0)
    _l.check(_l.load().nz_dev_flow_walk_reruns(C.byref(n)))
    return int(n.value)


def flowmap_scratch_bytes(width, rows, iterations):
    return int(_l.load().nz_dev_flowmap_scratch_bytes(width, rows, iterations))


def flowmap(height, tmp, scratch=None, iterations=5, norm_min=-0.1, norm_max=0.1, stream=None):
    """`tmp`: ping-pong partner of `height`; `scratch`: a CUDA tensor of at least flowmap_scratch_bytes(...)
    bytes (may be None when that is 0).  Returns the tensor (height or tmp) that holds the result."""
    _grid(height, "height"); _grid(tmp, "tmp")
    assert height.shape == tmp.shape
    rows, width = height.shape
    need = flowmap_scratch_bytes(width, rows, iterations)
    have = 0 if scratch is None else scratch.numel() * scratch.element_size()
    if have < need:
        raise ValueError(f"flowmap scratch too small: {have} < {need}")
    res = C.c_void_p()
    _l.check(_l.load().nz_dev_flowmap(height.data_ptr(), tmp.data_ptr(), None if scratch is None else scratch.data_ptr(),
                                      width, rows, iterations, norm_min, norm_max, C.byref(res), _l.stream_ptr(stream)))
    return _pick(res, height, tmp)


def heightmap_mesh(mesh_type, vertices, indices, resolution, input_resolution, margin_pix, tile_height, tile_size,
                   heights, h_row_first=0, vz_begin=0, vz_end=None, stream=None):
    """`heights` holds rows [h_row_first, h_row_first+heights.shape[0]) of the input grid.  `vertices` /
    `indices` are CUDA tensors whose first element is vertex row vz_begin / triangle row max(vz_begin,1)."""
    _grid(heights, "heights")
    if vz_end is None:
        vz_end = resolution + 1
    _l.check(_l.load().nz_dev_heightmap_mesh(int(mesh_type), vertices.data_ptr(), indices.data_ptr(), resolution,
                                             input_resolution, margin_pix, tile_height, tile_size, heights.data_ptr(),
                                             h_row_first, heights.shape[0], vz_begin, vz_end, _l.stream_ptr(stream)))


# ---- SURVEY.md section 8f rows (all in place on a square or rectangular device grid) -----------------------
def thermal_erosion(data, talus=45.0, increment_ratio=0.5, mesh_height_width_ratio=0.75, iterations=1, stream=None, tmp=None):
    """Without `tmp`: in place (four launches per iteration).  With `tmp` (same shape): one fused launch per iteration,
    ping-pong; returns the tensor (data or tmp) that holds the result."""
    _grid(data, "data")
    assert data.shape[0] == data.shape[1], "thermal erosion works on a square tile"
    if tmp is not None:
        _grid(tmp, "tmp")
        assert tmp.shape == data.shape
    res = C.c_void_p()
    _l.check(_l.load().nz_dev_thermal_erosion(data.data_ptr(), None if tmp is None else tmp.data_ptr(), data.shape[0], talus,
                                              increment_ratio, mesh_height_width_ratio, iterations, C.byref(res),
                                              _l.stream_ptr(stream)))
    return tmp if tmp is not None and res.value == tmp.data_ptr() else data


def subtractive_flow_erosion(height, erosive_iterations=5, erosive_factor=0.1, norm_min=-0.1, norm_max=0.1, stream=None):
    """In place on a device grid; water + 4 flow fields live in a scratch tensor for the call."""
    import torch
    _grid(height, "height")
    rows, width = height.shape
    need = _l.load().nz_dev_subtractive_flow_scratch_bytes(width, rows)
    scratch = torch.empty(need // 4, dtype=torch.float32, device=height.device) if erosive_iterations > 0 else None
    _l.check(_l.load().nz_dev_subtractive_flow_erosion(height.data_ptr(), scratch.data_ptr() if scratch is not None else None,
                                                       width, rows, erosive_iterations, erosive_factor, norm_min, norm_max,
                                                       _l.stream_ptr(stream)))
    return height


def constant(data, operation, value, stream=None):
    _grid(data, "data")
    _l.check(_l.load().nz_dev_constant(data.data_ptr(), data.numel(), int(operation), value, _l.stream_ptr(stream)))
    return data


def reduce(left, right, operation, stream=None):
    _grid(left, "left"); _grid(right, "right")
    assert left.shape == right.shape
    _l.check(_l.load().nz_dev_reduce(left.data_ptr(), right.data_ptr(), left.numel(), int(operation), _l.stream_ptr(stream)))
    return left


def curve(data, curve_samples, stream=None):
    _grid(data, "data")
    assert curve_samples.is_cuda and curve_samples.is_contiguous() and curve_samples.dim() == 1
    _l.check(_l.load().nz_dev_curve(data.data_ptr(), data.numel(), curve_samples.data_ptr(), curve_samples.numel(),
                                    _l.stream_ptr(stream)))
    return data


def crop(input, output, offset=0, stream=None):
    _grid(input, "input"); _grid(output, "output")
    assert input.shape[0] == input.shape[1] and output.shape[0] == output.shape[1]
    _l.check(_l.load().nz_dev_crop(input.data_ptr(), input.shape[0], output.data_ptr(), output.shape[0], offset,
                                   _l.stream_ptr(stream)))
    return output


def map_range(data, lim_min=float("inf"), lim_max=float("-inf"), stream=None):
    """Returns a 3-element CUDA tensor [min, max, range] (no synchronisation)."""
    import torch
    _grid(data, "data")
    res = torch.empty(3, dtype=torch.float32, device=data.device)
    scratch = torch.empty(int(_l.load().nz_dev_map_range_scratch_bytes()), dtype=torch.uint8, device=data.device)
    _l.check(_l.load().nz_dev_map_range(data.data_ptr(), data.numel(), lim_min, lim_max, res.data_ptr(), scratch.data_ptr(),
                                        _l.stream_ptr(stream)))
    return res


def normalize(data, vmin, vrange, stream=None):
    _grid(data, "data")
    _l.check(_l.load().nz_dev_normalize(data.data_ptr(), data.numel(), vmin, vrange, _l.stream_ptr(stream)))
    return data


def fma_peak(sink, grid, iters, stream=None):
    flops = C.c_double()
    _l.check(_l.load().nz_dev_fma_peak(sink.data_ptr(), grid, iters, C.byref(flops), _l.stream_ptr(stream)))
    return flops.value
