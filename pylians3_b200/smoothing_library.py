"""Drop-in for Pylians3's `smoothing_library` (library/smoothing_library/smoothing_library.pyx):
  FT_filter :20-120, FT_filter_2D :123-209, field_smoothing :215-235, field_smoothing_2D :243-262
with the same names, positional order and defaults.  The filter is placed on the grid, normalised, transformed,
multiplied into the field's transform and transformed back on the GPU (pyl_filter_fill, pyl_sum_f64,
pyl_divide_by_f64, pyl_cmul_inplace of include/pyl_b200.h; cuFFT for the transforms).  There is no CPU path.

Arrays may be NumPy arrays (results come back as NumPy arrays, like the reference) or torch CUDA tensors
(results stay on the device); `as_tensor=True` makes FT_filter return a CUDA tensor that can be handed to
field_smoothing without a host round trip."""
import ctypes

import numpy as np
import torch

from . import _device as D
from . import _lib as L
from . import Pk_library as _P
from . import _pk_more as _M
from ._pk_more import field_smoothing  # noqa: F401

__all__ = ["FT_filter", "FT_filter_2D", "field_smoothing", "field_smoothing_2D"]

_KINDS = {"Top-Hat": 0, "Gaussian": 1, "Top-Hat-k": 2}


def _ifft2d_c2r_device(ak_d):
    """(grid, grid//2+1) complex64 CUDA tensor -> (grid, grid) float32, NORMALISED like IFFT2Dr_f
    (Pk_library.pyx:213-226; pyfftw scales inverse transforms by 1/grid^2).  `ak_d` is consumed."""
    lib = L.load()
    grid = ak_d.shape[0]
    dev = ak_d.device
    with torch.cuda.device(dev):
        out = torch.empty((grid, grid), dtype=torch.float32, device=dev)
        need = lib.pyl_fft2d_c2r_workspace_bytes(grid)
        if need == ctypes.c_size_t(-1).value:
            L.check(-3, "pyl_fft2d_c2r_workspace_bytes")
        ws = D.workspace(need, dev, "fft")
        L.check(lib.pyl_fft2d_c2r(D.ptr(ak_d), D.ptr(out), grid, D.ptr(ws), need, D.stream_ptr(dev)), "pyl_fft2d_c2r")
        L.check(lib.pyl_scale_inplace(D.ptr(out), out.numel(), float(np.float32(1.0) / np.float32(grid ** 2)),
                                      D.stream_ptr(dev)), "pyl_scale_inplace")
    return out


def _ft_filter(BoxSize, R, dims, Filter, kmin, kmax, nd, as_tensor):
    if Filter not in ["Top-Hat", "Gaussian", "Top-Hat-k"]:
        raise Exception("Filter %s not implemented!" % Filter)
    D.require_cuda()
    lib = L.load()
    f32 = np.float32
    # the reference's `float` locals (smoothing_library.pyx:20-36)
    BoxSize, R, kmin, kmax = f32(BoxSize), f32(R), f32(kmin), f32(kmax)
    dims = int(dims)
    R_grid = f32(f32(R * f32(dims)) / BoxSize)
    R2 = f32(R_grid * R_grid)
    kF = f32(2.0 * np.pi / float(BoxSize))
    dev = torch.device("cuda", torch.cuda.current_device())
    kind = _KINDS[Filter]
    with torch.cuda.device(dev):
        s = D.stream_ptr(dev)
        if kind == 2:
            field_k = torch.empty((dims,) * (nd - 1) + (dims // 2 + 1,), dtype=torch.complex64, device=dev)
            L.check(lib.pyl_filter_fill(kind, D.ptr(field_k), dims, nd, float(R2), float(kF), float(kmin), float(kmax), s),
                    "pyl_filter_fill")
            field = _M.ifft3d_c2r_device(field_k) if nd == 3 else _ifft2d_c2r_device(field_k)
        else:
            field = torch.empty((dims,) * nd, dtype=torch.float32, device=dev)
            L.check(lib.pyl_filter_fill(kind, D.ptr(field), dims, nd, float(R2), float(kF), float(kmin), float(kmax), s),
                    "pyl_filter_fill")
        norm = torch.empty(1, dtype=torch.float64, device=dev)
        L.check(lib.pyl_sum_f64(D.ptr(field), field.numel(), D.ptr(norm), s), "pyl_sum_f64")
        L.check(lib.pyl_divide_by_f64(D.ptr(field), field.numel(), D.ptr(norm), s), "pyl_divide_by_f64")
    out = _P.fft3d_r2c_device(field) if nd == 3 else _M.fft2d_r2c_device(field)
    return out if as_tensor else out.cpu().numpy()


def FT_filter(BoxSize, R, dims, Filter, threads=1, kmin=0, kmax=0, *, as_tensor=False):
    """Fourier transform of a Top-Hat / Gaussian / Top-Hat-k filter of radius R on a dims^3 grid
    (smoothing_library.pyx:20-120).  Returns (dims,dims,dims//2+1) complex64."""
    return _ft_filter(BoxSize, R, dims, Filter, kmin, kmax, 3, as_tensor)


def FT_filter_2D(BoxSize, R, grid, Filter, threads=1, kmin=0, kmax=0, *, as_tensor=False):
    """2D version (smoothing_library.pyx:123-209).  Returns (grid, grid//2+1) complex64."""
    return _ft_filter(BoxSize, R, grid, Filter, kmin, kmax, 2, as_tensor)


def field_smoothing_2D(field, filter_k, threads=1):
    """IFFT2(FFT2(field) * filter_k) (smoothing_library.pyx:243-262)."""
    D.require_cuda()
    if field.shape[0] != filter_k.shape[0]:
        raise Exception("field and filter have different grids!!!")
    dev = D.pick_device(field, filter_k)
    img = _M._as_image(field, dev, "field")
    grid = img.shape[0]
    if isinstance(filter_k, torch.Tensor):
        if filter_k.dtype != torch.complex64:
            raise ValueError("filter_k must be complex64")
        fk = filter_k.to(dev).contiguous()
    else:
        fk = np.asarray(filter_k)
        if fk.dtype != np.complex64:
            raise ValueError("Buffer dtype mismatch, expected 'complex64_t' but got '%s'" % fk.dtype)
        fk = torch.from_numpy(np.ascontiguousarray(fk)).to(dev)
    if tuple(fk.shape) != (grid, grid // 2 + 1):
        raise ValueError("filter_k must have shape (grid, grid//2+1)")
    field_k = _M.fft2d_r2c_device(img)
    with torch.cuda.device(dev):
        L.check(L.load().pyl_cmul_inplace(D.ptr(field_k), D.ptr(fk), field_k.numel(), D.stream_ptr(dev)), "pyl_cmul_inplace")
    out = _ifft2d_c2r_device(field_k)
    return out if D.is_cuda_tensor(field) else out.cpu().numpy()
