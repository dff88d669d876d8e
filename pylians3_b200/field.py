"""Device-side glue between MA and Pk: delta = n/<n> - 1 without a host round trip.

The reference's callers do this step in NumPy between the two library calls
(docs/source/construction.rst:50; Pk_library/Pk_snapshot.py:88).  For device-resident grids
the same arithmetic (float64 mean, float64 division, float32 subtraction) runs in two small
kernels of libpyl_b200.so."""
import torch

from . import _device as D
from . import _lib as L


def grid_sum(grid):
    """float64 sum of a float32 CUDA tensor, returned as a 1-element CUDA float64 tensor."""
    if not D.is_cuda_tensor(grid) or grid.dtype != torch.float32 or not grid.is_contiguous():
        raise ValueError("grid must be a contiguous float32 CUDA tensor")
    out = torch.empty(1, dtype=torch.float64, device=grid.device)
    with torch.cuda.device(grid.device):
        L.check(L.load().pyl_sum_f64(D.ptr(grid), grid.numel(), D.ptr(out), D.stream_ptr(grid.device)),
                "pyl_sum_f64")
    return out


def overdensity_(grid, total=None, cells=None):
    """In place: grid <- grid/mean(grid) - 1.  `total` (CUDA float64[1]) and `cells` override the
    local sum / cell count (multi-GPU: the all-reduced sum and dims**3)."""
    if total is None:
        total = grid_sum(grid)
    if cells is None:
        cells = grid.numel()
    with torch.cuda.device(grid.device):
        L.check(L.load().pyl_overdensity_inplace(D.ptr(grid), grid.numel(), D.ptr(total), float(cells),
                                                 D.stream_ptr(grid.device)), "pyl_overdensity_inplace")
    return grid


def prebias_(grid, particles, W=None, cells=None, scale=None):
    """Start value for a deposit whose spectrum is taken with Pk(..., density=True, offset=c): fills `grid` with
    -c, c = the mean density the deposit will produce (sum of the weights / cells), and returns c -- the exact
    float32 fill value -- as a CUDA float64[1] tensor.

    Why: a float32 FFT carries rounding noise proportional to its largest amplitudes; a density grid has its whole
    mass in the DC mode.  With the constant taken out beforehand the transform sees n - c, the DC mode is what is
    left of c's rounding, and <n> = c + DC/cells exactly.  The sum of the weights is exact (one pass over W on the
    device, 4 bytes per particle; nothing for unweighted particles) because the residual DC must stay below the
    fluctuation modes: |<n> - c|/<n> << sigma/sqrt(cells), 1e-5 at 1024^3.  No host synchronisation: c stays on
    the device.  `scale` (CUDA float64[1]) replaces the sum by an agreed numerator (multi-GPU: the all-reduced
    one, see SlabContext.prebias_)."""
    if not D.is_cuda_tensor(grid) or grid.dtype != torch.float32 or not grid.is_contiguous():
        raise ValueError("grid must be a contiguous float32 CUDA tensor")
    cells = grid.numel() if cells is None else cells
    if scale is None:
        scale = weight_total(particles, W, grid.device)
    c = torch.empty(1, dtype=torch.float64, device=grid.device)
    with torch.cuda.device(grid.device):
        L.check(L.load().pyl_fill_negative(D.ptr(grid), grid.numel(), D.ptr(scale), float(cells), D.ptr(c),
                                           D.stream_ptr(grid.device)), "pyl_fill_negative")
    return c


HOST_SUM_LIMIT = 1 << 24


def weight_total(particles, W, device):
    """Sum of the weights (the particle count if W is None) as a float64[1] tensor on `device`.  CUDA weights: exact
    float64 sum (pyl_sum_f64).  Host weights: exact up to HOST_SUM_LIMIT elements; beyond that particles * mean of a
    strided sample of 2^20 weights -- an ESTIMATE (a pass over GBs of host memory would cost more than the
    overdensity_ pass prebias_ replaces), good to ~1e-3: use overdensity_ instead when weights of that size
    live on the host."""
    if W is None:
        return torch.full((1,), float(particles), dtype=torch.float64, device=device)
    if D.is_cuda_tensor(W):
        if W.dtype != torch.float32 or not W.is_contiguous():
            raise ValueError("W must be a contiguous float32 tensor")
        return grid_sum(W)
    if not torch.is_tensor(W):
        W = torch.from_numpy(W)
    W = W.reshape(-1)
    if W.numel() <= HOST_SUM_LIMIT:
        m = W.to(torch.float64).sum().reshape(1)
    else:
        stride = max(1, W.numel() >> 20)
        m = W[::stride].to(torch.float64).mean().reshape(1) * float(particles)
    return m.to(device, non_blocking=True)
