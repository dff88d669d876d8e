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
