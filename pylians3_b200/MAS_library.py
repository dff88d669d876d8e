"""Drop-in for the hot-path part of Pylians3's `MAS_library`: `MA` and the `*c3D/*c2D` wrappers.

Mirrors library/MAS_library/MAS_library.pyx:57-112 (MA) and :1305-1389 (NGPc3D ... PCSWc2D):
same names, positional order, defaults, in-place `+=` semantics on `number`, same printed
messages.  The arithmetic runs in libpyl_b200.so (hand-written sm_100a kernels) through the
C ABI of include/pyl_b200.h; there is no CPU path.

Accepted buffers
  * NumPy float32 arrays: copied host->device, deposited, and `number` copied back into the
    caller's array (the reference mutates `number` in place; so do we).
  * torch CUDA float32 tensors: used in place, zero-copy (`number` must be contiguous to be
    updated in place; a non-contiguous tensor is staged through a contiguous copy).
  The two can be mixed (e.g. host positions, device-resident grid).

Opt-in extras (keyword-only, defaults reproduce the reference): mode='auto'|'atomic'|'tiled'|
'deterministic' selects the deposit algorithm.
"""
import time

import numpy as np
import torch

from . import _device as D
from . import _lib as L
from .errors import reference_exit

__all__ = ["MA", "FLOAT_type",
           "NGPc3D", "NGPWc3D", "NGPc2D", "NGPWc2D", "CICc3D", "CICWc3D", "CICc2D", "CICWc2D",
           "TSCc3D", "TSCWc3D", "TSCc2D", "TSCWc2D", "PCSc3D", "PCSWc3D", "PCSc2D", "PCSWc2D"]

_RENORM = {"NGP": 1.0, "CIC": 2.0, "TSC": 3.0, "PCS": 4.0}


def FLOAT_type():
    """MAS_library.pyx:11-13 (FLOAT is float32 in the reference build)."""
    return np.float32


def _deposit_device(mas_id, pos_d, number_d, W_d, dims, axes, BoxSize, mode):
    lib = L.load()
    particles = pos_d.shape[0]
    dev = number_d.device
    with torch.cuda.device(dev):
        need = lib.pyl_deposit_workspace_bytes(mas_id, particles, dims, axes, mode)
        ws = D.workspace(need, dev, "deposit")
        st = lib.pyl_deposit(mas_id, D.ptr(pos_d), D.ptr(number_d), D.ptr(W_d), particles, dims, axes,
                             float(np.float32(BoxSize)), mode, D.ptr(ws), need, D.stream_ptr(dev))
    L.check(st, "pyl_deposit")


def MA(pos, number, BoxSize, MAS="CIC", W=None, verbose=False, renormalize_2D=True, *, mode="auto"):
    """Mass assignment: number += deposit(pos[, W]) with NGP/CIC/TSC/PCS, 2D or 3D.

    Reference: MAS_library.pyx:57-112.  Returns None; `number` is modified in place."""
    coord, coord_aux = pos.shape[1], number.ndim
    if coord != coord_aux:
        reference_exit("pos have %d dimensions and the density %d!!!" % (coord, coord_aux))
    if verbose:
        if W is None:
            print("\nUsing %s mass assignment scheme" % MAS)
        else:
            print("\nUsing %s mass assignment scheme with weights" % MAS)
    start = time.time()
    if MAS not in L.MAS_IDS or coord not in (2, 3):
        reference_exit("option not valid!!!")
    if mode not in L.MODE_IDS:
        raise ValueError("mode must be one of %s" % sorted(L.MODE_IDS))
    D.require_cuda()

    dims = number.shape[0]
    if any(s != dims for s in number.shape):
        raise ValueError("number must be a (dims,)*%d grid, got %s" % (coord, tuple(number.shape)))
    dev = D.pick_device(number, pos, W)
    pos_d, _ = D.to_device_f32(pos, dev, "pos")
    W_d = None
    if W is not None:
        W_d, _ = D.to_device_f32(W, dev, "W")
        if W_d.ndim != 1 or W_d.shape[0] != pos_d.shape[0]:
            raise ValueError("W must have one weight per particle")

    inplace = D.is_cuda_tensor(number) and number.is_contiguous()
    if inplace:
        if number.dtype != torch.float32:
            raise ValueError("number must be float32, got %s" % number.dtype)
        number_d = number
    else:
        number_d, _ = D.to_device_f32(number, dev, "number")   # staged copy (host or strided)

    _deposit_device(L.MAS_IDS[MAS], pos_d, number_d, W_d, dims, coord, BoxSize, L.MODE_IDS[mode])
    if coord == 2 and renormalize_2D and MAS != "NGP":
        # number2 /= 2.0|3.0|4.0 -- the WHOLE accumulated plane (MAS_library.pyx:90-107)
        with torch.cuda.device(dev):
            L.check(L.load().pyl_divide_inplace(D.ptr(number_d), number_d.numel(), _RENORM[MAS],
                                                D.stream_ptr(dev)), "pyl_divide_inplace")
    if not inplace:
        if isinstance(number, torch.Tensor):
            number.copy_(number_d)
        else:
            number[...] = number_d.cpu().numpy()
    if verbose:
        if D.is_cuda_tensor(number):
            torch.cuda.synchronize(dev)
        print("Time taken = %.3f seconds\n" % (time.time() - start))


# ---- MAS_c (OpenMP C core) wrappers, MAS_library.pyx:1305-1389 ---------------------------
# Same call shapes: pos, number[, W], BoxSize, threads.  `threads` is accepted and ignored.
def _c2d(pos, number, BoxSize, mas, W):
    """2D through the C core = the proper plane deposit, one add per cell (MAS_c.c: n_max=1),
    accumulated into `number` without touching what is already there."""
    if D.is_cuda_tensor(number):
        tmp = torch.zeros(number.shape, dtype=torch.float32, device=number.device)
        MA(pos, tmp, BoxSize, mas, W, renormalize_2D=True)
        number.add_(tmp)          # plumbing-level accumulate of two device planes
    else:
        tmp = np.zeros(number.shape, dtype=np.float32)
        MA(pos, tmp, BoxSize, mas, W, renormalize_2D=True)
        number += tmp


def _c_wrapper(mas, weighted, ndim):
    def fn(pos, number, *args):
        if weighted:
            W, BoxSize, threads = args
        else:
            (BoxSize, threads), W = args, None
        if number.ndim != ndim:
            raise ValueError("Buffer has wrong number of dimensions (expected %d, got %d)" % (ndim, number.ndim))
        if ndim == 2 and mas != "NGP":
            _c2d(pos, number, BoxSize, mas, W)
        else:
            MA(pos, number, BoxSize, mas, W)
    fn.__name__ = "%s%sc%dD" % (mas, "W" if weighted else "", ndim)
    fn.__doc__ = "MAS_c.%s on %dD grids (MAS_library.pyx:1305-1389); `threads` is ignored." % (mas, ndim)
    return fn


for _mas in ("NGP", "CIC", "TSC", "PCS"):
    for _w in (False, True):
        for _nd in (3, 2):
            _f = _c_wrapper(_mas, _w, _nd)
            globals()[_f.__name__] = _f
del _mas, _w, _nd, _f
