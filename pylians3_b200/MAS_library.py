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

__all__ = ["MA", "CIC_interp", "FLOAT_type", "NGP", "CIC", "TSC", "PCS", "NGPW", "CICW", "TSCW", "PCSW",
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


# Host-resident particles are streamed: chunk k+1 crosses PCIe (copy stream, double-buffered staging)
# while chunk k is deposited.  Each chunk is a complete deposit into the same grid (MA accumulates,
# MAS_library.pyx:57-112 / MAS_gadget.py:63-75), so nothing has to wait for the whole array.
STREAM_CHUNK = 16 * 1024 * 1024          # particles per chunk (192 MB of positions: ~3.5 ms of PCIe gen5)
_copy_streams = {}


def stream_chunk(dims, coord=3):
    """Particles per streamed chunk: every chunk is a complete tiled deposit, which adds each 8x16x32-cell tile
    (+ halo) to the grid once, so a chunk should bring a few hundred particles per tile (1024^3: 8 chunks of 134 M)."""
    if coord != 3:
        return STREAM_CHUNK
    tiles = ((dims + 7) // 8) * ((dims + 15) // 16) * ((dims + 31) // 32)
    return max(STREAM_CHUNK, 512 * tiles)


def _host_f32(x, name):
    """numpy array / CPU tensor -> contiguous float32 CPU tensor sharing memory when possible."""
    if isinstance(x, torch.Tensor):
        if x.dtype != torch.float32:
            raise ValueError("%s must be float32, got %s" % (name, x.dtype))
        return x.contiguous()
    a = np.asarray(x)
    if a.dtype != np.float32:
        raise ValueError("%s must be float32 (Buffer dtype mismatch, expected 'float32_t' but got '%s')"
                         % (name, a.dtype))
    return torch.from_numpy(np.ascontiguousarray(a))


def stream_host_chunks(pos_h, W_h, dev, chunk, consume):
    """Feed host-resident particles to `consume(pos_chunk_d, W_chunk_d)` chunk by chunk: the copy of chunk
    k+1 (copy stream, double-buffered staging in HBM) overlaps whatever `consume` enqueued for chunk k."""
    n, axes = pos_h.shape
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    copy_stream = _copy_streams.get(key)
    if copy_stream is None:
        copy_stream = _copy_streams[key] = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    stage = D.workspace(2 * chunk * (axes + 1) * 4, dev, "ma_stage").view(torch.float32)
    pbuf = [stage[j * chunk * axes:(j + 1) * chunk * axes].view(chunk, axes) for j in range(2)]
    wbuf = [stage[2 * chunk * axes + j * chunk:2 * chunk * axes + (j + 1) * chunk] for j in range(2)]
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    copy_stream.wait_stream(main)                 # staging buffers may still be read by earlier work
    for k, a in enumerate(range(0, n, chunk)):
        b, j = min(a + chunk, n), k & 1
        with torch.cuda.stream(copy_stream):
            if k >= 2:
                copy_stream.wait_event(done[j])   # the deposit that last read this buffer has finished
            pbuf[j][:b - a].copy_(pos_h[a:b], non_blocking=True)
            if W_h is not None:
                wbuf[j][:b - a].copy_(W_h[a:b], non_blocking=True)
            copied[j].record(copy_stream)
        main.wait_event(copied[j])
        consume(pbuf[j][:b - a], None if W_h is None else wbuf[j][:b - a])
        done[j].record(main)


def _deposit_streamed(mas_id, pos_h, number_d, W_h, dims, axes, BoxSize, mode, chunk):
    stream_host_chunks(pos_h, W_h, number_d.device, chunk,
                       lambda p, w: _deposit_device(mas_id, p, number_d, w, dims, axes, BoxSize, mode))


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
    chunk = stream_chunk(dims, coord)
    streamed = (not D.is_cuda_tensor(pos)) and (W is None or not D.is_cuda_tensor(W)) and \
        pos.shape[0] >= 2 * chunk
    pos_d = W_d = None
    if streamed:
        pos_h = _host_f32(pos, "pos")
        W_h = None if W is None else _host_f32(W, "W")
        if W_h is not None and (W_h.ndim != 1 or W_h.shape[0] != pos_h.shape[0]):
            raise ValueError("W must have one weight per particle")
    else:
        pos_d, _ = D.to_device_f32(pos, dev, "pos")
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

    if streamed:
        with torch.cuda.device(dev):
            _deposit_streamed(L.MAS_IDS[MAS], pos_h, number_d, W_h, dims, coord, BoxSize, L.MODE_IDS[mode], chunk)
    else:
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


def CIC_interp(density, BoxSize, pos, den):
    """den[i] = CIC-interpolated value of the 3D grid `density` at pos[i]; `den` is overwritten in place.

    Reference: MAS_library.pyx:558-599 (used for marked power spectra).  NumPy float32 arrays or torch
    CUDA float32 tensors (zero-copy); returns None."""
    D.require_cuda()
    if density.ndim != 3 or pos.ndim != 2 or pos.shape[1] != 3 or den.ndim != 1:
        raise ValueError("Buffer has wrong number of dimensions (density 3D, pos (N,3), den (N,))")
    dims = density.shape[0]
    if any(s != dims for s in density.shape):
        raise ValueError("density must be a (dims,dims,dims) grid, got %s" % (tuple(density.shape),))
    if den.shape[0] != pos.shape[0]:
        raise ValueError("den must have one entry per position")
    dev = D.pick_device(density, pos, den)
    density_d, _ = D.to_device_f32(density, dev, "density")
    pos_d, _ = D.to_device_f32(pos, dev, "pos")
    inplace = D.is_cuda_tensor(den) and den.is_contiguous()
    if inplace and den.dtype != torch.float32:
        raise ValueError("den must be float32, got %s" % den.dtype)
    if not inplace and not isinstance(den, torch.Tensor) and np.asarray(den).dtype != np.float32:
        raise ValueError("den must be float32 (Buffer dtype mismatch, expected 'float32_t' but got '%s')"
                         % np.asarray(den).dtype)
    den_d = den if inplace else torch.empty(pos_d.shape[0], dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L.check(L.load().pyl_cic_interp(D.ptr(density_d), dims, float(np.float32(BoxSize)), D.ptr(pos_d),
                                        pos_d.shape[0], D.ptr(den_d), D.stream_ptr(dev)), "pyl_cic_interp")
    if not inplace:
        if isinstance(den, torch.Tensor):
            den.copy_(den_d)
        else:
            den[...] = den_d.cpu().numpy()


# ---- the per-scheme entry points, MAS_library.pyx:123-545 ------------------------------------------------
# MASL.CIC(pos, number, BoxSize), MASL.CICW(pos, number, BoxSize, W), ...: cpdef functions of the reference that
# MA dispatches to and that scripts also call directly.  Same arguments; the deposit accumulates into `number`.
# Called directly on a plane they do NOT renormalise (the reference divides only inside MA, :84-110).
def _scheme(mas, weighted):
    def plane(pos, number):
        # the reference's 2D form of these functions takes the (dims, dims, 1) view MA builds (:84)
        return number[:, :, 0] if number.ndim == 3 and pos.shape[1] == 2 and number.shape[2] == 1 else number

    if weighted:
        def fn(pos, number, BoxSize, W):
            MA(pos, plane(pos, number), BoxSize, mas, W, renormalize_2D=False)
    else:
        def fn(pos, number, BoxSize):
            MA(pos, plane(pos, number), BoxSize, mas, None, renormalize_2D=False)
    fn.__name__ = mas + ("W" if weighted else "")
    fn.__doc__ = "MAS_library.%s (MAS_library.pyx): %s deposit%s, accumulated into `number`." % (
        fn.__name__, mas, " with weights" if weighted else "")
    return fn


for _m in ("NGP", "CIC", "TSC", "PCS"):
    globals()[_m] = _scheme(_m, False)
    globals()[_m + "W"] = _scheme(_m, True)
del _m


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
