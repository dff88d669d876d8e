"""Drop-in for the hot-path part of Pylians3's `Pk_library`: `Pk`, `XPk` and their helpers.

Mirrors library/Pk_library/Pk_library.pyx:
  frequencies :56-61, MAS_function :72-78, MAS_correction :83-84, check_number_modes :87-99,
  FFT3Dr_f :117-130, class Pk :263-420, class XPk :529-793
with the same names, positional order, defaults, printed messages and result attributes
(k3D, Pk, Nmodes3D, Pkphase, k1D, Pk1D, Nmodes1D, kpar, kper, Pk2D, Nmodes2D; XPk adds XPk,
PkX1D, PkX2D).  The FFT (cuFFT), the MAS-window deconvolution and the |delta_k|^2 binning run
on the GPU through the C ABI of include/pyl_b200.h; the O(bins) finalisation (units, averages;
:384-418 and :735-791) stays on the host, vectorised.  There is no CPU path.

`delta` may be a NumPy float32 array (copied host->device) or a torch CUDA float32 tensor
(zero-copy).  `threads` is accepted and ignored.
"""
import ctypes
import time

import numpy as np
import torch

from . import _device as D
from . import _lib as L
from .errors import reference_exit

__all__ = ["Pk", "XPk", "frequencies", "MAS_function", "MAS_correction", "check_number_modes", "FFT3Dr_f"]


# ---- helpers with the reference's names ---------------------------------------------------
def frequencies(BoxSize, dims):
    """Pk_library.pyx:56-61."""
    kF = 2.0 * np.pi / BoxSize
    middle = dims // 2
    kN = middle * kF
    kmax_par = middle
    kmax_per = int(np.sqrt(middle ** 2 + middle ** 2))
    kmax = int(np.sqrt(middle ** 2 + middle ** 2 + middle ** 2))
    return kF, kN, kmax_par, kmax_per, kmax


def MAS_function(MAS):
    """Pk_library.pyx:72-78: exponent of the window; anything unknown (None, 'None') -> 0."""
    MAS_index = 0
    if MAS == "NGP":
        MAS_index = 1
    if MAS == "CIC":
        MAS_index = 2
    if MAS == "TSC":
        MAS_index = 3
    if MAS == "PCS":
        MAS_index = 4
    return MAS_index


def MAS_correction(x, MAS_index):
    """Pk_library.pyx:83-84."""
    return 1.0 if x == 0.0 else float(np.power(x / np.sin(x), MAS_index))


def check_number_modes(Nmodes, dims):
    """Pk_library.pyx:87-99: every independent mode must have been counted exactly once."""
    own_modes = 1 if dims % 2 == 1 else 8
    repeated_modes = (dims ** 3 - own_modes) // 2
    indep_modes = repeated_modes + own_modes
    if int(np.sum(Nmodes)) != indep_modes:
        reference_exit("WARNING: Not all modes counted",
                       "Counted  %d independent modes" % (int(np.sum(Nmodes))),
                       "Expected %d independent modes" % indep_modes)


# ---- device stages ------------------------------------------------------------------------
def _as_delta(delta, dev, name="delta"):
    if isinstance(delta, torch.Tensor):
        if delta.ndim != 3:
            raise ValueError("Buffer has wrong number of dimensions (expected 3, got %d)" % delta.ndim)
    else:
        delta = np.asarray(delta) if not isinstance(delta, np.ndarray) else delta
        if delta.ndim != 3:
            raise ValueError("Buffer has wrong number of dimensions (expected 3, got %d)" % delta.ndim)
    t, _ = D.to_device_f32(delta, dev, name)
    if not (t.shape[0] == t.shape[1] == t.shape[2]):
        raise ValueError("%s must be a (dims,dims,dims) grid, got %s" % (name, tuple(t.shape)))
    return t


def fft3d_r2c_device(delta_d):
    """(dims,dims,dims) float32 CUDA tensor -> (dims,dims,dims//2+1) complex64 CUDA tensor (cuFFT)."""
    lib = L.load()
    dims = delta_d.shape[0]
    dev = delta_d.device
    with torch.cuda.device(dev):
        out = torch.empty((dims, dims, dims // 2 + 1), dtype=torch.complex64, device=dev)
        need = lib.pyl_fft_r2c_workspace_bytes(dims)
        if need == ctypes.c_size_t(-1).value:
            L.check(-3, "pyl_fft_r2c_workspace_bytes")
        ws = D.workspace(need, dev, "fft")
        L.check(lib.pyl_fft_r2c(D.ptr(delta_d), D.ptr(out), dims, D.ptr(ws), need, D.stream_ptr(dev)),
                "pyl_fft_r2c")
    return out


def FFT3Dr_f(a, threads=1):
    """Pk_library.pyx:117-130: unnormalised forward r2c of a float32 cube, returned as complex64
    ndarray (dims,dims,dims//2+1).  Given a CUDA tensor it returns a CUDA tensor."""
    D.require_cuda()
    dev = D.pick_device(a)
    out = fft3d_r2c_device(_as_delta(a, dev, "a"))
    return out if D.is_cuda_tensor(a) else out.cpu().numpy()


def bin_device(dk_list, mas_index, dims, axis, want_phase=False, ky_lo=0, nky=None, mirrored=False, flags=0):
    """Run pyl_pk_bin on `len(dk_list)` (<= L.MAX_FIELDS) half-spectra; returns (int64 CUDA tensor
    holding the raw accumulators, layout).  mirrored=True: the fields hold the |ky| rows [ky_lo, ky_lo+nky)
    and their mirrors (pyl_pk_bin_mirrored, multi-GPU slabs).  flags: extra PYL_PK_* bits (L.PK_CROSS_IMAG)."""
    lib = L.load()
    F = len(dk_list)
    nky = dims if nky is None else nky
    dev = dk_list[0].device
    lay = L.pk_layout(dims, F)
    with torch.cuda.device(dev):
        # room behind the accumulators for kpar/kper (finalize_device)
        out = torch.empty(lay.total_words + 2 * lay.n2d, dtype=torch.int64, device=dev)
        need = lib.pyl_pk_bin_workspace_bytes(dims, F)
        ws = D.workspace(need, dev, "pkbin")
        ptrs = (ctypes.c_void_p * F)(*[D.ptr(t) for t in dk_list])
        mi = (ctypes.c_int * F)(*[int(i) for i in mas_index])
        fn = lib.pyl_pk_bin_mirrored if mirrored else lib.pyl_pk_bin
        st = fn(ptrs, F, mi, dims, ky_lo, nky, axis, (L.PK_PHASE if want_phase else 0) | int(flags), D.ptr(out), D.ptr(ws), need,
                D.stream_ptr(dev))
    L.check(st, "pyl_pk_bin")
    return out, lay


def unpack_raw(words, lay):
    """Split the flat accumulator buffer (host int64 ndarray) into named float64 arrays."""
    F, X = lay.fields, lay.xfields
    n3, n1, n2 = lay.kmax + 1, lay.kmax_par + 1, lay.n2d
    f64 = words.view(np.float64)

    def cnt(off, n):
        return words[off:off + n].astype(np.float64)

    def dbl(off, *shape):
        n = int(np.prod(shape))
        return f64[off:off + n].reshape(shape).copy()

    Nm1D = cnt(lay.Nm1D, n1)
    return dict(k3D=dbl(lay.k3D, n3), Nm3D=cnt(lay.Nm3D, n3), Pk3D=dbl(lay.Pk3D, n3, 3, F),
                PkX3D=dbl(lay.PkX3D, n3, 3, X), phase=dbl(lay.phase, n3),
                Nm1D=Nm1D, k1D=Nm1D * np.arange(n1, dtype=np.float64),     # sum of k_par over the bin
                Pk1D=dbl(lay.Pk1D, n1, F), PkX1D=dbl(lay.PkX1D, n1, X),
                Nm2D=cnt(lay.Nm2D, n2), Pk2D=dbl(lay.Pk2D, n2, F), PkX2D=dbl(lay.PkX2D, n2, X))


def bin_fields(dk_list, mas_index, dims, axis, want_phase=False, ky_lo=0, nky=None, reduce_fn=None,
               mirrored=False, flags=0):
    """Raw accumulators for ANY number of fields.

    Up to L.MAX_FIELDS fields go through one launch.  More fields are covered by launches over
    pairs of field blocks (each launch yields the autos of its fields and all their crosses), and
    the results are scattered into the (.., F) / (.., X) arrays in the reference's pair order.
    `reduce_fn(int64 cuda tensor, layout)` lets the multi-GPU path all-reduce before the D2H."""
    F = len(dk_list)

    def run(idx):
        out, lay = bin_device([dk_list[i] for i in idx], [mas_index[i] for i in idx], dims, axis,
                              want_phase and len(idx) == 1, ky_lo, nky, mirrored, flags)
        if reduce_fn is not None:
            out = reduce_fn(out, lay)
        return unpack_raw(D.to_host_numpy(out), lay)

    if F <= L.MAX_FIELDS:
        return run(list(range(F)))

    pair_id = {}
    for i in range(F):
        for j in range(i + 1, F):
            pair_id[(i, j)] = len(pair_id)
    X = len(pair_id)
    half = L.MAX_FIELDS // 2
    blocks = [list(range(b, min(b + half, F))) for b in range(0, F, half)]
    res, seen_auto, seen_pair = None, set(), set()
    for bi in range(len(blocks)):
        for bj in range(bi + 1, len(blocks)) if len(blocks) > 1 else []:
            idx = blocks[bi] + blocks[bj]
            r = run(idx)
            if res is None:
                res = {k: v for k, v in r.items() if k in ("k3D", "Nm3D", "Nm1D", "k1D", "Nm2D", "phase")}
                for k in ("Pk3D", "Pk1D", "Pk2D"):
                    res[k] = np.zeros(r[k].shape[:-1] + (F,))
                for k in ("PkX3D", "PkX1D", "PkX2D"):
                    res[k] = np.zeros(r[k].shape[:-1] + (X,))
            for a, ga in enumerate(idx):
                if ga not in seen_auto:
                    seen_auto.add(ga)
                    for k in ("Pk3D", "Pk1D", "Pk2D"):
                        res[k][..., ga] = r[k][..., a]
            lx = 0
            for a in range(len(idx)):
                for b in range(a + 1, len(idx)):
                    g = (idx[a], idx[b])
                    if g not in seen_pair:
                        seen_pair.add(g)
                        for k in ("PkX3D", "PkX1D", "PkX2D"):
                            res[k][..., pair_id[g]] = r[k][..., lx]
                    lx += 1
    return res


class _Spectra(dict):
    """Result dict of finalize_device: "kpar"/"kper" appear on first access (host index arithmetic from "kgrid")."""

    def __missing__(self, key):
        if key in ("kpar", "kper"):
            self["kpar"], self["kper"] = _kpar_kper(*self["kgrid"])
            return self[key]
        raise KeyError(key)


def finalize_device(out, lay, BoxSize, dims, counts_are_f64=False, k2d_on_device=False):
    """Finalise the accumulators on the device (pyl_pk_finalize) and return the reference's attribute arrays
    (same dict as `_finalize`).

    The finished block crosses PCIe ONCE into a pinned host block
    (_device.result_block) and the returned arrays are views of that block: they keep it out of circulation
    until the caller drops them.  No multi-MB NumPy temporaries: at 1024^3 their page faults alone
    cost several ms per call."""
    dev = out.device
    n2 = lay.n2d
    with torch.cuda.device(dev):
        f64 = out.view(torch.float64)
        kp = kq = None
        if k2d_on_device:         # kpar/kper written behind the accumulators by the kernel and shipped with them
            kp = f64[lay.total_words:lay.total_words + n2]
            kq = f64[lay.total_words + n2:lay.total_words + 2 * n2]
        else:
            f64 = f64[:lay.total_words]
        L.check(L.load().pyl_pk_finalize(D.ptr(out), int(dims), int(lay.fields), float(BoxSize),
                                         1 if counts_are_f64 else 0, D.ptr(kp) if kp is not None else None,
                                         D.ptr(kq) if kq is not None else None, D.stream_ptr(dev)),
                "pyl_pk_finalize")
        block, f = D.result_block(f64.numel())
        block[:f64.numel()].copy_(f64, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
    F, X = lay.fields, lay.xfields
    n3, n1 = lay.kmax + 1, lay.kmax_par + 1
    kF = frequencies(BoxSize, dims)[0]

    def arr(off, *shape, skip=0):
        return f[off:off + int(np.prod(shape))].reshape(shape)[skip:]

    o = _Spectra()
    Nm1 = arr(lay.Nm1D, n1, skip=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        o["k1D"] = ((Nm1 * np.arange(1, n1, dtype=np.float64)) / Nm1) * kF
    o["Nmodes1D"] = Nm1
    o["Pk1D"], o["PkX1D"] = arr(lay.Pk1D, n1, F, skip=1), arr(lay.PkX1D, n1, X, skip=1)
    # kpar/kper are pure index arithmetic (Pk_library.pyx:394-399): the result classes compute them on first access
    # (K2D) instead of shipping 16 bytes per 2D bin across PCIe with every result (95 MB at 4096^3)
    o["kgrid"] = (lay.kmax_par, lay.kmax_per, kF)
    if k2d_on_device:
        o["kpar"], o["kper"] = arr(lay.total_words, n2), arr(lay.total_words + n2, n2)
    o["Nmodes2D"], o["Pk2D"], o["PkX2D"] = arr(lay.Nm2D, n2), arr(lay.Pk2D, n2, F), arr(lay.PkX2D, n2, X)
    Nm3 = arr(lay.Nm3D, n3)
    check_number_modes(Nm3, dims)
    o["Nmodes3D"], o["k3D"] = Nm3[1:], arr(lay.k3D, n3, skip=1)
    o["Pk"], o["XPk"] = arr(lay.Pk3D, n3, 3, F, skip=1), arr(lay.PkX3D, n3, 3, X, skip=1)
    o["Pkphase"] = arr(lay.phase, n3, skip=1)
    return o


def take_dc(dk_list, holds_dc=True):
    """pyl_pk_take_dc: DC modes of the half-spectra as a float64 CUDA tensor (one per field), zeroed in the
    spectra.  With the transform of a DENSITY n this is dims^3 <n> (see `density=` of Pk / XPk)."""
    dev = dk_list[0].device
    dc = torch.empty(len(dk_list), dtype=torch.float64, device=dev)
    for b in range(0, len(dk_list), L.MAX_FIELDS):
        part = dk_list[b:b + L.MAX_FIELDS]
        ptrs = (ctypes.c_void_p * len(part))(*[D.ptr(t) for t in part])
        with torch.cuda.device(dev):
            L.check(L.load().pyl_pk_take_dc(ptrs, len(part), 1 if holds_dc else 0, D.ptr(dc[b:]), D.stream_ptr(dev)),
                    "pyl_pk_take_dc")
    return dc


def _offsets(offset, fields, dev):
    """`offset=` of Pk/XPk as a float64 CUDA tensor [fields] (None stays None): a number, a sequence of numbers
    and/or CUDA tensors (what field.prebias_ returned, one per field), or one tensor."""
    if offset is None:
        return None
    if torch.is_tensor(offset):
        t = offset.to(device=dev, dtype=torch.float64).reshape(-1)
    elif np.isscalar(offset):
        t = torch.full((fields,), float(offset), dtype=torch.float64, device=dev)
    else:
        t = torch.cat([o.to(device=dev, dtype=torch.float64).reshape(-1) if torch.is_tensor(o)
                       else torch.full((1,), float(o), dtype=torch.float64, device=dev) for o in offset])
    if t.numel() == 1 and fields > 1:
        t = t.expand(fields)
    if t.numel() != fields:
        raise ValueError("offset must give one constant per field")
    return t.contiguous()


def density_scale_(out, lay, dims, dc, offset=None):
    """pyl_pk_density_scale on the raw accumulators: sums of |FFT(n - c)|^2 become sums of |FFT(n/<n> - 1)|^2,
    <n> = c + dc/dims^3 (offset = c per field as a float64 CUDA tensor, None = 0)."""
    with torch.cuda.device(out.device):
        L.check(L.load().pyl_pk_density_scale(D.ptr(out), int(dims), int(lay.fields), D.ptr(dc),
                                              D.ptr(offset) if offset is not None else None,
                                              D.stream_ptr(out.device)), "pyl_pk_density_scale")


def spectra(dk_list, mas_index, dims, axis, BoxSize, want_phase=False, flags=0, density=False, offset=None):
    """bin + finalise: device finalisation for up to L.MAX_FIELDS fields, host assembly beyond.
    density=True: the fields are transforms of densities n (minus the constants `offset`); the spectra are those
    of n/<n> - 1."""
    if density and offset is None:
        raise ValueError("density=True needs offset=: the constant taken out of the grid before the deposit "
                         "(c = prebias_(grid, particles, W)); offset=0.0 transforms the raw density, whose float32 "
                         "FFT noise is ~1e-7 sqrt(cells)/sigma relative to the fluctuation modes")
    dc = take_dc(dk_list) if density else None
    off = _offsets(offset, len(dk_list), dk_list[0].device) if density else None
    if len(dk_list) <= L.MAX_FIELDS:
        out, lay = bin_device(dk_list, mas_index, dims, axis, want_phase and len(dk_list) == 1, flags=flags)
        if density:
            density_scale_(out, lay, dims, dc, off)
        return finalize_device(out, lay, BoxSize, dims)
    raw = bin_fields(dk_list, mas_index, dims, axis, want_phase, flags=flags)
    if density:
        inv = 1.0 / ((off.cpu().numpy() if off is not None else 0.0) + dc.cpu().numpy() / float(dims) ** 3)
        pairs = np.array([inv[i] * inv[j] for i in range(len(inv)) for j in range(i + 1, len(inv))])
        for k in ("Pk3D", "Pk1D", "Pk2D"):
            raw[k] = raw[k] * (inv * inv)
        for k in ("PkX3D", "PkX1D", "PkX2D"):
            raw[k] = raw[k] * pairs
    return _finalize(raw, BoxSize, dims)


# ---- host finalisation (vectorised restatement of :384-418 / :735-791) --------------------
_kpk_cache = {}


def _kpar_kper(kmax_par, kmax_per, kF):
    """Bin-centre kpar/kper of the 2D array (Pk_library.pyx:394-399); the index arithmetic is cached per
    geometry, callers get fresh copies."""
    key = (kmax_par, kmax_per, kF)
    hit = _kpk_cache.get(key)
    if hit is None:
        k_par = np.tile(np.arange(kmax_par + 1), kmax_per + 1)          # i2 % (kmax_par+1)
        k_per = np.repeat(np.arange(kmax_per + 1), kmax_par + 1)        # i2 // (kmax_par+1)
        hit = (0.5 * (k_par + k_par + 1) * kF, 0.5 * (k_per + k_per + 1) * kF)
        if len(_kpk_cache) > 8:
            _kpk_cache.clear()
        _kpk_cache[key] = hit
    return hit[0].copy(), hit[1].copy()


def _finalize(raw, BoxSize, dims):
    """Returns dict with the reference's attribute arrays, field/pair axes kept last."""
    kF, kN, kmax_par, kmax_per, kmax = frequencies(BoxSize, dims)
    fact = (BoxSize / dims ** 2) ** 3
    o = {}
    # 1D: discard the DC bin, give units, perpendicular-area weight
    Nm1 = raw["Nm1D"][1:]
    with np.errstate(invalid="ignore", divide="ignore"):
        k1D = (raw["k1D"][1:] / Nm1) * kF
        kmaxper = np.sqrt(kN ** 2 - k1D ** 2)
        w1 = (np.pi * kmaxper ** 2 / Nm1)
        o["Pk1D"] = (raw["Pk1D"][1:] * fact) * w1[:, None] / (2.0 * np.pi) ** 2
        o["PkX1D"] = (raw["PkX1D"][1:] * fact) * w1[:, None] / (2.0 * np.pi) ** 2
    o["k1D"], o["Nmodes1D"] = k1D, Nm1
    # 2D: DC bin kept
    o["kpar"], o["kper"] = _kpar_kper(kmax_par, kmax_per, kF)
    o["kgrid"] = (kmax_par, kmax_per, kF)
    with np.errstate(invalid="ignore", divide="ignore"):
        o["Pk2D"] = raw["Pk2D"] * fact / raw["Nm2D"][:, None]
        o["PkX2D"] = raw["PkX2D"] * fact / raw["Nm2D"][:, None]
    o["Nmodes2D"] = raw["Nm2D"]
    # 3D: check modes, discard the DC bin, (2l+1) factors, units
    check_number_modes(raw["Nm3D"], dims)
    Nm3 = raw["Nm3D"][1:]
    ell = np.array([1.0, 5.0, 9.0])[None, :, None]
    with np.errstate(invalid="ignore", divide="ignore"):
        o["k3D"] = (raw["k3D"][1:] / Nm3) * kF
        o["Pk"] = (raw["Pk3D"][1:] * ell / Nm3[:, None, None]) * fact
        o["XPk"] = (raw["PkX3D"][1:] * ell / Nm3[:, None, None]) * fact
        o["Pkphase"] = (raw["phase"][1:] / Nm3) * fact
    o["Nmodes3D"] = Nm3
    return o


class K2D:
    """kpar / kper of the 2D arrays (bin-centre coordinates, Pk_library.pyx:394-399): pure index arithmetic, computed
    on the host at first access from `_kgrid` = (kmax_par, kmax_per, kF) and kept."""
    _kgrid = None

    def _k2d(self):
        v = self.__dict__.get("_k2d_v")
        if v is None:
            v = self.__dict__["_k2d_v"] = _kpar_kper(*self._kgrid)
        return v

    @property
    def kpar(self):
        return self._k2d()[0]

    @property
    def kper(self):
        return self._k2d()[1]


class Pk(K2D):
    """1D, 2D and 3D power spectrum of a density field (Pk_library.pyx:263-420).

    Attributes: k3D, Pk (kmax,3: l=0,2,4), Nmodes3D, Pkphase, k1D, Pk1D, Nmodes1D, kpar, kper,
    Pk2D, Nmodes2D.

    density=True (not in the reference): `delta` is a DENSITY n (what MA deposited), and the result is the
    spectrum of n/<n> - 1 -- the caller's `delta /= np.mean(delta); delta -= 1` (Pk_snapshot.py:88-89) folded into
    the scale of the binned sums, <n> read from the DC mode, which is then dropped (its Pk2D[0] slot reads 0
    where the reference leaves the squared rounding residue of sum(delta)).
    offset=c (required with density=True): the grid holds n - c (the deposit started from -c:
    `c = prebias_(grid, particles, W)` instead of zeroing it), <n> = c + DC/dims^3.  A float32 transform carries
    rounding noise proportional to its largest amplitude; with the whole mass in the DC mode that is
    ~1e-7 sqrt(cells)/sigma relative to the fluctuation modes (3e-4 in power at 64^3 with 8 particles per cell,
    1e-2 at 1024^3), so the constant has to go before the transform.  offset=0.0 is accepted for small grids."""

    def __init__(self, delta, BoxSize, axis=2, MAS="CIC", threads=1, verbose=True, density=False, offset=None):
        start = time.time()
        if verbose:
            print("\nComputing power spectrum of the field...")
        D.require_cuda()
        if axis not in (0, 1, 2):
            raise ValueError("axis must be 0, 1 or 2")
        dev = D.pick_device(delta)
        delta_d = _as_delta(delta, dev)
        dims = delta_d.shape[0]
        delta_k = fft3d_r2c_device(delta_d)
        start2 = time.time()
        o = spectra([delta_k], [MAS_function(MAS)], dims, axis, BoxSize, want_phase=True, density=density,
                    offset=offset)
        if verbose:
            print("Time to complete loop = %.2f" % (time.time() - start2))
        self.k1D, self.Pk1D, self.Nmodes1D = o["k1D"], o["Pk1D"][:, 0], o["Nmodes1D"]
        self._kgrid = o["kgrid"]
        self.Pk2D, self.Nmodes2D = o["Pk2D"][:, 0], o["Nmodes2D"]
        self.k3D, self.Nmodes3D = o["k3D"], o["Nmodes3D"]
        self.Pk, self.Pkphase = np.ascontiguousarray(o["Pk"][:, :, 0]), o["Pkphase"]
        if verbose:
            print("Time taken = %.2f seconds" % (time.time() - start))


class XPk(K2D):
    """Auto- and cross-power spectra of several fields (Pk_library.pyx:529-793).

    Attributes: k3D, Nmodes3D, Pk (kmax,3,F), XPk (kmax,3,X), k1D, Nmodes1D, Pk1D (.,F),
    PkX1D (.,X), kpar, kper, Nmodes2D, Pk2D (.,F), PkX2D (.,X); pairs i<j in lexicographic order.
    density=True: the fields are densities, the spectra those of n_i/<n_i> - 1 (see Pk)."""

    def __init__(self, delta, BoxSize, axis=2, MAS=None, threads=1, density=False, offset=None):
        start = time.time()
        print("\nComputing power spectra of the fields...")
        D.require_cuda()
        if axis not in (0, 1, 2):
            raise ValueError("axis must be 0, 1 or 2")
        fields = len(delta)
        dims = len(delta[0])
        for i in range(1, fields):
            if len(delta[i]) != dims:
                reference_exit("Fields have different grid sizes!!!")
        if MAS is None or len(MAS) != fields:
            raise TypeError("MAS must be a list with one scheme per field")
        dev = D.pick_device(*delta)
        dk = [fft3d_r2c_device(_as_delta(d, dev, "delta[%d]" % i)) for i, d in enumerate(delta)]
        if D.is_cuda_tensor(delta[0]):
            torch.cuda.synchronize(dev)
        print("Time FFTS = %.2f" % (time.time() - start))
        start2 = time.time()
        o = spectra(dk, [MAS_function(m) for m in MAS], dims, axis, BoxSize, density=density, offset=offset)
        del dk
        print("Time loop = %.2f" % (time.time() - start2))
        self.k1D, self.Nmodes1D = o["k1D"], o["Nmodes1D"]
        self.Pk1D, self.PkX1D = o["Pk1D"], o["PkX1D"]
        self._kgrid, self.Nmodes2D = o["kgrid"], o["Nmodes2D"]
        self.Pk2D, self.PkX2D = o["Pk2D"], o["PkX2D"]
        self.k3D, self.Nmodes3D = o["k3D"], o["Nmodes3D"]
        self.Pk, self.XPk = o["Pk"], o["XPk"]
        print("Time taken = %.2f seconds" % (time.time() - start))


# ---- sibling estimators on the same machinery (Pk_plane, XPk_imag, XPk_plane, Pk_theta, XPk_dv, XPk_vv, XPk_2D,
# correct_MAS, expected_Pk, Xi, XXi, ...): defined in _pk_more.py, exported here under the reference's names
from . import _pk_more  # noqa: E402
from ._pk_more import *  # noqa: E402,F401,F403

__all__ += [n for n in _pk_more.__all__ if n != "field_smoothing"]
