"""Sibling estimators of Pylians3's `Pk_library` on the same GPU machinery.

Mirrors library/Pk_library/Pk_library.pyx, same names / positional order / defaults / printed messages / results:
  frequencies_2D :64-69, check_number_modes_2D :102-114, IFFT3Dr_f :149-163, FFT2Dr_f :181-194,
  class Pk_plane :435-511, class XPk_imag :811-1077, class XPk_plane :1093-1220, Pk_theta :1238-1326,
  XPk_dv :1345-1444, XPk_vv :1467-1580, XPk_2D :1754-1865, correct_MAS :1882-1939, expected_Pk :1956-2047,
  class Xi :2168-2282, class XXi :2298-2427, class XXi_multi :2443-2670
and smoothing_library.field_smoothing (library/smoothing_library/smoothing_library.pyx:215-235).

Every loop over modes or cells runs on the GPU (pyl_shell_bin, pyl_modes_*, pyl_pk_bin of include/pyl_b200.h; cuFFT for
the transforms); only the O(bins) unit conversions stay on the host, spelled in the reference's operation order.
Fields may be NumPy float32 arrays (copied to the device) or torch CUDA float32 tensors (zero-copy).  There is no CPU
path.  Imported into `Pk_library`'s namespace, so `PKL.Pk_theta(...)` etc. work as with the reference.
"""
import ctypes
import math
import time

import numpy as np
import torch

from . import _device as D
from . import _lib as L
from . import Pk_library as _P
from .errors import reference_exit

__all__ = ["frequencies_2D", "check_number_modes_2D", "IFFT3Dr_f", "FFT2Dr_f", "Pk_plane", "XPk_imag", "XPk_plane",
           "Pk_theta", "XPk_dv", "XPk_vv", "XPk_2D", "correct_MAS", "expected_Pk", "Xi", "XXi", "XXi_multi", "XXi_projected", "field_smoothing"]


def frequencies_2D(BoxSize, dims):
    """Pk_library.pyx:64-69."""
    kF = 2.0 * np.pi / BoxSize
    middle = dims // 2
    kN = middle * kF
    kmax_par = middle
    kmax_per = middle
    kmax = int(np.sqrt(middle ** 2 + middle ** 2))
    return kF, kN, kmax_par, kmax_per, kmax


def check_number_modes_2D(Nmodes, dims):
    """Pk_library.pyx:102-114."""
    own_modes = 1 if dims % 2 == 1 else 4
    repeated_modes = (dims ** 2 - own_modes) // 2
    indep_modes = repeated_modes + own_modes
    if int(np.sum(Nmodes)) != indep_modes:
        reference_exit("WARNING: Not all modes counted",
                       "Counted  %d independent modes" % (int(np.sum(Nmodes))),
                       "Expected %d independent modes" % indep_modes)


# ---- device stages ------------------------------------------------------------------------------------------
def _as_image(a, dev, name):
    if getattr(a, "ndim", None) != 2:
        raise ValueError("Buffer has wrong number of dimensions (expected 2, got %s)" % getattr(a, "ndim", "?"))
    t, _ = D.to_device_f32(a, dev, name)
    if t.shape[0] != t.shape[1]:
        raise ValueError("%s must be a (grid,grid) image, got %s" % (name, tuple(t.shape)))
    return t


def fft2d_r2c_device(img_d):
    """(grid,grid) float32 CUDA tensor -> (grid,grid//2+1) complex64 CUDA tensor: one 2D r2c of cuFFT."""
    lib = L.load()
    grid = img_d.shape[0]
    dev = img_d.device
    with torch.cuda.device(dev):
        out = torch.empty((grid, grid // 2 + 1), dtype=torch.complex64, device=dev)
        need = lib.pyl_fft_slab_workspace_bytes(grid, 1, 0)
        if need == ctypes.c_size_t(-1).value:
            L.check(-3, "pyl_fft_slab_workspace_bytes")
        ws = D.workspace(need, dev, "fft")
        L.check(lib.pyl_fft_slab_yz(D.ptr(img_d), D.ptr(out), grid, 1, D.ptr(ws), need, D.stream_ptr(dev)), "pyl_fft_slab_yz")
    return out


def ifft3d_c2r_device(ak_d, normalise=True):
    """(dims,dims,dims//2+1) complex64 CUDA tensor -> (dims,dims,dims) float32 CUDA tensor.  `ak_d` is consumed
    (cuFFT's c2r may overwrite its input).  normalise=True gives IFFT3Dr_f's result (pyfftw scales by 1/dims^3)."""
    lib = L.load()
    dims = ak_d.shape[0]
    dev = ak_d.device
    with torch.cuda.device(dev):
        out = torch.empty((dims, dims, dims), dtype=torch.float32, device=dev)
        need = lib.pyl_fft_c2r_workspace_bytes(dims)
        if need == ctypes.c_size_t(-1).value:
            L.check(-3, "pyl_fft_c2r_workspace_bytes")
        ws = D.workspace(need, dev, "fft")
        L.check(lib.pyl_fft_c2r(D.ptr(ak_d), D.ptr(out), dims, D.ptr(ws), need, D.stream_ptr(dev)), "pyl_fft_c2r")
        if normalise:
            L.check(lib.pyl_scale_inplace(D.ptr(out), out.numel(), float(np.float32(1.0) / np.float32(dims ** 3)),
                                          D.stream_ptr(dev)), "pyl_scale_inplace")
    return out


def FFT2Dr_f(a, threads=1):
    """Pk_library.pyx:181-194."""
    D.require_cuda()
    dev = D.pick_device(a)
    out = fft2d_r2c_device(_as_image(a, dev, "a"))
    return out if D.is_cuda_tensor(a) else out.cpu().numpy()


def IFFT3Dr_f(a, threads=1):
    """Pk_library.pyx:149-163: normalised inverse c2r of a (dims,dims,dims//2+1) complex64 array."""
    D.require_cuda()
    dev = D.pick_device(a)
    if isinstance(a, torch.Tensor):
        ak = a.to(dev).to(torch.complex64).contiguous().clone()
    else:
        ak = torch.from_numpy(np.ascontiguousarray(a, dtype=np.complex64)).to(dev)
    out = ifft3d_c2r_device(ak)
    return out if D.is_cuda_tensor(a) else out.cpu().numpy()


def shell_bin(kind, fields, mas_index, dims, axis=2, scale=1.0, table=None):
    """pyl_shell_bin on CUDA tensors; returns the RAW sums on the host: dict(ksum, Nm, vals=[...])."""
    lib = L.load()
    kid = L.SHELL_KINDS[kind]
    nb, nv = ctypes.c_int(0), ctypes.c_int(0)
    L.check(lib.pyl_shell_layout(kid, int(dims), ctypes.byref(nb), ctypes.byref(nv)), "pyl_shell_layout")
    nb, nv = nb.value, nv.value
    dev = fields[0].device if fields else torch.device("cuda", torch.cuda.current_device())
    with torch.cuda.device(dev):
        out = torch.empty((2 + nv) * nb, dtype=torch.int64, device=dev)
        need = lib.pyl_shell_bin_workspace_bytes(kid, int(dims))
        ws = D.workspace(need, dev, "shell")
        nf = len(fields)
        ptrs = (ctypes.c_void_p * max(nf, 1))(*[D.ptr(t) for t in fields])
        mi = (ctypes.c_int * 2)(*(list(mas_index) + [0, 0])[:2])
        tab = None
        if table is not None:
            tk, tP, kF32, log10_kmin, deltak = table
            tab = L.ShellTable(D.ptr(tk), D.ptr(tP), int(tk.numel()), float(kF32), float(log10_kmin), float(deltak))
        L.check(lib.pyl_shell_bin(kid, ptrs, nf, mi, int(dims), int(axis), float(scale),
                                  None if tab is None else ctypes.addressof(tab), D.ptr(out), D.ptr(ws), need,
                                  D.stream_ptr(dev)), "pyl_shell_bin")
        words = D.to_host_numpy(out)
    f64 = words.view(np.float64)
    return dict(ksum=f64[:nb].copy(), Nm=words[nb:2 * nb].astype(np.float64),
                vals=[f64[(2 + j) * nb:(3 + j) * nb].copy() for j in range(nv)])


def _modes(op, a_k, b_k, dims, mas_a, mas_b):
    lib = L.load()
    dev = a_k.device
    with torch.cuda.device(dev):
        need = lib.pyl_modes_workspace_bytes(int(dims))
        ws = D.workspace(need, dev, "modes")
        if op == "deconvolve":
            st = lib.pyl_modes_deconvolve(D.ptr(a_k), int(dims), int(mas_a), D.ptr(ws), need, D.stream_ptr(dev))
        else:
            st = lib.pyl_modes_power(D.ptr(a_k), D.ptr(b_k), int(dims), int(mas_a), int(mas_b), D.ptr(ws), need,
                                     D.stream_ptr(dev))
    L.check(st, "pyl_modes_" + op)


def _cube(x, dev, name):
    return _P._as_delta(x, dev, name)


# ---- 2D images ----------------------------------------------------------------------------------------------
class Pk_plane:
    """Power spectrum of a 2D field (Pk_library.pyx:435-511).  Attributes: k, Pk, Nmodes."""

    def __init__(self, delta, BoxSize, MAS="CIC", threads=1, verbose=True):
        start = time.time()
        if verbose:
            print("\nComputing power spectrum of the field...")
        D.require_cuda()
        dev = D.pick_device(delta)
        img = _as_image(delta, dev, "delta")
        grid = img.shape[0]
        kF, kN, kmax_par, kmax_per, kmax = frequencies_2D(BoxSize, grid)
        delta_k = fft2d_r2c_device(img)
        start2 = time.time()
        r = shell_bin("plane", [delta_k], [_P.MAS_function(MAS)], grid)
        if verbose:
            print("Time to complete loop = %.2f" % (time.time() - start2))
        check_number_modes_2D(r["Nm"], grid)
        k2D, Nmodes, Pk2D = r["ksum"][1:], r["Nm"][1:], r["vals"][0][1:]
        self.k = (k2D / Nmodes) * kF                                       # :506
        self.Pk = (Pk2D / Nmodes) * (BoxSize / grid ** 2) ** 2              # :507
        self.Nmodes = Nmodes
        if verbose:
            print("Time taken = %.2f seconds" % (time.time() - start))


class XPk_plane:
    """Auto- and cross-power spectra of two images (Pk_library.pyx:1093-1220).  Attributes: k, Nmodes, Pk (.,2),
    XPk, r."""

    def __init__(self, delta1, delta2, BoxSize, MAS1=None, MAS2=None, threads=1):
        start = time.time()
        print("\nComputing power spectra of the fields...")
        D.require_cuda()
        if delta1.shape[0] != delta2.shape[1]:
            raise Exception("Images have different grid sizes!!!")
        dev = D.pick_device(delta1, delta2)
        grid = delta1.shape[0]
        kF, kN, kmax_par, kmax_per, kmax = frequencies_2D(BoxSize, grid)
        d1 = fft2d_r2c_device(_as_image(delta1, dev, "delta1"))
        d2 = fft2d_r2c_device(_as_image(delta2, dev, "delta2"))
        print("Time FFTS = %.2f" % (time.time() - start))
        start2 = time.time()
        r = shell_bin("xplane", [d1, d2], [_P.MAS_function(MAS1), _P.MAS_function(MAS2)], grid)
        print("Time loop = %.2f" % (time.time() - start2))
        fact = (BoxSize / grid ** 2) ** 2
        Nm = r["Nm"][1:]
        self.k = (r["ksum"][1:] / Nm) * kF                                  # :1211
        self.Nmodes = Nm
        self.Pk = np.stack([(r["vals"][0][1:] / Nm) * fact, (r["vals"][1][1:] / Nm) * fact], axis=1)
        self.XPk = (r["vals"][2][1:] / Nm) * fact
        self.r = self.XPk / np.sqrt(self.Pk[:, 0] * self.Pk[:, 1])
        print("Time taken = %.2f seconds" % (time.time() - start))


# ---- 3D: variants of XPk --------------------------------------------------------------------------------------
class XPk_imag(_P.K2D):
    """XPk with the cross term imag_i*real_j - real_i*imag_j (Pk_library.pyx:811-1077); same attributes as XPk."""

    def __init__(self, delta, BoxSize, axis=2, MAS=None, threads=1):
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
        dk = [_P.fft3d_r2c_device(_cube(d, dev, "delta[%d]" % i)) for i, d in enumerate(delta)]
        print("Time FFTS = %.2f" % (time.time() - start))
        start2 = time.time()
        o = _P.spectra(dk, [_P.MAS_function(m) for m in MAS], dims, axis, BoxSize, flags=L.PK_CROSS_IMAG)
        del dk
        print("Time loop = %.2f" % (time.time() - start2))
        self.k1D, self.Nmodes1D, self.Pk1D, self.PkX1D = o["k1D"], o["Nmodes1D"], o["Pk1D"], o["PkX1D"]
        self._kgrid, self.Nmodes2D = o["kgrid"], o["Nmodes2D"]
        self.Pk2D, self.PkX2D = o["Pk2D"], o["PkX2D"]
        self.k3D, self.Nmodes3D, self.Pk, self.XPk = o["k3D"], o["Nmodes3D"], o["Pk"], o["XPk"]
        print("Time taken = %.2f seconds" % (time.time() - start))


def XPk_2D(delta1, delta2, BoxSize, axis=2, MAS1="CIC", MAS2="CIC", threads=1):
    """2D (k_par, k_per) auto- and cross-spectra of two fields (Pk_library.pyx:1754-1865).
    Returns [kpar, kper, Pk1, Pk2, PkX, Nmodes]."""
    start = time.time()
    print("Computing power spectra of the fields...")
    D.require_cuda()
    dims = len(delta1)
    if dims != len(delta2):
        reference_exit("Different grids in the two fields!!!")
    dev = D.pick_device(delta1, delta2)
    dk = [_P.fft3d_r2c_device(_cube(delta1, dev, "delta1")), _P.fft3d_r2c_device(_cube(delta2, dev, "delta2"))]
    start2 = time.time()
    out, lay = _P.bin_device(dk, [_P.MAS_function(MAS1), _P.MAS_function(MAS2)], dims, axis)
    o = _P.finalize_device(out, lay, BoxSize, dims)          # Pk2D = sum*fact/Nmodes2D, the expression of :1860-1862
    print("Time compute modulus = %.2f" % (time.time() - start2))
    kpar, kper = _P._kpar_kper(*o["kgrid"])
    res = [kpar, kper, np.ascontiguousarray(o["Pk2D"][:, 0]), np.ascontiguousarray(o["Pk2D"][:, 1]),
           np.ascontiguousarray(o["PkX2D"][:, 0]), o["Nmodes2D"]]
    print("Time taken = %.2f seconds" % (time.time() - start))
    return res


# ---- 3D: velocity-divergence estimators -----------------------------------------------------------------------
def _momentum_(V, delta_d, dev, names):
    """V[i] *= (1 + delta) in float32 on the device; host arrays receive the product back, like the reference's
    in-place NumPy statement (Pk_library.pyx:1367)."""
    lib = L.load()
    out = []
    for v, name in zip(V, names):
        v_d = _cube(v, dev, name)
        if v_d.shape != delta_d.shape:
            raise ValueError("operands could not be broadcast together with shapes %s %s" % (tuple(v_d.shape), tuple(delta_d.shape)))
        with torch.cuda.device(dev):
            L.check(lib.pyl_mul_one_plus(D.ptr(v_d), D.ptr(delta_d), v_d.numel(), D.stream_ptr(dev)), "pyl_mul_one_plus")
        if isinstance(v, np.ndarray):
            v[...] = v_d.cpu().numpy()
        elif isinstance(v, torch.Tensor) and v.data_ptr() != v_d.data_ptr():
            v.copy_(v_d)
        out.append(v_d)
    return out


def Pk_theta(Vx, Vy, Vz, BoxSize, axis=2, MAS="CIC", threads=1):
    """Power spectrum of theta = div V (Pk_library.pyx:1238-1326).  Returns k, Pk, Nmodes."""
    start = time.time()
    print("Computing power spectrum of theta...")
    D.require_cuda()
    dev = D.pick_device(Vx, Vy, Vz)
    dims = len(Vx)
    kF, kN, kmax_par, kmax_per, kmax = _P.frequencies(BoxSize, dims)
    Vk = [_P.fft3d_r2c_device(_cube(v, dev, n)) for v, n in zip((Vx, Vy, Vz), ("Vx", "Vy", "Vz"))]
    start2 = time.time()
    r = shell_bin("theta", Vk, [_P.MAS_function(MAS)], dims)
    print("Time compute modulus = %.2f" % (time.time() - start2))
    _P.check_number_modes(r["Nm"], dims)
    k = r["ksum"][1:]; Nmodes = r["Nm"][1:]; k = (k / Nmodes) * kF
    Pk = r["vals"][0][1:] * (BoxSize / dims ** 2) ** 3 * kF ** 2; Pk *= (1.0 / Nmodes)          # :1322-1323
    print("Time taken = %.2f seconds" % (time.time() - start))
    return k, Pk, Nmodes


def XPk_dv(delta, Vx, Vy, Vz, BoxSize, axis=2, MAS="CIC", threads=1):
    """Auto- and cross-spectra of delta and div[(1+delta) V] (Pk_library.pyx:1345-1444).  Like the reference,
    Vx, Vy, Vz are multiplied by (1+delta) IN PLACE.  Returns k, Pk1, Pk2, PkX, Nmodes."""
    start = time.time()
    print("Computing power spectra of the fields...")
    D.require_cuda()
    dev = D.pick_device(delta, Vx, Vy, Vz)
    dims = len(delta)
    kF, kN, kmax_par, kmax_per, kmax = _P.frequencies(BoxSize, dims)
    delta_d = _cube(delta, dev, "delta")
    V = _momentum_([Vx, Vy, Vz], delta_d, dev, ("Vx", "Vy", "Vz"))
    fk = [_P.fft3d_r2c_device(delta_d)] + [_P.fft3d_r2c_device(v) for v in V]
    start2 = time.time()
    r = shell_bin("dv", fk, [_P.MAS_function(MAS)], dims)
    print("Time compute modulus = %.2f" % (time.time() - start2))
    _P.check_number_modes(r["Nm"], dims)
    k = r["ksum"][1:]; Nmodes = r["Nm"][1:]; k = (k / Nmodes) * kF
    Pk1 = r["vals"][0][1:] * (BoxSize / dims ** 2) ** 3; Pk1 *= (1.0 / Nmodes)                  # :1439-1441
    Pk2 = r["vals"][1][1:] * (BoxSize / dims ** 2) ** 3 * kF ** 2; Pk2 *= (1.0 / Nmodes)
    PkX = r["vals"][2][1:] * (BoxSize / dims ** 2) ** 3 * kF; PkX *= (1.0 / Nmodes)
    print("Time taken = %.2f seconds" % (time.time() - start))
    return k, Pk1, Pk2, PkX, Nmodes


def XPk_vv(delta1, Vx1, Vy1, Vz1, delta2, Vx2, Vy2, Vz2, BoxSize, axis=2, MAS="CIC", threads=1):
    """Auto- and cross-spectra of two momentum divergences (Pk_library.pyx:1467-1580); the velocity fields are
    multiplied by (1+delta) in place.  The reference also transforms and deconvolves delta1 and delta2 but never
    reads them (:1539,1543 vs :1549-1568), so those two transforms are not done.  Returns k, Pk1, Pk2, PkX, Nmodes."""
    start = time.time()
    print("Computing power spectra of the fields...")
    D.require_cuda()
    dev = D.pick_device(delta1, Vx1, delta2, Vx2)
    dims = len(delta1)
    assert len(delta2) == dims
    kF, kN, kmax_par, kmax_per, kmax = _P.frequencies(BoxSize, dims)
    V1 = _momentum_([Vx1, Vy1, Vz1], _cube(delta1, dev, "delta1"), dev, ("Vx1", "Vy1", "Vz1"))
    V2 = _momentum_([Vx2, Vy2, Vz2], _cube(delta2, dev, "delta2"), dev, ("Vx2", "Vy2", "Vz2"))
    fk = [_P.fft3d_r2c_device(v) for v in V1 + V2]
    start2 = time.time()
    r = shell_bin("vv", fk, [_P.MAS_function(MAS)], dims)
    print("Time compute modulus = %.2f" % (time.time() - start2))
    _P.check_number_modes(r["Nm"], dims)
    k = r["ksum"][1:]; Nmodes = r["Nm"][1:]; k = (k / Nmodes) * kF
    res = []
    for j in range(3):                                                                         # :1575-1577
        P = r["vals"][j][1:] * (BoxSize / dims ** 2) ** 3 * kF ** 2; P *= (1.0 / Nmodes)
        res.append(P)
    print("Time taken = %.2f seconds" % (time.time() - start))
    return k, res[0], res[1], res[2], Nmodes


# ---- expected_Pk ------------------------------------------------------------------------------------------------
def _expected_table(k_in, Pk_in, bins):
    """Pk_library.pyx:1971-1993: the log-spaced float32 table the mode loop interpolates in (O(bins), host)."""
    f32 = np.float32
    kmin_in, kmax_in = k_in[0], k_in[-1]
    deltak = f32((math.log10(float(kmax_in)) - math.log10(float(kmin_in))) / (bins - 1.0))
    tk, tP = np.zeros(bins, f32), np.zeros(bins, f32)
    j = 1
    for i in range(bins):
        tk[i] = f32(10.0 ** (math.log10(float(kmin_in)) + float(f32(deltak) * f32(i))))
        while tk[i] > k_in[j] and j < len(k_in) - 1:
            j += 1
        tP[i] = (Pk_in[j] - Pk_in[j - 1]) / (k_in[j] - k_in[j - 1]) * (tk[i] - k_in[j - 1]) + Pk_in[j - 1]
    return tk, tP, kmin_in, deltak


def expected_Pk(k_in, Pk_in, BoxSize, dims, bins=750):
    """Binned k, Pk a (dims^3, BoxSize) grid would measure for an input Pk (Pk_library.pyx:1956-2047).
    Returns k3D, Pk3D, Nmodes3D."""
    start2 = time.time()
    D.require_cuda()
    k_in = np.asarray(k_in)
    Pk_in = np.asarray(Pk_in)
    if k_in.dtype != np.float32 or Pk_in.dtype != np.float32:
        raise ValueError("Buffer dtype mismatch, expected 'float32_t'")
    BoxSize = float(np.float32(BoxSize))                      # `float BoxSize`
    kF, kN, kmax_par, kmax_per, kmax = _P.frequencies(BoxSize, dims)
    if np.any(k_in[1:] <= k_in[:-1]):
        raise Exception("Input k-array not sorted!!!")
    kF32 = np.float32(kF)                                     # `cdef float kF`
    if kF32 < k_in[0] or kmax * float(kF32) > float(k_in[-1]):
        raise Exception("k value in grid outside input k range")
    tk, tP, kmin_in, deltak = _expected_table(k_in, Pk_in, int(bins))
    dev = torch.device("cuda", torch.cuda.current_device())
    table = (torch.from_numpy(tk).to(dev), torch.from_numpy(tP).to(dev), kF32, math.log10(float(kmin_in)),
             float(deltak))
    r = shell_bin("expected", [], [], int(dims), table=table)
    k3D, Nmodes3D, Pk3D = r["ksum"][1:], r["Nm"][1:], r["vals"][0][1:]
    k3D = k3D / Nmodes3D
    Pk3D = Pk3D / Nmodes3D
    print("Time take = %.2f" % (time.time() - start2))
    return k3D, Pk3D, Nmodes3D


# ---- deconvolved field, correlation functions, smoothing ----------------------------------------------------------
def correct_MAS(delta, BoxSize, MAS="CIC", threads=1):
    """FFT, MAS deconvolution of the modes, inverse FFT (Pk_library.pyx:1882-1939).  Returns the corrected field
    (NumPy in -> NumPy out, CUDA tensor in -> CUDA tensor out); `delta` itself is not modified."""
    start = time.time()
    print("\nComputing power spectrum of the field...")
    D.require_cuda()
    dev = D.pick_device(delta)
    delta_d = _cube(delta, dev, "delta")
    dims = delta_d.shape[0]
    delta_k = _P.fft3d_r2c_device(delta_d)
    start2 = time.time()
    _modes("deconvolve", delta_k, None, dims, _P.MAS_function(MAS), 0)
    print("Time to complete loop = %.2f" % (time.time() - start2))
    out = ifft3d_c2r_device(delta_k)
    print("Time taken = %.2f seconds" % (time.time() - start))
    return out if D.is_cuda_tensor(delta) else out.cpu().numpy()


def _xi_results(self, xi_k, dims, BoxSize, axis):
    """c2r of the (|delta_k|^2, 0) modes, radial l = 0,2,4 binning of every cell, units (:2221-2280)."""
    xi_grid = ifft3d_c2r_device(xi_k, normalise=False)
    # the 1/dims^3 of the normalised inverse transform is applied to each value as it is read (same float32 product)
    scale = float(np.float32(1.0) / np.float32(dims ** 3))
    r = shell_bin("xi", [xi_grid], [], dims, axis=axis, scale=scale)
    Nm = r["Nm"][1:]
    self.r3D = (r["ksum"][1:] / Nm) * (BoxSize * 1.0 / dims)                                   # :2275
    self.Nmodes3D = Nm
    xi0 = (r["vals"][0][1:] / Nm) * (1.0 / dims ** 3)                                          # :2276-2278
    xi2 = (r["vals"][1][1:] * 5.0 / Nm) * (1.0 / dims ** 3)
    xi4 = (r["vals"][2][1:] * 9.0 / Nm) * (1.0 / dims ** 3)
    self.xi = np.stack([xi0, xi2, xi4], axis=1)


class Xi:
    """Correlation function multipoles of a field (Pk_library.pyx:2168-2282).  Attributes: r3D, xi (.,3), Nmodes3D."""

    def __init__(self, delta, BoxSize, MAS="CIC", axis=2, threads=1):
        start = time.time()
        print("\nComputing correlation function of the field...")
        D.require_cuda()
        if axis not in (0, 1, 2):
            raise ValueError("axis must be 0, 1 or 2")
        BoxSize = float(np.float32(BoxSize))                  # `float BoxSize` in the reference's signature
        dev = D.pick_device(delta)
        delta_d = _cube(delta, dev, "delta")
        dims = delta_d.shape[0]
        delta_k = _P.fft3d_r2c_device(delta_d)
        _modes("power", delta_k, None, dims, _P.MAS_function(MAS), 0)
        start2 = time.time()
        _xi_results(self, delta_k, dims, BoxSize, axis)
        print("Time to complete loop = %.2f" % (time.time() - start2))
        print("Time taken = %.2f seconds" % (time.time() - start))


class XXi:
    """Cross-correlation function multipoles of two fields (Pk_library.pyx:2298-2427)."""

    def __init__(self, delta1, delta2, BoxSize, MAS=["CIC", "CIC"], axis=2, threads=1):
        start = time.time()
        print("\nComputing correlation function of the field...")
        D.require_cuda()
        if axis not in (0, 1, 2):
            raise ValueError("axis must be 0, 1 or 2")
        BoxSize = float(np.float32(BoxSize))
        grid = delta1.shape[0]
        if grid != delta2.shape[0]:
            raise Exception("grid sizes differ!!!")
        dev = D.pick_device(delta1, delta2)
        d1 = _P.fft3d_r2c_device(_cube(delta1, dev, "delta1"))
        d2 = _P.fft3d_r2c_device(_cube(delta2, dev, "delta2"))
        _modes("power", d1, d2, grid, _P.MAS_function(MAS[0]), _P.MAS_function(MAS[1]))
        del d2
        start2 = time.time()
        _xi_results(self, d1, grid, BoxSize, axis)
        print("Time to complete loop = %.2f" % (time.time() - start2))
        print("Time taken = %.2f seconds" % (time.time() - start))


class XXi_projected:
    """Projected cross-correlation function of two images (Pk_library.pyx:2684-2789).  Attributes: r_p, xi_p,
    Nmodes_p.  2D r2c of both images -> pyl_modes_power_2d -> 2D c2r -> pyl_radial_bin_2d, all on the device."""

    def __init__(self, delta1, delta2, BoxSize, MAS=["CIC", "CIC"], threads=1):
        start = time.time()
        print("\nComputing correlation function of the field...")
        D.require_cuda()
        BoxSize = float(np.float32(BoxSize))
        grid = delta1.shape[0]
        if grid != delta2.shape[0]:
            raise Exception("grid sizes differ!!!")
        dev = D.pick_device(delta1, delta2)
        lib = L.load()
        d1 = fft2d_r2c_device(_as_image(delta1, dev, "delta1"))
        d2 = fft2d_r2c_device(_as_image(delta2, dev, "delta2"))
        with torch.cuda.device(dev):
            need = lib.pyl_modes_workspace_bytes(int(grid))
            ws = D.workspace(need, dev, "modes")
            L.check(lib.pyl_modes_power_2d(D.ptr(d1), D.ptr(d2), int(grid), _P.MAS_function(MAS[0]),
                                           _P.MAS_function(MAS[1]), D.ptr(ws), need, D.stream_ptr(dev)),
                    "pyl_modes_power_2d")
            del d2
            xi_grid = torch.empty((grid, grid), dtype=torch.float32, device=dev)
            need = lib.pyl_fft2d_c2r_workspace_bytes(int(grid))
            if need == ctypes.c_size_t(-1).value:
                L.check(-3, "pyl_fft2d_c2r_workspace_bytes")
            ws = D.workspace(need, dev, "fft")
            L.check(lib.pyl_fft2d_c2r(D.ptr(d1), D.ptr(xi_grid), int(grid), D.ptr(ws), need, D.stream_ptr(dev)),
                    "pyl_fft2d_c2r")
            start2 = time.time()
            nb = int((grid // 2) * np.sqrt(2)) + 1
            out = torch.empty(3 * nb, dtype=torch.int64, device=dev)
            # the 1/grid^2 of the normalised inverse transform is applied to each value as it is read
            L.check(lib.pyl_radial_bin_2d(D.ptr(xi_grid), int(grid), float(np.float32(1.0) / np.float32(grid ** 2)),
                                          D.ptr(out), D.stream_ptr(dev)), "pyl_radial_bin_2d")
            words = D.to_host_numpy(out).copy()
        print("Time to complete loop = %.2f" % (time.time() - start2))
        f64 = words.view(np.float64)
        Nm = words[nb + 1:2 * nb].astype(np.float64)
        with np.errstate(invalid="ignore", divide="ignore"):
            self.r_p = (f64[1:nb] / Nm) * (BoxSize * 1.0 / grid)                 # :2783
            self.xi_p = (f64[2 * nb + 1:3 * nb] / Nm) * (1.0 / grid ** 2)         # :2784
        self.Nmodes_p = Nm
        print("Time taken = %.2f seconds" % (time.time() - start))


class XXi_multi:
    """Auto- and cross-correlation function multipoles of several fields (Pk_library.pyx:2443-2670).
    Attributes: r3D, Nmodes3D, xi (.,3,F), Xxi (.,3,X); pairs i<j in lexicographic order.  As in the reference, the
    cross term (i, j) is deconvolved with field i's window for BOTH fields (:2548-2553)."""

    def __init__(self, delta, BoxSize, axis=2, MAS=None, threads=1):
        start = time.time()
        print("\nComputing correlation functions of the fields...")
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
        dk = [_P.fft3d_r2c_device(_cube(d, dev, "delta[%d]" % i)) for i, d in enumerate(delta)]
        print("Time FFTS = %.2f" % (time.time() - start))
        mi = [_P.MAS_function(m) for m in MAS]

        class _Holder:
            pass
        cols, xcols = [], []
        for i in range(fields):                                    # autos: |delta_i|^2 with its own window
            a = torch.clone(dk[i])
            _modes("power", a, None, dims, mi[i], 0)
            h = _Holder()
            _xi_results(h, a, dims, BoxSize, axis)
            cols.append(h)
        for i in range(fields):                                    # crosses, both fields with window i
            for j in range(i + 1, fields):
                a = torch.clone(dk[i])
                _modes("power", a, dk[j], dims, mi[i], mi[i])
                h = _Holder()
                _xi_results(h, a, dims, BoxSize, axis)
                xcols.append(h)
        del dk
        self.r3D, self.Nmodes3D = cols[0].r3D, cols[0].Nmodes3D
        self.xi = np.stack([h.xi for h in cols], axis=2)
        self.Xxi = np.stack([h.xi for h in xcols], axis=2) if xcols else np.zeros(self.xi.shape[:2] + (0,))
        print("Time taken = %.2f seconds" % (time.time() - start))


def field_smoothing(field, filter_k, threads=1):
    """smoothing_library.field_smoothing (smoothing_library.pyx:215-235): IFFT(FFT(field) * filter_k).
    `filter_k`: (dims,dims,dims//2+1) complex64 NumPy array or CUDA tensor."""
    D.require_cuda()
    if field.shape[0] != filter_k.shape[0]:
        raise Exception("field and filter have different grids!!!")
    dev = D.pick_device(field, filter_k)
    field_d = _cube(field, dev, "field")
    dims = field_d.shape[0]
    if isinstance(filter_k, torch.Tensor):
        if filter_k.dtype != torch.complex64:
            raise ValueError("filter_k must be complex64")
        fk = filter_k.to(dev).contiguous()
    else:
        fk = np.asarray(filter_k)
        if fk.dtype != np.complex64:
            raise ValueError("Buffer dtype mismatch, expected 'complex64_t' but got '%s'" % fk.dtype)
        fk = torch.from_numpy(np.ascontiguousarray(fk)).to(dev)
    if tuple(fk.shape) != (dims, dims, dims // 2 + 1):
        raise ValueError("filter_k must have shape (dims,dims,dims//2+1)")
    field_k = _P.fft3d_r2c_device(field_d)
    with torch.cuda.device(dev):
        L.check(L.load().pyl_cmul_inplace(D.ptr(field_k), D.ptr(fk), field_k.numel(), D.stream_ptr(dev)), "pyl_cmul_inplace")
    out = ifft3d_c2r_device(field_k)
    return out if D.is_cuda_tensor(field) else out.cpu().numpy()
