"""Python face of the CPU oracle for the sibling estimators (TEST INFRASTRUCTURE ONLY).

Restates, with the reference's names, arguments and return values:
  Pk_plane   Pk_library.pyx:435-511      XPk_plane  :1093-1220     XPk_imag  :811-1077
  Pk_theta   :1238-1326    XPk_dv :1345-1444    XPk_vv :1467-1580   XPk_2D    :1754-1865
  correct_MAS :1882-1939   expected_Pk :1956-2047   Xi :2168-2282   XXi :2298-2427
  frequencies_2D :64-69    check_number_modes_2D :102-114
and smoothing_library.field_smoothing (smoothing_library.pyx:215-235).
Hot loops live in pk_oracle.c; transforms are scipy's pocketfft in float32, forward
unnormalised, inverse normalised (what pyfftw gives the reference, see pyfftw_shim).
Pinned against tests/golden/pk_more_golden.npz (outputs of the compiled reference).
"""
import ctypes
import math

import numpy as np
import scipy.fft as _sfft

from . import cpu as _c

_fpp = ctypes.POINTER(ctypes.c_float)
_dpp = ctypes.POINTER(ctypes.c_double)
_ipp = ctypes.POINTER(ctypes.c_int)
_ready = False


def _lib():
    global _ready
    L = _c.lib()
    if not _ready:
        i, f = ctypes.c_int, ctypes.c_float
        L.oracle_pk_bin_imag.argtypes = [_fpp, i, i, _ipp, i, i, i, i] + [_dpp] * 11
        L.oracle_plane_bin.argtypes = [_fpp, i, i, _ipp] + [_dpp] * 4
        L.oracle_vel_bin.argtypes = [i, _fpp, i, i] + [_dpp] * 5
        L.oracle_expected_pk.argtypes = [i, f, _fpp, _fpp, f, f] + [_dpp] * 3
        L.oracle_correct_mas_modes.argtypes = [_fpp, i, i]
        L.oracle_xi_modes.argtypes = [_fpp, _fpp, i, i, i]
        L.oracle_xi_bin.argtypes = [_fpp, i, i] + [_dpp] * 3
        for n in ("oracle_pk_bin_imag", "oracle_plane_bin", "oracle_vel_bin", "oracle_expected_pk",
                  "oracle_correct_mas_modes", "oracle_xi_modes", "oracle_xi_bin"):
            getattr(L, n).restype = None
        _ready = True
    return L


def _cf(a):
    """complex64 ndarray -> float* of its interleaved storage"""
    return a.view(np.float32).ctypes.data_as(_fpp)


def _z(*s):
    return np.zeros(s, dtype=np.float64)


def frequencies_2D(BoxSize, dims):
    """Pk_library.pyx:64-69."""
    kF = 2.0 * np.pi / BoxSize
    middle = dims // 2
    kN = middle * kF
    return kF, kN, middle, middle, int(np.sqrt(middle ** 2 + middle ** 2))


def expected_modes_2D(dims):
    """Pk_library.pyx:102-108."""
    own = 1 if dims % 2 == 1 else 4
    return (dims ** 2 - own) // 2 + own


def fft2d_r2c(a, threads=1):
    return _sfft.rfftn(np.asarray(a, dtype=np.float32), axes=(0, 1), workers=threads).astype(np.complex64, copy=False)


def ifft3d_c2r(ak, dims, threads=1):
    """IFFT3Dr_f (Pk_library.pyx:149-163): normalised inverse, float32."""
    return _sfft.irfftn(ak, s=(dims, dims, dims), axes=(0, 1, 2), workers=threads).astype(np.float32, copy=False)


class Pk_plane:
    """Pk_library.pyx:435-511."""

    def __init__(self, delta, BoxSize, MAS="CIC", threads=1, verbose=True, delta_k=None):
        grid = len(delta) if delta_k is None else delta_k.shape[0]
        kF, kN, kmax_par, kmax_per, kmax = frequencies_2D(BoxSize, grid)
        dk = np.ascontiguousarray(fft2d_r2c(delta, threads) if delta_k is None else delta_k)
        k2D, Nm, Pk2D = _z(kmax + 1), _z(kmax + 1), _z(kmax + 1)
        mi = np.array([_c.MAS_function(MAS)], dtype=np.int32)
        _lib().oracle_plane_bin(_cf(dk), grid, 1, mi.ctypes.data_as(_ipp), _c._dp(k2D), _c._dp(Nm), _c._dp(Pk2D), None)
        if int(np.sum(Nm)) != expected_modes_2D(grid):
            raise RuntimeError("WARNING: Not all modes counted")
        k2D, Nm, Pk2D = k2D[1:], Nm[1:], Pk2D[1:]
        for i in range(len(k2D)):                                   # :505-507
            k2D[i] = (k2D[i] / Nm[i]) * kF
            Pk2D[i] = (Pk2D[i] / Nm[i]) * (BoxSize / grid ** 2) ** 2
        self.k, self.Nmodes, self.Pk = k2D, Nm, Pk2D


class XPk_plane:
    """Pk_library.pyx:1093-1220."""

    def __init__(self, delta1, delta2, BoxSize, MAS1=None, MAS2=None, threads=1):
        grid = delta1.shape[0]
        if delta1.shape[0] != delta2.shape[1]:
            raise Exception("Images have different grid sizes!!!")
        kF, kN, kmax_par, kmax_per, kmax = frequencies_2D(BoxSize, grid)
        dk = np.ascontiguousarray(np.stack([fft2d_r2c(delta1, threads), fft2d_r2c(delta2, threads)]))
        k2D, Nm, Pk2D, PkX = _z(kmax + 1), _z(kmax + 1), _z(kmax + 1, 2), _z(kmax + 1)
        mi = np.array([_c.MAS_function(MAS1), _c.MAS_function(MAS2)], dtype=np.int32)
        _lib().oracle_plane_bin(_cf(dk), grid, 2, mi.ctypes.data_as(_ipp), _c._dp(k2D), _c._dp(Nm), _c._dp(Pk2D),
                                _c._dp(PkX))
        fact = (BoxSize / grid ** 2) ** 2
        k2D, Nm, Pk2D, PkX = k2D[1:], Nm[1:], Pk2D[1:], PkX[1:]
        for i in range(len(k2D)):                                   # :1210-1214
            k2D[i] = (k2D[i] / Nm[i]) * kF
            for j in range(2):
                Pk2D[i, j] = (Pk2D[i, j] / Nm[i]) * fact
            PkX[i] = (PkX[i] / Nm[i]) * fact
        self.k, self.Nmodes, self.Pk, self.XPk = k2D, Nm, Pk2D, PkX
        self.r = self.XPk / np.sqrt(self.Pk[:, 0] * self.Pk[:, 1])


class XPk_imag(_c.XPk):
    """Pk_library.pyx:811-1077: XPk with the cross term imag_i*real_j - real_i*imag_j."""

    def __init__(self, delta, BoxSize, axis=2, MAS=None, threads=1, delta_k=None):
        src = delta if delta_k is None else delta_k
        dims = len(src[0]); F = len(src); X = F * (F - 1) // 2
        kF, kN, kmax_par, kmax_per, kmax = _c.frequencies(BoxSize, dims)
        dks = [_c.fft3d_r2c(d, threads) for d in delta] if delta_k is None else list(delta_k)
        dk = np.ascontiguousarray(np.stack(dks).astype(np.complex64, copy=False))
        n2 = (kmax_par + 1) * (kmax_per + 1)
        r = dict(k3D=_z(kmax + 1), Nm3D=_z(kmax + 1), Pk3D=_z(kmax + 1, 3, F), PkX3D=_z(kmax + 1, 3, max(X, 1)),
                 k1D=_z(kmax_par + 1), Nm1D=_z(kmax_par + 1), Pk1D=_z(kmax_par + 1, F),
                 PkX1D=_z(kmax_par + 1, max(X, 1)), Nm2D=_z(n2), Pk2D=_z(n2, F), PkX2D=_z(n2, max(X, 1)))
        mi = np.asarray([_c.MAS_function(m) for m in MAS], dtype=np.int32)
        _lib().oracle_pk_bin_imag(_cf(dk), dims, F, mi.ctypes.data_as(_ipp), axis, kmax_par, kmax_per, kmax,
                                  *[_c._dp(r[n]) for n in ("k3D", "Nm3D", "Pk3D", "PkX3D", "k1D", "Nm1D", "Pk1D",
                                                           "PkX1D", "Nm2D", "Pk2D", "PkX2D")])
        self._finish(r, F, X, BoxSize, dims)

    def _finish(self, r, F, X, BoxSize, dims):
        """Finalisation :1019-1075 (same expressions as XPk's)."""
        kF, kN, kmax_par, kmax_per, kmax = _c.frequencies(BoxSize, dims)
        fact = (BoxSize / dims ** 2) ** 3
        Nm1 = r["Nm1D"][1:]
        k1D = (r["k1D"][1:] / Nm1) * kF
        kmaxper = np.sqrt(kN ** 2 - k1D ** 2)
        w1 = (np.pi * kmaxper ** 2 / Nm1)[:, None]
        self.k1D, self.Nmodes1D = k1D, Nm1
        self.Pk1D = (r["Pk1D"][1:] * fact) * w1 / (2.0 * np.pi) ** 2
        self.PkX1D = (r["PkX1D"][1:, :X] * fact) * w1 / (2.0 * np.pi) ** 2
        n2 = (kmax_par + 1) * (kmax_per + 1)
        i2 = np.arange(n2)
        self.kpar = 0.5 * (2 * (i2 % (kmax_par + 1)) + 1) * kF
        self.kper = 0.5 * (2 * (i2 // (kmax_par + 1)) + 1) * kF
        with np.errstate(invalid="ignore", divide="ignore"):
            self.Pk2D = r["Pk2D"] * fact / r["Nm2D"][:, None]
            self.PkX2D = r["PkX2D"][:, :X] * fact / r["Nm2D"][:, None]
        self.Nmodes2D = r["Nm2D"]
        if int(np.sum(r["Nm3D"])) != _c.expected_modes(dims):
            raise RuntimeError("WARNING: Not all modes counted")
        Nm3 = r["Nm3D"][1:]
        ell = np.array([1.0, 5.0, 9.0])[None, :, None]
        self.k3D, self.Nmodes3D = (r["k3D"][1:] / Nm3) * kF, Nm3
        self.Pk = (r["Pk3D"][1:] * ell / Nm3[:, None, None]) * fact
        self.XPk = (r["PkX3D"][1:, :, :X] * ell / Nm3[:, None, None]) * fact


def XPk_2D(delta1, delta2, BoxSize, axis=2, MAS1="CIC", MAS2="CIC", threads=1):
    """Pk_library.pyx:1754-1865: [kpar, kper, Pk1, Pk2, PkX, Nmodes] of the 2D (k_par, k_per) bins."""
    dims = len(delta1)
    if dims != len(delta2):
        raise ValueError("Different grids in the two fields!!!")
    kF = 2.0 * np.pi / BoxSize
    r = _c.bin_raw([_c.fft3d_r2c(delta1, threads), _c.fft3d_r2c(delta2, threads)], dims,
                   [_c.MAS_function(MAS1), _c.MAS_function(MAS2)], axis, BoxSize)
    middle = dims // 2
    imax_par, imax_per = middle, int(np.sqrt(middle ** 2 + middle ** 2))
    n2 = (imax_par + 1) * (imax_per + 1)
    kpar, kper = np.zeros(n2), np.zeros(n2)
    for ipar in range(imax_par + 1):
        for iper in range(imax_per + 1):
            index = (imax_par + 1) * iper + ipar
            kpar[index] = 0.5 * (ipar + ipar + 1) * kF
            kper[index] = 0.5 * (iper + iper + 1) * kF
    Nm = r["Nm2D"]
    with np.errstate(invalid="ignore", divide="ignore"):
        Pk1 = r["Pk2D"][:, 0] * (BoxSize / dims ** 2) ** 3 / Nm       # :1860-1862
        Pk2 = r["Pk2D"][:, 1] * (BoxSize / dims ** 2) ** 3 / Nm
        PkX = r["PkX2D"][:, 0] * (BoxSize / dims ** 2) ** 3 / Nm
    return [kpar, kper, Pk1, Pk2, PkX, Nm]


def _vel(kind, fields, BoxSize, MAS, threads):
    dims = len(fields[0])
    kF, kN, kmax_par, kmax_per, kmax = _c.frequencies(BoxSize, dims)
    dk = np.ascontiguousarray(np.stack([_c.fft3d_r2c(f, threads) for f in fields]))
    k, Nm, P1, P2, PX = (_z(kmax + 1) for _ in range(5))
    _lib().oracle_vel_bin(kind, _cf(dk), dims, _c.MAS_function(MAS), *[_c._dp(a) for a in (k, Nm, P1, P2, PX)])
    if int(np.sum(Nm)) != _c.expected_modes(dims):
        raise RuntimeError("WARNING: Not all modes counted")
    return dims, kF, k, Nm, P1, P2, PX


def Pk_theta(Vx, Vy, Vz, BoxSize, axis=2, MAS="CIC", threads=1):
    """Pk_library.pyx:1238-1326."""
    dims, kF, k, Nm, P1, _, _ = _vel(0, [Vx, Vy, Vz], BoxSize, MAS, threads)
    k = k[1:]; Nm = Nm[1:]; k = (k / Nm) * kF
    Pk = P1[1:] * (BoxSize / dims ** 2) ** 3 * kF ** 2; Pk *= (1.0 / Nm)            # :1322-1323
    return k, Pk, Nm


def XPk_dv(delta, Vx, Vy, Vz, BoxSize, axis=2, MAS="CIC", threads=1):
    """Pk_library.pyx:1345-1444.  Like the reference, Vx, Vy, Vz are multiplied by (1+delta) IN PLACE."""
    Vx *= (1.0 + delta); Vy *= (1.0 + delta); Vz *= (1.0 + delta)
    dims, kF, k, Nm, P1, P2, PX = _vel(1, [delta, Vx, Vy, Vz], BoxSize, MAS, threads)
    k = k[1:]; Nm = Nm[1:]; k = (k / Nm) * kF
    Pk1 = P1[1:] * (BoxSize / dims ** 2) ** 3; Pk1 *= (1.0 / Nm)
    Pk2 = P2[1:] * (BoxSize / dims ** 2) ** 3 * kF ** 2; Pk2 *= (1.0 / Nm)
    PkX = PX[1:] * (BoxSize / dims ** 2) ** 3 * kF; PkX *= (1.0 / Nm)
    return k, Pk1, Pk2, PkX, Nm


def XPk_vv(delta1, Vx1, Vy1, Vz1, delta2, Vx2, Vy2, Vz2, BoxSize, axis=2, MAS="CIC", threads=1):
    """Pk_library.pyx:1467-1580 (velocities multiplied by (1+delta) in place)."""
    Vx1 *= (1.0 + delta1); Vy1 *= (1.0 + delta1); Vz1 *= (1.0 + delta1)
    Vx2 *= (1.0 + delta2); Vy2 *= (1.0 + delta2); Vz2 *= (1.0 + delta2)
    dims, kF, k, Nm, P1, P2, PX = _vel(2, [delta1, Vx1, Vy1, Vz1, delta2, Vx2, Vy2, Vz2], BoxSize, MAS, threads)
    k = k[1:]; Nm = Nm[1:]; k = (k / Nm) * kF
    out = []
    for P in (P1, P2, PX):
        Q = P[1:] * (BoxSize / dims ** 2) ** 3 * kF ** 2; Q *= (1.0 / Nm)
        out.append(Q)
    return k, out[0], out[1], out[2], Nm


def correct_MAS(delta, BoxSize, MAS="CIC", threads=1):
    """Pk_library.pyx:1882-1939: FFT, deconvolve the independent modes, normalised inverse FFT."""
    dims = len(delta)
    dk = np.ascontiguousarray(_c.fft3d_r2c(delta, threads))
    _lib().oracle_correct_mas_modes(_cf(dk), dims, _c.MAS_function(MAS))
    return ifft3d_c2r(dk, dims, threads)


def expected_table(k_in, Pk_in, bins):
    """Pk_library.pyx:1971-1993: sortedness check and the log-spaced float32 interpolation table."""
    k_in = np.asarray(k_in, dtype=np.float32); Pk_in = np.asarray(Pk_in, dtype=np.float32)
    if np.any(k_in[1:] <= k_in[:-1]):
        raise Exception("Input k-array not sorted!!!")
    f32 = np.float32
    kmin_in, kmax_in = k_in[0], k_in[-1]
    # `cdef float deltak = (log10(kmax_in) - log10(kmin_in))/(bins-1.0)`: double arithmetic, rounded on store
    deltak = f32((math.log10(float(kmax_in)) - math.log10(float(kmin_in))) / (bins - 1.0))
    tk, tP = np.zeros(bins, f32), np.zeros(bins, f32)
    j = 1
    for i in range(bins):
        # `deltak*i` is a C float*int product, i.e. float32; the sum and the power are double (:1988)
        tk[i] = f32(10.0 ** (math.log10(float(kmin_in)) + float(f32(deltak) * f32(i))))
        while tk[i] > k_in[j] and j < len(k_in) - 1:     # the reference would read past the end here (:1990)
            j += 1
        tP[i] = (Pk_in[j] - Pk_in[j - 1]) / (k_in[j] - k_in[j - 1]) * (tk[i] - k_in[j - 1]) + Pk_in[j - 1]
    return tk, tP, kmin_in, deltak


def expected_Pk(k_in, Pk_in, BoxSize, dims, bins=750):
    """Pk_library.pyx:1956-2047."""
    BoxSize = float(np.float32(BoxSize))
    kF, kN, kmax_par, kmax_per, kmax = _c.frequencies(BoxSize, dims)
    k_in = np.asarray(k_in, dtype=np.float32)
    if kF < k_in[0] or kmax * kF > k_in[-1]:
        raise Exception("k value in grid outside input k range")
    tk, tP, kmin_in, deltak = expected_table(k_in, Pk_in, bins)
    k3D, Pk3D, Nm = _z(kmax + 1), _z(kmax + 1), _z(kmax + 1)
    _lib().oracle_expected_pk(dims, np.float32(kF), tk.ctypes.data_as(_fpp), tP.ctypes.data_as(_fpp), kmin_in, deltak,
                              _c._dp(k3D), _c._dp(Pk3D), _c._dp(Nm))
    return k3D[1:] / Nm[1:], Pk3D[1:] / Nm[1:], Nm[1:]


def _xi_finish(self, xi_grid, dims, BoxSize, axis):
    kmax = _c.frequencies(BoxSize, dims)[4]
    r3D, xi3D, Nm = _z(kmax + 1), _z(kmax + 1, 3), _z(kmax + 1)
    xi_grid = np.ascontiguousarray(xi_grid, dtype=np.float32)
    _lib().oracle_xi_bin(xi_grid.ctypes.data_as(_fpp), dims, axis, _c._dp(r3D), _c._dp(xi3D), _c._dp(Nm))
    r3D, Nm, xi3D = r3D[1:], Nm[1:], xi3D[1:]
    for i in range(len(r3D)):                                       # :2274-2278
        r3D[i] = (r3D[i] / Nm[i]) * (BoxSize * 1.0 / dims)
        xi3D[i, 0] = (xi3D[i, 0] / Nm[i]) * (1.0 / dims ** 3)
        xi3D[i, 1] = (xi3D[i, 1] * 5.0 / Nm[i]) * (1.0 / dims ** 3)
        xi3D[i, 2] = (xi3D[i, 2] * 9.0 / Nm[i]) * (1.0 / dims ** 3)
    self.r3D, self.Nmodes3D, self.xi = r3D, Nm, xi3D


class Xi:
    """Pk_library.pyx:2168-2282 (`BoxSize` is a C float there)."""

    def __init__(self, delta, BoxSize, MAS="CIC", axis=2, threads=1):
        BoxSize = float(np.float32(BoxSize))
        dims = delta.shape[0]
        dk = np.ascontiguousarray(_c.fft3d_r2c(delta, threads))
        _lib().oracle_xi_modes(_cf(dk), None, dims, _c.MAS_function(MAS), 0)
        _xi_finish(self, ifft3d_c2r(dk, dims, threads), dims, BoxSize, axis)


class XXi:
    """Pk_library.pyx:2298-2427."""

    def __init__(self, delta1, delta2, BoxSize, MAS=["CIC", "CIC"], axis=2, threads=1):
        BoxSize = float(np.float32(BoxSize))
        grid = delta1.shape[0]
        if grid != delta2.shape[0]:
            raise Exception("grid sizes differ!!!")
        d1 = np.ascontiguousarray(_c.fft3d_r2c(delta1, threads))
        d2 = np.ascontiguousarray(_c.fft3d_r2c(delta2, threads))
        _lib().oracle_xi_modes(_cf(d1), _cf(d2), grid, _c.MAS_function(MAS[0]), _c.MAS_function(MAS[1]))
        _xi_finish(self, ifft3d_c2r(d1, grid, threads), grid, BoxSize, axis)


class XXi_projected:
    """Pk_library.pyx:2684-2789: projected cross-correlation function of two images.  Every stored mode of the 2D
    half-spectrum is deconvolved (no duplicate-mode rule here, :2735-2758), the product re1*re2 + im1*im2 is formed
    in float32, the inverse transform is the normalised 2D c2r and every cell of the (grid, grid) result is binned
    by int(sqrt(kx^2 + ky^2)) (:2776-2787).  Attributes: r_p, xi_p, Nmodes_p (bin 0 dropped)."""

    def __init__(self, delta1, delta2, BoxSize, MAS=["CIC", "CIC"], threads=1):
        BoxSize = float(np.float32(BoxSize))
        grid = delta1.shape[0]
        middle = grid // 2
        if grid != delta2.shape[0]:
            raise Exception("grid sizes differ!!!")
        kmax = int((grid // 2) * np.sqrt(2))
        i1, i2 = _c.MAS_function(MAS[0]), _c.MAS_function(MAS[1])
        d1, d2 = fft2d_r2c(delta1, threads), fft2d_r2c(delta2, threads)
        prefact = np.pi / grid
        kx = _signed(grid)[:, None]
        ky = np.arange(middle + 1)[None, :]

        def corr(k, p):
            x = prefact * k
            with np.errstate(invalid="ignore", divide="ignore"):
                return np.where(k == 0, 1.0, (x / np.sin(x)) ** p)
        f1 = (corr(kx, i1) * corr(ky, i1)).astype(np.float32)          # cdef float MAS_factor (:2697)
        f2 = (corr(kx, i2) * corr(ky, i2)).astype(np.float32)
        a = (d1 * f1).astype(np.complex64)
        b = (d2 * f2).astype(np.complex64)
        prod = (a.real * b.real + a.imag * b.imag).astype(np.float32)  # float32 products and sum (:2752-2757)
        xi_grid = _sfft.irfft2(prod.astype(np.complex64), s=(grid, grid)).astype(np.float32)
        kxf = _signed(grid)[:, None]
        kyf = _signed(grid)[None, :]
        k = np.sqrt((kxf * kxf + kyf * kyf).astype(np.float64))
        idx = k.astype(np.int64).ravel()
        r_p = np.bincount(idx, weights=k.ravel(), minlength=kmax + 1)[1:]
        xi_p = np.bincount(idx, weights=xi_grid.astype(np.float64).ravel(), minlength=kmax + 1)[1:]
        Nm = np.bincount(idx, minlength=kmax + 1).astype(np.float64)[1:]
        with np.errstate(invalid="ignore", divide="ignore"):
            self.r_p = (r_p / Nm) * (BoxSize * 1.0 / grid)
            self.xi_p = (xi_p / Nm) * (1.0 / grid ** 2)
        self.Nmodes_p = Nm


def field_smoothing(field, filter_k, threads=1):
    """smoothing_library.pyx:215-235: IFFT(FFT(field) * filter_k), complex64 product."""
    dims = field.shape[0]
    if field.shape[0] != filter_k.shape[0]:
        raise Exception("field and filter have different grids!!!")
    fk = _c.fft3d_r2c(field, threads)
    fk = (fk * np.asarray(filter_k, dtype=np.complex64)).astype(np.complex64)
    return ifft3d_c2r(fk, dims, threads)


# ---- smoothing_library: filters (smoothing_library.pyx:20-120, 123-209) and the 2D smoothing (:243-262) -------
def _signed(n):
    i = np.arange(n)
    return np.where(i > n // 2, i - n, i)


def _ft_filter(BoxSize, R, dims, Filter, kmin, kmax, nd):
    """Common body of FT_filter / FT_filter_2D.  C-typed locals of the reference: BoxSize, R, kmin, kmax,
    R_grid, R2, kF, k, factor are `float`; d2 is `int`; normalization is `double`."""
    if Filter not in ["Top-Hat", "Gaussian", "Top-Hat-k"]:
        raise Exception("Filter %s not implemented!" % Filter)
    f32 = np.float32
    BoxSize, R, kmin, kmax = f32(BoxSize), f32(R), f32(kmin), f32(kmax)
    middle = dims // 2
    R_grid = f32(f32(R * f32(dims)) / BoxSize)                  # (R*dims/BoxSize), float arithmetic
    R2 = f32(R_grid * R_grid)
    kF = f32(2.0 * np.pi / float(BoxSize))
    s = _signed(dims)
    if nd == 3:
        d2 = s[:, None, None] ** 2 + s[None, :, None] ** 2 + s[None, None, :] ** 2
        d2h = s[:, None, None] ** 2 + s[None, :, None] ** 2 + np.arange(middle + 1)[None, None, :] ** 2
    else:
        d2 = s[:, None] ** 2 + s[None, :] ** 2
        d2h = s[:, None] ** 2 + np.arange(middle + 1)[None, :] ** 2
    if Filter == "Top-Hat":
        field = np.where(d2.astype(f32) <= R2, f32(1.0), f32(0.0)).astype(f32)          # `d2<=R2`: int vs float
        normalization = float(np.sum(field, dtype=np.float64))
    elif Filter == "Gaussian":
        field = np.exp(-d2.astype(np.float64) / (2.0 * float(R2))).astype(f32)          # double exp, float store
        normalization = float(np.sum(field, dtype=np.float64))
    else:
        k = (float(kF) * np.sqrt(d2h.astype(np.float64))).astype(f32)                   # `cdef float k`
        field_k = np.where((k >= kmin) & (k < kmax), 1.0, 0.0).astype(np.complex64)
        field_k[(0,) * nd] = 1.0                                                        # "dont mess with DC mode"
        field = _sfft.irfftn(field_k, s=(dims,) * nd, axes=tuple(range(nd))).astype(f32)
        normalization = float(np.sum(field, dtype=np.float64))
    field = (field.astype(np.float64) / normalization).astype(f32)                      # float / double
    return _sfft.rfftn(field, axes=tuple(range(nd))).astype(np.complex64)


def FT_filter(BoxSize, R, dims, Filter, threads=1, kmin=0, kmax=0):
    """smoothing_library.pyx:20-120."""
    return _ft_filter(BoxSize, R, dims, Filter, kmin, kmax, 3)


def FT_filter_2D(BoxSize, R, grid, Filter, threads=1, kmin=0, kmax=0):
    """smoothing_library.pyx:123-209."""
    return _ft_filter(BoxSize, R, grid, Filter, kmin, kmax, 2)


def field_smoothing_2D(field, filter_k, threads=1):
    """smoothing_library.pyx:243-262."""
    grid = field.shape[0]
    if field.shape[0] != filter_k.shape[0]:
        raise Exception("field and filter have different grids!!!")
    fk = fft2d_r2c(field, threads)
    fk = (fk * np.asarray(filter_k, dtype=np.complex64)).astype(np.complex64)
    return _sfft.irfftn(fk, s=(grid, grid), axes=(0, 1), workers=threads).astype(np.float32)


class XXi_multi:
    """Pk_library.pyx:2443-2670: auto- and cross-correlation function multipoles of several fields.
    As in the reference, a cross term (i, j) is deconvolved with field i's window for BOTH fields (:2548-2553)."""

    def __init__(self, delta, BoxSize, axis=2, MAS=None, threads=1):
        dims = len(delta[0]); F = len(delta)
        for d in delta[1:]:
            if len(d) != dims:
                raise ValueError("Fields have different grid sizes!!!")
        kmax = _c.frequencies(BoxSize, dims)[4]
        dk = [np.ascontiguousarray(_c.fft3d_r2c(d, threads)) for d in delta]
        mi = [_c.MAS_function(m) for m in MAS]
        grids = []
        for i in range(F):
            a = dk[i].copy()
            _lib().oracle_xi_modes(_cf(a), None, dims, mi[i], 0)
            grids.append(ifft3d_c2r(a, dims, threads))
        xgrids = []
        for i in range(F):
            for j in range(i + 1, F):
                a = dk[i].copy()
                _lib().oracle_xi_modes(_cf(a), _cf(dk[j]), dims, mi[i], mi[i])
                xgrids.append(ifft3d_c2r(a, dims, threads))
        X = len(xgrids)
        xi = np.zeros((kmax, 3, F)); Xxi = np.zeros((kmax, 3, X))

        class _R:
            pass
        for n, (src, dst) in enumerate([(g, xi) for g in grids] + [(g, Xxi) for g in xgrids]):
            r = _R()
            _xi_finish(r, src, dims, BoxSize, axis)
            dst[:, :, n if n < F else n - F] = r.xi
            self.r3D, self.Nmodes3D = r.r3D, r.Nmodes3D
        self.xi, self.Xxi = xi, Xxi
