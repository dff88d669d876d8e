"""Python face of the CPU oracle (TEST INFRASTRUCTURE ONLY): MA / Pk / XPk restated.

`MA`, `Pk`, `XPk` here take the same arguments and return the same attributes as the
reference's MAS_library.MA (library/MAS_library/MAS_library.pyx:57-112) and
Pk_library.Pk / XPk (library/Pk_library/Pk_library.pyx:263-420, 529-793).  The hot loops
live in ma_oracle.c / pk_oracle.c; the finalisation below restates Pk_library.pyx:384-418
and :735-791 line by line.  Pinned against tests/golden (outputs of the compiled reference).
"""
import ctypes
import os

import numpy as np
import scipy.fft as _sfft

from . import build as _build

_lib = None
_MAS = {"NGP": 0, "CIC": 1, "TSC": 2, "PCS": 3}
_NREP = {"NGP": 1.0, "CIC": 2.0, "TSC": 3.0, "PCS": 4.0}


def lib():
    global _lib
    if _lib is None:
        so = _build.SO
        if not os.path.exists(so):
            so = _build.build()
        L = ctypes.CDLL(so)
        fp = ctypes.POINTER(ctypes.c_float)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        L.oracle_ma.argtypes = [ctypes.c_int, fp, fp, fp, ctypes.c_long, ctypes.c_int,
                                ctypes.c_int, ctypes.c_float]
        L.oracle_ma.restype = None
        L.oracle_cic_interp.argtypes = [fp, ctypes.c_int, ctypes.c_float, fp, ctypes.c_long, fp]
        L.oracle_cic_interp.restype = None
        L.oracle_pos_redshift_space.argtypes = [fp, fp, ctypes.c_long, ctypes.c_float, ctypes.c_float,
                                                ctypes.c_float, ctypes.c_int]
        L.oracle_pos_redshift_space.restype = None
        L.oracle_pk_bin.argtypes = [fp, ctypes.c_int, ctypes.c_int, ip, ctypes.c_int,
                                    ctypes.c_int, ctypes.c_int, ctypes.c_int] + [dp] * 12
        L.oracle_pk_bin.restype = None
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _dp(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def MA(pos, number, BoxSize, MAS="CIC", W=None, verbose=False, renormalize_2D=True):
    """MAS_library.pyx:57-112.  `number` (float32, C-contiguous) is accumulated in place."""
    coord = pos.shape[1]
    if coord != number.ndim:
        raise ValueError("pos have %d dimensions and the density %d!!!" % (coord, number.ndim))
    if MAS not in _MAS:
        raise ValueError("option not valid!!!")
    assert number.dtype == np.float32 and number.flags.c_contiguous
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    Wc = None if W is None else np.ascontiguousarray(W, dtype=np.float32)
    lib().oracle_ma(_MAS[MAS], _fp(pos), _fp(number), None if Wc is None else _fp(Wc),
                    pos.shape[0], number.shape[0], coord, np.float32(BoxSize))
    if coord == 2 and renormalize_2D and MAS != "NGP":
        number /= _NREP[MAS]          # :90-107 -- divides the whole accumulated plane


def CIC_interp(density, BoxSize, pos, den):
    """MAS_library.pyx:558-599: den[i] = CIC-interpolated value of the 3D grid at pos[i] (overwritten)."""
    assert density.dtype == np.float32 and density.flags.c_contiguous and density.ndim == 3
    assert den.dtype == np.float32 and den.flags.c_contiguous
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    lib().oracle_cic_interp(_fp(density), density.shape[0], np.float32(BoxSize), _fp(pos), pos.shape[0], _fp(den))


def pos_redshift_space(pos, vel, BoxSize, Hubble, redshift, axis):
    """redshift_space_library.pyx:29-46: pos[:,axis] += (1+z)/H * vel[:,axis], wrapped into the box, IN PLACE."""
    assert pos.dtype == np.float32 and pos.flags.c_contiguous and pos.shape[1] == 3
    assert vel.dtype == np.float32 and vel.flags.c_contiguous and vel.shape == pos.shape
    lib().oracle_pos_redshift_space(_fp(pos), _fp(vel), pos.shape[0], np.float32(BoxSize), np.float32(Hubble),
                                    np.float32(redshift), int(axis))


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
    """Pk_library.pyx:72-78."""
    return {"NGP": 1, "CIC": 2, "TSC": 3, "PCS": 4}.get(MAS, 0)


def expected_modes(dims):
    """Pk_library.pyx:87-93."""
    own = 1 if dims % 2 == 1 else 8
    return (dims ** 3 - own) // 2 + own


def fft3d_r2c(delta, threads=1):
    """The transform the pyfftw shim gives the compiled reference (Pk_library.pyx:117-130)."""
    return _sfft.rfftn(np.asarray(delta, dtype=np.float32), axes=(0, 1, 2), workers=threads) \
        .astype(np.complex64, copy=False)


def bin_raw(delta_k_list, dims, mas_indices, axis, BoxSize, phase=False):
    """Run the C loop on already-transformed fields; returns the RAW accumulators."""
    F = len(delta_k_list)
    X = F * (F - 1) // 2
    kF, kN, kmax_par, kmax_per, kmax = frequencies(BoxSize, dims)
    dk = np.ascontiguousarray(np.stack(delta_k_list).astype(np.complex64, copy=False))
    n2 = (kmax_par + 1) * (kmax_per + 1)
    z = lambda *s: np.zeros(s, dtype=np.float64)
    r = dict(k3D=z(kmax + 1), Nm3D=z(kmax + 1), Pk3D=z(kmax + 1, 3, F), PkX3D=z(kmax + 1, 3, max(X, 1)),
             phase=z(kmax + 1) if phase else None,
             k1D=z(kmax_par + 1), Nm1D=z(kmax_par + 1), Pk1D=z(kmax_par + 1, F),
             PkX1D=z(kmax_par + 1, max(X, 1)), Nm2D=z(n2), Pk2D=z(n2, F), PkX2D=z(n2, max(X, 1)))
    mi = np.asarray(mas_indices, dtype=np.int32)
    lib().oracle_pk_bin(dk.view(np.float32).ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                        dims, F, mi.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), axis,
                        kmax_par, kmax_per, kmax,
                        _dp(r["k3D"]), _dp(r["Nm3D"]), _dp(r["Pk3D"]), _dp(r["PkX3D"]), _dp(r["phase"]),
                        _dp(r["k1D"]), _dp(r["Nm1D"]), _dp(r["Pk1D"]), _dp(r["PkX1D"]),
                        _dp(r["Nm2D"]), _dp(r["Pk2D"]), _dp(r["PkX2D"]))
    if X == 0:
        r["PkX3D"] = r["PkX3D"][:, :, :0]; r["PkX1D"] = r["PkX1D"][:, :0]; r["PkX2D"] = r["PkX2D"][:, :0]
    return r


class Pk:
    """Pk_library.pyx:263-420."""

    def __init__(self, delta, BoxSize, axis=2, MAS="CIC", threads=1, verbose=True, delta_k=None):
        """`delta_k` (test hook): bin this half-spectrum instead of transforming `delta`."""
        dims = len(delta) if delta_k is None else delta_k.shape[0]
        kF, kN, kmax_par, kmax_per, kmax = frequencies(BoxSize, dims)
        dk = fft3d_r2c(delta, threads) if delta_k is None else delta_k
        r = bin_raw([dk], dims, [MAS_function(MAS)], axis, BoxSize, phase=True)
        fact = (BoxSize / dims ** 2) ** 3

        # 1D (:384-391): drop the DC bin, units, <k_par>, perpendicular-area weight
        k1D = r["k1D"][1:].copy(); Nm1D = r["Nm1D"][1:].copy(); Pk1D = r["Pk1D"][1:, 0].copy()
        for i in range(len(k1D)):
            Pk1D[i] = Pk1D[i] * fact
            k1D[i] = (k1D[i] / Nm1D[i]) * kF
            kmaxper = np.sqrt(kN ** 2 - k1D[i] ** 2)
            Pk1D[i] = Pk1D[i] * (np.pi * kmaxper ** 2 / Nm1D[i]) / (2.0 * np.pi) ** 2
        self.k1D, self.Pk1D, self.Nmodes1D = k1D, Pk1D, Nm1D

        # 2D (:394-404): DC bin kept
        n2 = (kmax_par + 1) * (kmax_per + 1)
        kpar = np.zeros(n2); kper = np.zeros(n2)
        for k_par in range(kmax_par + 1):
            for k_per in range(kmax_per + 1):
                i2 = (kmax_par + 1) * k_per + k_par
                kpar[i2] = 0.5 * (k_par + k_par + 1) * kF
                kper[i2] = 0.5 * (k_per + k_per + 1) * kF
        with np.errstate(invalid="ignore", divide="ignore"):
            Pk2D = r["Pk2D"][:, 0] * fact / r["Nm2D"]
        self.kpar, self.kper, self.Pk2D, self.Nmodes2D = kpar, kper, Pk2D, r["Nm2D"]

        # 3D (:408-418)
        if int(np.sum(r["Nm3D"])) != expected_modes(dims):
            raise RuntimeError("WARNING: Not all modes counted")
        k3D = r["k3D"][1:].copy(); Nm3D = r["Nm3D"][1:].copy()
        Pk3D = r["Pk3D"][1:, :, 0].copy(); ph = r["phase"][1:].copy()
        for i in range(len(k3D)):
            k3D[i] = (k3D[i] / Nm3D[i]) * kF
            Pk3D[i, 0] = (Pk3D[i, 0] / Nm3D[i]) * fact
            Pk3D[i, 1] = (Pk3D[i, 1] * 5.0 / Nm3D[i]) * fact
            Pk3D[i, 2] = (Pk3D[i, 2] * 9.0 / Nm3D[i]) * fact
            ph[i] = (ph[i] / Nm3D[i]) * fact
        self.k3D, self.Nmodes3D, self.Pk, self.Pkphase = k3D, Nm3D, Pk3D, ph


class XPk:
    """Pk_library.pyx:529-793."""

    def __init__(self, delta, BoxSize, axis=2, MAS=None, threads=1, delta_k=None):
        """`delta_k` (test hook): list of half-spectra to bin instead of transforming `delta`."""
        src = delta if delta_k is None else delta_k
        dims = len(src[0]); F = len(src); X = F * (F - 1) // 2
        for d in src[1:]:
            if len(d) != dims:
                raise ValueError("Fields have different grid sizes!!!")
        kF, kN, kmax_par, kmax_per, kmax = frequencies(BoxSize, dims)
        dks = [fft3d_r2c(d, threads) for d in delta] if delta_k is None else list(delta_k)
        r = bin_raw(dks, dims, [MAS_function(m) for m in MAS], axis, BoxSize)
        fact = (BoxSize / dims ** 2) ** 3

        k1D = r["k1D"][1:].copy(); Nm1D = r["Nm1D"][1:].copy()
        Pk1D = r["Pk1D"][1:].copy(); PkX1D = r["PkX1D"][1:].copy()
        for i in range(len(k1D)):
            k1D[i] = (k1D[i] / Nm1D[i]) * kF
            kmaxper = np.sqrt(kN ** 2 - k1D[i] ** 2)
            for j in range(F):
                Pk1D[i, j] = Pk1D[i, j] * fact
                Pk1D[i, j] = Pk1D[i, j] * (np.pi * kmaxper ** 2 / Nm1D[i]) / (2.0 * np.pi) ** 2
            for j in range(X):
                PkX1D[i, j] = PkX1D[i, j] * fact
                PkX1D[i, j] = PkX1D[i, j] * (np.pi * kmaxper ** 2 / Nm1D[i]) / (2.0 * np.pi) ** 2
        self.k1D, self.Nmodes1D, self.Pk1D, self.PkX1D = k1D, Nm1D, Pk1D, PkX1D

        n2 = (kmax_par + 1) * (kmax_per + 1)
        kpar = np.zeros(n2); kper = np.zeros(n2)
        for k_par in range(kmax_par + 1):
            for k_per in range(kmax_per + 1):
                i2 = (kmax_par + 1) * k_per + k_par
                kpar[i2] = 0.5 * (k_par + k_par + 1) * kF
                kper[i2] = 0.5 * (k_per + k_per + 1) * kF
        with np.errstate(invalid="ignore", divide="ignore"):
            Pk2D = r["Pk2D"] * fact / r["Nm2D"][:, None]
            PkX2D = r["PkX2D"] * fact / r["Nm2D"][:, None]
        self.kpar, self.kper, self.Nmodes2D, self.Pk2D, self.PkX2D = kpar, kper, r["Nm2D"], Pk2D, PkX2D

        if int(np.sum(r["Nm3D"])) != expected_modes(dims):
            raise RuntimeError("WARNING: Not all modes counted")
        k3D = r["k3D"][1:].copy(); Nm3D = r["Nm3D"][1:].copy()
        Pk3D = r["Pk3D"][1:].copy(); PkX3D = r["PkX3D"][1:].copy()
        ell = np.array([1.0, 5.0, 9.0])
        for i in range(len(k3D)):
            k3D[i] = (k3D[i] / Nm3D[i]) * kF
            for l in range(3):
                Pk3D[i, l, :] = (Pk3D[i, l, :] * ell[l] / Nm3D[i]) * fact
                PkX3D[i, l, :] = (PkX3D[i, l, :] * ell[l] / Nm3D[i]) * fact
        self.k3D, self.Nmodes3D, self.Pk, self.XPk = k3D, Nm3D, Pk3D, PkX3D
