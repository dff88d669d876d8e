"""CPU: the per-thread bodies of the CUDA shell kernels (pylians3_b200/csrc/shell_body.cuh), run serially by
tests/harness/shell_host.cpp over the launch geometry of the device launcher, against the oracle's raw sums.

This is how the kernels' logic is checked in the GPU-less authoring container; the -m gpu tests
(test_gpu_pk_more.py) check the real launches."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

import more_cases as MC
from conftest import ROOT

HARNESS_DIR = os.path.join(ROOT, "tests", "harness")
SO = os.path.join(HARNESS_DIR, "libshell_host.so")
KIND = {"theta": 0, "dv": 1, "vv": 2, "expected": 3, "plane": 4, "xplane": 5, "xi": 6}
NV = {"theta": 1, "dv": 3, "vv": 3, "expected": 1, "plane": 1, "xplane": 3, "xi": 3}


@pytest.fixture(scope="module")
def harness():
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    src = os.path.join(HARNESS_DIR, "shell_host.cpp")
    hdr = os.path.join(ROOT, "pylians3_b200", "csrc", "shell_body.cuh")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                               "-I" + os.path.dirname(hdr), "-o", SO, src])
    L = ctypes.CDLL(SO)
    L.harness_shell.restype = ctypes.c_int
    L.harness_shell.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int), ctypes.c_int,
                                ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_float,
                                ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_void_p]
    L.harness_shell_bins.restype = ctypes.c_int
    L.harness_filter.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_float] * 4
    L.harness_modes.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    return L


def run_shell(L, kind, fields, mas, dims, axis=2, scale=1.0, table=None, sms=148):
    n3 = L.harness_shell_bins(KIND[kind], dims)
    out = np.zeros((2 + NV[kind]) * n3, dtype=np.uint64)
    fields = [np.ascontiguousarray(f) for f in fields]
    ptrs = (ctypes.c_void_p * max(len(fields), 1))(*[f.ctypes.data for f in fields])
    mi = (ctypes.c_int * 2)(*(list(mas) + [0, 0])[:2])
    tk = tP = None
    tn, kF, lk, dk = 0, 0.0, 0.0, 1.0
    if table is not None:
        tk, tP, kF, lk, dk = table
        tn = len(tk)
    nseg = L.harness_shell(KIND[kind], ptrs, mi, dims, axis, scale, None if tk is None else tk.ctypes.data,
                           None if tP is None else tP.ctypes.data, tn, kF, lk, dk, sms, out.ctypes.data)
    assert nseg >= 1
    f = out.view(np.float64)
    return dict(ksum=f[:n3].copy(), Nm=out[n3:2 * n3].astype(np.float64),
                vals=[f[(2 + j) * n3:(3 + j) * n3].copy() for j in range(NV[kind])], nseg=nseg)


def close(a, b, tol):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    den = np.maximum(np.abs(b), 1e-6 * max(np.abs(b).max(), 1e-300))
    return float(np.max(np.abs(a - b) / den)) < tol


def z(n):
    return np.zeros(n, np.float64)


# sms = 1 -> one segment (a thread walks the whole ky axis); 148 -> the B200 geometry; 4000 -> 8-step segments
@pytest.mark.parametrize("sms", [1, 148, 4000])
@pytest.mark.parametrize("N", [2, 3, 4, 5, 12, 9, 32])
def test_velocity_kinds(harness, oracle, N, sms):
    from oracle import cpu as C, cpu_more as M
    rng = np.random.default_rng(N)
    F = [rng.standard_normal((N, N, N)).astype(np.float32) for _ in range(7)]
    dks = [np.ascontiguousarray(C.fft3d_r2c(f)) for f in F]
    kmax = C.frequencies(1000.0, N)[4]
    for kind, idx, oidx, mas in (("theta", [1, 2, 3], [1, 2, 3], 2), ("dv", [0, 1, 2, 3], [0, 1, 2, 3], 3),
                                 ("vv", [1, 2, 3, 4, 5, 6], [0, 1, 2, 3, 0, 4, 5, 6], 4)):
        got = run_shell(harness, kind, [dks[i] for i in idx], [mas], N, sms=sms)
        stack = np.ascontiguousarray(np.stack([dks[i] for i in oidx]))
        k, Nm, P1, P2, PX = z(kmax + 1), z(kmax + 1), z(kmax + 1), z(kmax + 1), z(kmax + 1)
        M._lib().oracle_vel_bin(KIND[kind], M._cf(stack), N, mas, *[C._dp(a) for a in (k, Nm, P1, P2, PX)])
        assert np.array_equal(got["Nm"], Nm), kind
        assert int(Nm.sum()) == C.expected_modes(N)
        assert close(got["ksum"], k, 1e-12), kind
        assert close(got["vals"][0], P1, 2e-6), kind
        if kind != "theta":
            assert close(got["vals"][1], P2, 2e-6), kind
            # cross term: signed sum, float32 k.V round-off scales with the autos
            assert np.max(np.abs(got["vals"][2] - PX) / np.maximum(np.sqrt(P1 * P2), 1e-300)) < 1e-6, kind


@pytest.mark.parametrize("sms", [1, 148])
@pytest.mark.parametrize("N", [2, 3, 4, 5, 12, 9, 64])
def test_plane_kinds(harness, oracle, N, sms):
    from oracle import cpu as C, cpu_more as M
    rng = np.random.default_rng(100 + N)
    imgs = [np.ascontiguousarray(M.fft2d_r2c(rng.standard_normal((N, N)).astype(np.float32))) for _ in range(2)]
    kmax = M.frequencies_2D(1000.0, N)[4]
    for kind, F, mas in (("plane", 1, [2]), ("xplane", 2, [2, 4])):
        got = run_shell(harness, kind, imgs[:F], mas, N, sms=sms)
        k2D, Nm, Pk, PkX = z(kmax + 1), z(kmax + 1), np.zeros((kmax + 1, F)), z(kmax + 1)
        mi = np.array(mas, np.int32)
        M._lib().oracle_plane_bin(M._cf(np.ascontiguousarray(np.stack(imgs[:F]))), N, F, mi.ctypes.data_as(M._ipp),
                                  C._dp(k2D), C._dp(Nm), C._dp(Pk), C._dp(PkX) if F == 2 else None)
        assert np.array_equal(got["Nm"], Nm) and int(Nm.sum()) == M.expected_modes_2D(N)
        assert close(got["ksum"], k2D, 1e-12)
        assert close(got["vals"][0], Pk[:, 0], 1e-10)
        if F == 2:
            assert close(got["vals"][1], Pk[:, 1], 1e-10) and close(got["vals"][2], PkX, 1e-9)


@pytest.mark.parametrize("sms", [1, 148])
@pytest.mark.parametrize("N,axis", [(2, 0), (3, 1), (4, 2), (5, 0), (12, 0), (9, 1), (20, 2)])
def test_xi_kind(harness, oracle, N, axis, sms):
    from oracle import cpu as C, cpu_more as M
    rng = np.random.default_rng(200 + N)
    grid = rng.standard_normal((N, N, N)).astype(np.float32)
    scale = np.float32(1.0 / N ** 3)
    got = run_shell(harness, "xi", [grid], [], N, axis=axis, scale=float(scale), sms=sms)
    kmax = C.frequencies(1000.0, N)[4]
    r3D, xi3D, Nm = z(kmax + 1), np.zeros((kmax + 1, 3)), z(kmax + 1)
    scaled = (grid * scale).astype(np.float32)
    M._lib().oracle_xi_bin(scaled.ctypes.data_as(M._fpp), N, axis, C._dp(r3D), C._dp(xi3D), C._dp(Nm))
    assert np.array_equal(got["Nm"], Nm) and int(Nm.sum()) == N ** 3
    assert close(got["ksum"], r3D, 1e-12)
    for l in range(3):
        assert close(got["vals"][l], xi3D[:, l], 1e-10)


@pytest.mark.parametrize("N", [12, 9])
def test_expected_kind(harness, oracle, N):
    from oracle import cpu as C, cpu_more as M
    import math
    I = MC.inputs(N)
    tk, tP, kmin_in, deltak = M.expected_table(I["k_in"], I["Pk_in"], 300)
    kF = np.float32(2.0 * np.pi / 1000.0)
    got = run_shell(harness, "expected", [], [], N, table=(tk, tP, float(kF), math.log10(float(kmin_in)), float(deltak)))
    kmax = C.frequencies(1000.0, N)[4]
    k3D, Pk3D, Nm = z(kmax + 1), z(kmax + 1), z(kmax + 1)
    M._lib().oracle_expected_pk(N, kF, tk.ctypes.data_as(M._fpp), tP.ctypes.data_as(M._fpp), kmin_in, deltak,
                                C._dp(k3D), C._dp(Pk3D), C._dp(Nm))
    assert np.array_equal(got["Nm"], Nm) and Nm[0] == 0
    assert close(got["ksum"], k3D, 1e-12) and close(got["vals"][0], Pk3D, 1e-6)


@pytest.mark.parametrize("N", [12, 9])
def test_mode_passes(harness, oracle, N):
    from oracle import cpu as C, cpu_more as M
    rng = np.random.default_rng(300 + N)
    d1 = rng.standard_normal((N, N, N)).astype(np.float32)
    d2 = rng.standard_normal((N, N, N)).astype(np.float32)
    a, b = np.ascontiguousarray(C.fft3d_r2c(d1)), np.ascontiguousarray(C.fft3d_r2c(d2))
    # Xi / XXi mode loop: identical float32 expressions up to FMA contraction
    for second, masb in ((None, 0), (b, 4)):
        g, r = a.copy(), a.copy()
        harness.harness_modes(1, g.ctypes.data, None if second is None else second.ctypes.data, N, 2, masb if second is not None else 2)
        M._lib().oracle_xi_modes(M._cf(r), None if second is None else M._cf(second), N, 2, masb)
        assert np.max(np.abs(g - r)) < 2e-6 * np.abs(r).max()
    # correct_MAS: the product applies the Hermitian part of the reference's half-corrected planes explicitly;
    # after the inverse transform both give the same real field
    g, r = a.copy(), a.copy()
    harness.harness_modes(0, g.ctypes.data, None, N, 4, 4)
    M._lib().oracle_correct_mas_modes(M._cf(r), N, 4)
    xg, xr = M.ifft3d_c2r(g, N), M.ifft3d_c2r(r, N)
    assert np.max(np.abs(xg - xr)) < 2e-6 * np.abs(xr).max()
    assert np.max(np.abs(xr - d1)) > 0.05 * np.abs(d1).max()          # the deconvolution did something


@pytest.mark.parametrize("N", [2, 3, 12, 9])
@pytest.mark.parametrize("nd", [3, 2])
def test_filter_bodies(harness, oracle, N, nd):
    """filter_element (Top-Hat, Gaussian, Top-Hat-k) -> normalise -> transform == the oracle's FT_filter."""
    from oracle import cpu_more as M
    import scipy.fft as sfft
    f32 = np.float32
    for name, (F, R, kmin, kmax) in MC.FILTERS.items():
        R_grid = f32(f32(f32(R) * f32(N)) / f32(MC.BOX))
        R2, kF = f32(R_grid * R_grid), f32(2.0 * np.pi / MC.BOX)
        kind = {"Top-Hat": 0, "Gaussian": 1, "Top-Hat-k": 2}[F]
        if kind == 2:
            out = np.zeros((N,) * (nd - 1) + (N // 2 + 1,), np.complex64)
        else:
            out = np.zeros((N,) * nd, np.float32)
        harness.harness_filter(kind, out.ctypes.data, N, nd, R2, kF, f32(kmin), f32(kmax))
        field = sfft.irfftn(out, s=(N,) * nd, axes=tuple(range(nd))).astype(f32) if kind == 2 else out
        field = (field.astype(np.float64) / float(np.sum(field, dtype=np.float64))).astype(f32)
        got = sfft.rfftn(field, axes=tuple(range(nd))).astype(np.complex64)
        ref = (M.FT_filter if nd == 3 else M.FT_filter_2D)(MC.BOX, R, N, F, 1, kmin, kmax)
        assert np.max(np.abs(got - ref)) < 1e-6 * np.max(np.abs(ref)), (name, nd)


def test_shell_bodies_random_geometries(harness, oracle):
    """Randomised sweep (seeded): grid side, window exponent, line of sight and SM count (i.e. segmentation) drawn
    at random; counts must be exact and sums must match the oracle for every draw."""
    from oracle import cpu as C, cpu_more as M
    rng = np.random.default_rng(2026)
    for _ in range(30):
        N = int(rng.integers(2, 27))
        mas = int(rng.integers(0, 5))
        axis = int(rng.integers(0, 3))
        sms = int(rng.choice([1, 3, 148, 1000, 50000]))
        kmax = C.frequencies(1000.0, N)[4]
        # velocity-divergence estimator on three random fields
        dks = [np.ascontiguousarray(C.fft3d_r2c(rng.standard_normal((N, N, N)).astype(np.float32))) for _ in range(3)]
        got = run_shell(harness, "theta", dks, [mas], N, sms=sms)
        k, Nm, P1, P2, PX = z(kmax + 1), z(kmax + 1), z(kmax + 1), z(kmax + 1), z(kmax + 1)
        M._lib().oracle_vel_bin(0, M._cf(np.ascontiguousarray(np.stack(dks))), N, mas, *[C._dp(a) for a in (k, Nm, P1, P2, PX)])
        assert np.array_equal(got["Nm"], Nm) and int(Nm.sum()) == C.expected_modes(N), (N, sms)
        assert close(got["ksum"], k, 1e-12) and close(got["vals"][0], P1, 2e-6), (N, mas, sms)
        # real-space binning
        grid = rng.standard_normal((N, N, N)).astype(np.float32)
        got = run_shell(harness, "xi", [grid], [], N, axis=axis, scale=1.0, sms=sms)
        r3D, xi3D, Nm = z(kmax + 1), np.zeros((kmax + 1, 3)), z(kmax + 1)
        M._lib().oracle_xi_bin(grid.ctypes.data_as(M._fpp), N, axis, C._dp(r3D), C._dp(xi3D), C._dp(Nm))
        assert np.array_equal(got["Nm"], Nm) and int(Nm.sum()) == N ** 3, (N, sms)
        amp = np.sqrt(np.maximum(Nm, 1.0)) * float(np.abs(grid).max())
        for l in range(3):
            assert float(np.max(np.abs(got["vals"][l] - xi3D[:, l]) / amp)) < 1e-12, (N, axis, l)
