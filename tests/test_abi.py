"""CPU: the C-ABI shared library loads and exports every symbol include/pyl_b200.h declares; the
host-side logic (layout, finalisation, error behaviour) matches the reference's.  No compute
calls are made here (there is no GPU in this tier)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import BOX, ROOT, rel_err


def header_symbols():
    src = open(os.path.join(ROOT, "include", "pyl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pyl_[A-Za-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from pylians3_b200 import build, _lib
    build.build()
    return _lib.load()


def test_every_declared_symbol_is_exported_and_bound(lib):
    from pylians3_b200 import _lib
    declared = header_symbols()
    assert len(declared) >= 25
    raw = ctypes.CDLL(_lib.SO_PATH)
    for name in declared:
        assert hasattr(raw, name), "libpyl_b200.so does not export %s" % name
        assert name in _lib.PROTOTYPES, "no ctypes prototype for %s" % name
    assert sorted(_lib.PROTOTYPES) == declared


def test_version_and_error_strings(lib):
    assert b"sm_100a" in lib.pyl_version()
    assert lib.pyl_error_string(0) == b"ok"
    assert lib.pyl_error_string(-4).startswith(b"workspace")


def test_no_torch_types_in_abi():
    src = open(os.path.join(ROOT, "include", "pyl_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)        # declarations only, comments stripped
    assert "torch" not in code.lower() and "at::" not in code and "c10::" not in code
    assert "#include <cuda" not in code                      # plain C consumers (Cython, cgo, ctypes)


@pytest.mark.parametrize("dims,F", [(16, 1), (15, 3), (512, 1), (4096, 1), (2048, 3)])
def test_layout_matches_frequencies(lib, oracle, dims, F):
    from pylians3_b200 import _lib, Pk_library as PKL
    L = _lib.pk_layout(dims, F)
    kF, kN, kmax_par, kmax_per, kmax = oracle.frequencies(BOX, dims)
    assert (L.kmax_par, L.kmax_per, L.kmax) == (kmax_par, kmax_per, kmax)
    assert PKL.frequencies(BOX, dims) == (kF, kN, kmax_par, kmax_per, kmax)
    X = F * (F - 1) // 2
    n3, n1, n2 = kmax + 1, kmax_par + 1, (kmax_par + 1) * (kmax_per + 1)
    assert L.n2d == n2
    assert L.total_words == n3 * (3 + 3 * F + 3 * X) + n1 * (1 + F + X) + n2 * (1 + F + X)
    # sections are disjoint and ordered
    offs = [L.k3D, L.Nm3D, L.Pk3D, L.PkX3D, L.phase, L.Nm1D, L.Pk1D, L.PkX1D, L.Nm2D, L.Pk2D, L.PkX2D]
    assert offs == sorted(offs) and offs[0] == 0


def test_workspace_queries_without_gpu(lib):
    assert lib.pyl_deposit_workspace_bytes(1, 1000, 64, 3, 0) == 0          # atomic needs none
    assert lib.pyl_pk_bin_workspace_bytes(64, 1) > 0
    assert lib.pyl_pk_bin_workspace_bytes(64, 9) == 0                       # beyond PYL_MAX_FIELDS


def test_argument_errors_return_status_not_crash(lib):
    from pylians3_b200 import _lib
    st = lib.pyl_deposit(7, None, None, None, 10, 8, 3, 1.0, 0, None, 0, None)
    assert st == -1 and b"unknown scheme" in lib.pyl_last_error()
    st = lib.pyl_deposit(1, None, None, None, 10, 8, 4, 1.0, 0, None, 0, None)
    assert st == -1
    st = lib.pyl_pk_bin(None, 1, None, 16, 0, 16, 2, 0, None, None, 0, None)
    assert st == -1
    with pytest.raises(_lib.PylError):
        _lib.check(st, "pyl_pk_bin")


def test_reference_error_behaviour(capsys):
    from pylians3_b200 import MAS_library as MASL
    from pylians3_b200.errors import ReferenceExit
    # dimension mismatch: message + exit (MAS_library.pyx:64-66)
    with pytest.raises(SystemExit) as e:
        MASL.MA(np.zeros((4, 3), np.float32), np.zeros((4, 4), np.float32), 1.0)
    assert isinstance(e.value, ValueError) and isinstance(e.value, ReferenceExit)
    assert "pos have 3 dimensions and the density 2!!!" in capsys.readouterr().out
    # unknown scheme (MAS_library.pyx:82)
    with pytest.raises(ValueError):
        MASL.MA(np.zeros((4, 3), np.float32), np.zeros((4, 4, 4), np.float32), 1.0, "XYZ")
    assert "option not valid!!!" in capsys.readouterr().out


def test_finalisation_matches_reference_loops(oracle, pk_golden):
    """Host logic: feed the oracle's RAW accumulators to the product's vectorised finalisation and
    compare with the golden outputs of the reference (Pk_library.pyx:384-418, 735-791)."""
    from pylians3_b200 import Pk_library as PKL
    for N in (16, 15):
        delta = pk_golden["N%d_delta" % N]
        for axis in (0, 2):
            raw = oracle.bin_raw([oracle.fft3d_r2c(d) for d in delta], N, [4, 2, 0], axis, BOX)
            raw["phase"] = np.zeros_like(raw["k3D"])
            o = PKL._finalize(raw, BOX, N)
            t = "N%d_XPk_a%d_" % (N, axis)
            for nm in ("k3D", "Pk", "XPk", "Nmodes3D", "k1D", "Pk1D", "PkX1D", "Nmodes1D", "kpar", "kper",
                       "Pk2D", "PkX2D", "Nmodes2D"):
                ref = pk_golden[t + nm]
                tol_floor = 1e-9 * float(np.nanmax(np.abs(pk_golden[t + "Pk"])))
                assert o[nm].shape == ref.shape, nm
                assert rel_err(o[nm], ref, floor=1e-300) < 1e-9 or \
                    float(np.nanmax(np.abs(o[nm] - ref))) < tol_floor, nm


def test_missing_extension_fails_loudly(monkeypatch):
    from pylians3_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "SO_PATH", "/nonexistent/libpyl_b200.so")
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.load()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pylians3_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("# oracle", ""), f


@pytest.mark.parametrize("dims", [16, 15, 512, 4096])
def test_shell_layout_matches_frequencies(lib, oracle, dims):
    """pyl_shell_layout: bins = kmax+1 of frequencies (3D kinds) / frequencies_2D (images), values per kind."""
    from pylians3_b200 import _lib, Pk_library as PKL
    from oracle import cpu_more
    nvals = {"theta": 1, "dv": 3, "vv": 3, "expected": 1, "plane": 1, "xplane": 3, "xi": 3}
    for kind, kid in _lib.SHELL_KINDS.items():
        b, v = ctypes.c_int(0), ctypes.c_int(0)
        assert lib.pyl_shell_layout(kid, dims, ctypes.byref(b), ctypes.byref(v)) == 0
        kmax = (cpu_more.frequencies_2D if kind in ("plane", "xplane") else oracle.frequencies)(BOX, dims)[4]
        assert (b.value, v.value) == (kmax + 1, nvals[kind]), kind
        assert lib.pyl_shell_bin_workspace_bytes(kid, dims) >= 8 * 8 * (2 + v.value) * b.value
    assert PKL.frequencies_2D(BOX, dims) == cpu_more.frequencies_2D(BOX, dims)
    assert lib.pyl_shell_layout(9, dims, None, None) == -1


def test_sibling_argument_errors(lib):
    # wrong field count for the kind, NULL output, missing table, missing workspace: status codes, no crash
    one = (ctypes.c_void_p * 1)(ctypes.c_void_p(16))
    mi = (ctypes.c_int * 2)(2, 2)
    assert lib.pyl_shell_bin(0, one, 1, mi, 16, 2, 1.0, None, ctypes.c_void_p(16), None, 0, None) == -1
    three = (ctypes.c_void_p * 3)(16, 16, 16)
    assert lib.pyl_shell_bin(0, three, 3, mi, 16, 2, 1.0, None, None, None, 0, None) == -1
    assert lib.pyl_shell_bin(0, three, 3, mi, 16, 2, 1.0, None, ctypes.c_void_p(16), None, 0, None) == -4
    assert lib.pyl_shell_bin(3, one, 0, None, 16, 2, 1.0, None, ctypes.c_void_p(16), None, 0, None) == -1
    assert lib.pyl_modes_deconvolve(None, 16, 2, None, 0, None) == -1
    assert lib.pyl_modes_deconvolve(ctypes.c_void_p(16), 16, 7, None, 0, None) == -1
    assert lib.pyl_modes_power(ctypes.c_void_p(16), None, 16, 2, 2, None, 0, None) == -4
    assert lib.pyl_fft_c2r(None, None, 16, None, 0, None) == -1
    assert lib.pyl_cmul_inplace(None, None, 10, None) == -1 and lib.pyl_cmul_inplace(None, None, 0, None) == 0
    assert lib.pyl_mul_one_plus(None, None, 10, None) == -1


def test_sibling_python_errors_match_reference(capsys):
    """Host-side argument checks happen before any device work is attempted (messages of Pk_library.pyx:1119,
    1773, 1975, 1980, 2323 and smoothing_library.pyx:223)."""
    from pylians3_b200 import Pk_library as PKL, smoothing_library as SL, _device
    if _device.torch.cuda.is_available():
        pytest.skip("checked on the CPU tier")
    k = np.logspace(-3, 1, 50).astype(np.float32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        PKL.Pk_theta(np.zeros((4, 4, 4), np.float32), np.zeros((4, 4, 4), np.float32), np.zeros((4, 4, 4), np.float32), 1.0)
    for name in ("Pk_plane", "XPk_plane", "XPk_imag", "XPk_2D", "Pk_theta", "XPk_dv", "XPk_vv", "correct_MAS",
                 "expected_Pk", "Xi", "XXi", "IFFT3Dr_f", "FFT2Dr_f", "frequencies_2D", "check_number_modes_2D"):
        assert hasattr(PKL, name), name
    assert callable(SL.field_smoothing)
