"""CPU: the product's Python layer for the sibling estimators (pylians3_b200/_pk_more.py: argument handling, call
sequence, unit conversions) driven end to end WITHOUT a GPU.  The device primitives it calls are monkeypatched with
test doubles: transforms by pocketfft, and -- for the kernels -- the real per-thread bodies of shell_body.cuh run
serially by tests/harness/shell_host.cpp.  Results are held to the golden outputs of the compiled reference.
The product itself has no such path: without the patches every entry point raises (no CUDA device)."""
import os

import numpy as np
import pytest
import scipy.fft as sfft

import more_cases as MC
from conftest import GOLDEN
from test_shell_bodies_host import harness, run_shell  # noqa: F401  (fixture + helper)


@pytest.fixture(params=["pocketfft32", "rounded64"])
def fake_device(request, monkeypatch, harness):  # noqa: F811
    """params: the forward transform double.  "pocketfft32" is the transform the golden reference used;
    "rounded64" (a float64 transform rounded to complex64) stands in for ANOTHER float32 FFT library -- cuFFT on
    the GPU -- and rehearses the tolerances of tests/test_gpu_pk_more.py, FFT floor included."""
    from pylians3_b200 import _pk_more as PM, Pk_library as P, _device as D
    other_fft = request.param == "rounded64"

    def rfft(d, axes):
        x = d.astype(np.float64) if other_fft else d
        return np.ascontiguousarray(sfft.rfftn(x, axes=axes).astype(np.complex64))

    def as_np(x, dev=None, name=None):
        assert x.dtype == np.float32
        return np.ascontiguousarray(x)

    def shell_bin(kind, fields, mas_index, dims, axis=2, scale=1.0, table=None):
        if table is not None:
            tk, tP, kF32, lk, dk = table
            table = (tk, tP, float(kF32), lk, dk)
        r = run_shell(harness, kind, fields, mas_index, dims, axis=axis, scale=scale, table=table)
        return dict(ksum=r["ksum"], Nm=r["Nm"], vals=r["vals"])

    def modes(op, a_k, b_k, dims, mas_a, mas_b):
        harness.harness_modes(0 if op == "deconvolve" else 1, a_k.ctypes.data, None if b_k is None else b_k.ctypes.data,
                              dims, mas_a, mas_b if b_k is not None else mas_a)

    def c2r(ak, normalise=True):
        n = ak.shape[0]
        out = sfft.irfftn(ak.astype(np.complex128) if other_fft else ak, s=(n, n, n), axes=(0, 1, 2)).astype(np.float32)
        return out if normalise else (out * np.float32(n ** 3)).astype(np.float32)

    def momentum(V, delta_d, dev, names):
        for v in V:
            v *= (np.float32(1.0) + delta_d)
        return list(V)

    monkeypatch.setattr(D, "require_cuda", lambda: None)
    monkeypatch.setattr(D, "pick_device", lambda *a: None)
    monkeypatch.setattr(PM, "_cube", as_np)
    monkeypatch.setattr(PM, "_as_image", as_np)
    monkeypatch.setattr(P, "fft3d_r2c_device", lambda d: rfft(d, (0, 1, 2)))
    monkeypatch.setattr(PM, "fft2d_r2c_device", lambda d: rfft(d, (0, 1)))
    monkeypatch.setattr(PM, "ifft3d_c2r_device", c2r)
    monkeypatch.setattr(PM, "shell_bin", shell_bin)
    monkeypatch.setattr(PM, "_modes", modes)
    monkeypatch.setattr(PM, "_momentum_", momentum)
    monkeypatch.setattr(D, "is_cuda_tensor", lambda x: True)          # "device in -> device out": keep arrays as they are

    class _T:                                                          # torch.from_numpy(x).to(dev) -> x ; torch.device(...)
        @staticmethod
        def from_numpy(x):
            class W:
                def to(self, dev):
                    return x
            return W()

        @staticmethod
        def device(*a):
            return None

        @staticmethod
        def clone(x):
            return x.copy()

        class cuda:
            @staticmethod
            def current_device():
                return 0
    monkeypatch.setattr(PM, "torch", _T)
    PM.other_fft = other_fft
    return PM


def _impl(PM):

    class Impl:                                 # the entry points whose kernels can run on the host harness
        Pk_plane, XPk_plane, Pk_theta, XPk_dv, XPk_vv = PM.Pk_plane, PM.XPk_plane, PM.Pk_theta, PM.XPk_dv, PM.XPk_vv
        correct_MAS, expected_Pk, Xi, XXi = PM.correct_MAS, PM.expected_Pk, PM.Xi, PM.XXi

        class _Skip:
            def __getattr__(self, n):
                return np.zeros(0)

        @staticmethod
        def XPk_imag(*a, **k):
            return Impl._Skip()

        @staticmethod
        def XPk_2D(*a, **k):
            return [np.zeros(0)] * 6

    return Impl


@pytest.mark.parametrize("N", MC.SIZES)
def test_python_layer_against_reference_golden(fake_device, N):
    PM = fake_device
    golden = dict(np.load(os.path.join(GOLDEN, "pk_more_golden.npz")))
    got = MC.run_all(_impl(PM), N)
    got = {k: v for k, v in got.items() if "imag" not in k and "x2d" not in k}
    if PM.other_fft:
        bad = MC.compare(got, golden, tol=1e-4, fft_eps=1e-5)      # the GPU tests' bars
    else:
        bad = MC.compare(got, golden, tol=1e-5)     # float32 transforms differ in association (scaled c2r)
    assert not bad, bad


@pytest.mark.parametrize("N", MC.SIZES)
def test_xxi_multi_python_layer(fake_device, N):
    PM = fake_device
    golden = dict(np.load(os.path.join(GOLDEN, "pk_more_golden.npz")))
    bad = MC.compare_xxi_multi(MC.run_xxi_multi(PM, N), golden, 1e-4 if PM.other_fft else 2e-5, N)
    assert not bad, bad


@pytest.mark.parametrize("N", [64, 45])
def test_python_layer_against_oracle_medium(fake_device, oracle, N):
    """The medium-size comparison of the GPU suite, rehearsed on the CPU."""
    PM = fake_device
    from oracle import cpu_more
    I = MC.inputs(N)
    cp = lambda: {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in I.items()}  # noqa: E731
    ref = MC.run_all(cpu_more, N, cp())
    got = MC.run_all(_impl(PM), N, cp())
    got = {k: v for k, v in got.items() if "imag" not in k and "x2d" not in k}
    bad = MC.compare(got, ref, tol=1e-4, fft_eps=1e-5 if PM.other_fft else 0.0)
    assert not bad, bad
