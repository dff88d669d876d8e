"""GPU parity tests of the sibling estimators (Pk_plane, XPk_plane, XPk_imag, XPk_2D, Pk_theta, XPk_dv, XPk_vv,
correct_MAS, expected_Pk, Xi, XXi, field_smoothing) against the golden outputs of the compiled reference and the
CPU oracle.

Bars (BASELINE.json north_star): mode counts bit-exact, wavenumbers / radii 1e-12, spectra and correlation
functions 1e-4 relative per bin.  Cross terms and high multipoles are sums of signed terms, so their error is
measured against max(|value|, 0.1*sqrt(P_i P_j)) (more_cases.compare).  On IDENTICAL transformed input the kernels
are held to 1e-10 (float64 summation order only), 2e-6 where the reference itself works in float32 (k.V)."""
import contextlib
import io
import os

import numpy as np
import pytest

import more_cases as MC
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def env():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from pylians3_b200 import Pk_library as PKL, _pk_more as PM, _lib
    _lib.load()
    return torch, PKL, PM


@pytest.fixture(scope="module")
def golden():
    return dict(np.load(os.path.join(GOLDEN, "pk_more_golden.npz")))


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


@pytest.mark.parametrize("N", MC.SIZES)
def test_siblings_vs_reference_golden(env, golden, N):
    torch, PKL, PM = env
    n0 = __import__("pylians3_b200")._lib.load().pyl_kernel_launches()
    got = MC.run_all(PKL, N)
    assert __import__("pylians3_b200")._lib.load().pyl_kernel_launches() > n0
    assert set(got) == {k for k in golden if k.startswith("N%d_" % N) and "_sm_" not in k and "_xm_" not in k}
    bad = MC.compare(got, golden, tol=TOL, fft_eps=1e-5)
    assert not bad, bad


@pytest.mark.parametrize("N", [64, 45])
def test_siblings_vs_oracle_medium(env, oracle, N):
    """Fresh seeded inputs at a size where every bin holds many modes; the oracle runs the same calls on the CPU."""
    torch, PKL, PM = env
    from oracle import cpu_more
    I = MC.inputs(N)
    ref = MC.run_all(cpu_more, N, {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in I.items()})
    got = MC.run_all(PKL, N, {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in I.items()})
    bad = MC.compare(got, ref, tol=TOL, fft_eps=1e-5)
    assert not bad, bad


def _raw(PM, torch, kind, fields, mas, dims, **kw):
    dev = torch.device("cuda", 0)
    t = [torch.from_numpy(np.ascontiguousarray(f)).to(dev) for f in fields]
    if "table" in kw and kw["table"] is not None:
        tk, tP, kF, lk, dk = kw["table"]
        kw["table"] = (torch.from_numpy(tk).to(dev), torch.from_numpy(tP).to(dev), kF, lk, dk)
    return PM.shell_bin(kind, t, mas, dims, **kw)


def _close(a, b, tol, floor=None):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    den = np.maximum(np.abs(b), 1e-9 * max(np.abs(b).max(), 1e-300))
    if floor is not None:
        den = np.maximum(den, floor)
    return float(np.max(np.abs(a - b) / den)) < tol


@pytest.mark.parametrize("N", [12, 9, 96, 61])
def test_shell_kernels_same_input(env, oracle, N):
    """The kernels on the oracle's own transformed fields: only summation order (float64) and, for k.V, the
    float32 association differ."""
    torch, PKL, PM = env
    from oracle import cpu as C, cpu_more as M
    z = lambda n: np.zeros(n, np.float64)  # noqa: E731
    rng = np.random.default_rng(N)
    F = [rng.standard_normal((N, N, N)).astype(np.float32) for _ in range(7)]
    dks = [np.ascontiguousarray(C.fft3d_r2c(f)) for f in F]
    kmax = C.frequencies(1000.0, N)[4]
    for kind, idx, oidx, mas in (("theta", [1, 2, 3], [1, 2, 3], 2), ("dv", [0, 1, 2, 3], [0, 1, 2, 3], 3),
                                 ("vv", [1, 2, 3, 4, 5, 6], [0, 1, 2, 3, 0, 4, 5, 6], 4)):
        got = _raw(PM, torch, kind, [dks[i] for i in idx], [mas], N)
        stack = np.ascontiguousarray(np.stack([dks[i] for i in oidx]))
        k, Nm, P1, P2, PX = z(kmax + 1), z(kmax + 1), z(kmax + 1), z(kmax + 1), z(kmax + 1)
        M._lib().oracle_vel_bin({"theta": 0, "dv": 1, "vv": 2}[kind], M._cf(stack), N, mas, *[C._dp(a) for a in (k, Nm, P1, P2, PX)])
        assert np.array_equal(got["Nm"], Nm) and int(Nm.sum()) == C.expected_modes(N), kind
        assert _close(got["ksum"], k, 1e-12), kind
        assert _close(got["vals"][0], P1, 2e-6), kind
        if kind != "theta":
            assert _close(got["vals"][1], P2, 2e-6), kind
            assert _close(got["vals"][2], PX, 2e-6, floor=np.sqrt(P1 * P2)), kind
    # images
    imgs = [np.ascontiguousarray(M.fft2d_r2c(rng.standard_normal((N, N)).astype(np.float32))) for _ in range(2)]
    kmax2 = M.frequencies_2D(1000.0, N)[4]
    got = _raw(PM, torch, "xplane", imgs, [2, 4], N)
    k2D, Nm, Pk, PkX = z(kmax2 + 1), z(kmax2 + 1), np.zeros((kmax2 + 1, 2)), z(kmax2 + 1)
    mi = np.array([2, 4], np.int32)
    M._lib().oracle_plane_bin(M._cf(np.ascontiguousarray(np.stack(imgs))), N, 2, mi.ctypes.data_as(M._ipp),
                              C._dp(k2D), C._dp(Nm), C._dp(Pk), C._dp(PkX))
    assert np.array_equal(got["Nm"], Nm) and int(Nm.sum()) == M.expected_modes_2D(N)
    assert _close(got["ksum"], k2D, 1e-12)
    assert _close(got["vals"][0], Pk[:, 0], 1e-10) and _close(got["vals"][1], Pk[:, 1], 1e-10)
    assert _close(got["vals"][2], PkX, 1e-10, floor=np.sqrt(Pk[:, 0] * Pk[:, 1]))
    got1 = _raw(PM, torch, "plane", imgs[:1], [2], N)
    assert np.array_equal(got1["Nm"], Nm) and _close(got1["vals"][0], Pk[:, 0], 1e-10)
    # real-space binning
    for axis in (0, 1, 2):
        grid = F[axis]
        scale = np.float32(1.0 / N ** 3)
        got = _raw(PM, torch, "xi", [grid], [], N, axis=axis, scale=float(scale))
        r3D, xi3D, Nm = z(kmax + 1), np.zeros((kmax + 1, 3)), z(kmax + 1)
        scaled = (grid * scale).astype(np.float32)
        M._lib().oracle_xi_bin(scaled.ctypes.data_as(M._fpp), N, axis, C._dp(r3D), C._dp(xi3D), C._dp(Nm))
        assert np.array_equal(got["Nm"], Nm) and int(Nm.sum()) == N ** 3
        assert _close(got["ksum"], r3D, 1e-12)
        amp = np.sqrt(Nm) * float(np.abs(scaled).max())              # scale of a sum of Nm signed values
        for l in range(3):
            assert _close(got["vals"][l], xi3D[:, l], 1e-10, floor=amp), (axis, l)


@pytest.mark.parametrize("N", [12, 33])
def test_mode_passes_same_input(env, oracle, N):
    torch, PKL, PM = env
    from oracle import cpu as C, cpu_more as M
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(300 + N)
    d1 = rng.standard_normal((N, N, N)).astype(np.float32)
    d2 = rng.standard_normal((N, N, N)).astype(np.float32)
    a, b = np.ascontiguousarray(C.fft3d_r2c(d1)), np.ascontiguousarray(C.fft3d_r2c(d2))
    for second, masb in ((None, 0), (b, 4)):
        g = torch.from_numpy(a.copy()).to(dev)
        s = None if second is None else torch.from_numpy(second).to(dev)
        PM._modes("power", g, s, N, 2, masb)
        r = a.copy()
        M._lib().oracle_xi_modes(M._cf(r), None if second is None else M._cf(second), N, 2, masb)
        assert np.max(np.abs(g.cpu().numpy() - r)) < 2e-6 * np.abs(r).max()
    g = torch.from_numpy(a.copy()).to(dev)
    PM._modes("deconvolve", g, None, N, 4, 0)
    r = a.copy()
    M._lib().oracle_correct_mas_modes(M._cf(r), N, 4)
    xg = PM.ifft3d_c2r_device(g).cpu().numpy()
    xr = M.ifft3d_c2r(r, N)
    assert np.max(np.abs(xg - xr)) < 5e-6 * np.abs(xr).max()


def test_cuda_tensor_inputs_and_side_effects(env, oracle):
    """Zero-copy inputs: same results as host arrays; XPk_dv multiplies the caller's velocity tensors in place
    (Pk_library.pyx:1367), correct_MAS / field_smoothing return device tensors."""
    torch, PKL, PM = env
    from oracle import cpu_more
    N = 24
    I = MC.inputs(N)
    dev = torch.device("cuda", 0)
    V = [torch.from_numpy(I[n].copy()).to(dev) for n in ("Vx1", "Vy1", "Vz1")]
    d = torch.from_numpy(I["d1"]).to(dev)
    got = quiet(PKL.XPk_dv, d, V[0], V[1], V[2], MC.BOX, 2, "CIC", 1)
    Vh = [I[n].copy() for n in ("Vx1", "Vy1", "Vz1")]
    ref = quiet(cpu_more.XPk_dv, I["d1"], Vh[0], Vh[1], Vh[2], MC.BOX, 2, "CIC", 1)
    assert np.array_equal(V[0].cpu().numpy(), Vh[0])                # same float32 product, in place
    assert np.array_equal(got[4], ref[4])
    for g, r in zip(got[1:3], ref[1:3]):
        assert np.max(np.abs(g / r - 1)) < TOL
    back = quiet(PKL.correct_MAS, d, MC.BOX, "None", 1)              # no window: FFT -> IFFT is the identity
    assert back.is_cuda and float((back - d).abs().max()) < 2e-6 * float(d.abs().max())
    assert float((d - torch.from_numpy(I["d1"]).to(dev)).abs().max()) == 0.0        # input untouched
    # smoothing with a Gaussian filter given in Fourier space
    k = np.fft.fftfreq(N, 1.0 / N)
    kz = np.arange(N // 2 + 1)
    k2 = k[:, None, None] ** 2 + k[None, :, None] ** 2 + kz[None, None, :] ** 2
    filt = np.exp(-0.5 * k2 * (2 * np.pi / N * 1.5) ** 2).astype(np.complex64)
    from pylians3_b200 import smoothing_library as SL
    sm = SL.field_smoothing(I["d1"], filt, 1)
    ref_sm = cpu_more.field_smoothing(I["d1"], filt, 1)
    assert isinstance(sm, np.ndarray) and np.max(np.abs(sm - ref_sm)) < 5e-6 * np.abs(ref_sm).max()


def test_properties_at_size(env):
    """Size-independent properties at 256^3 / 2048^2 (the oracle would take minutes there)."""
    torch, PKL, PM = env
    N = 256
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(5)
    V = [torch.randn((N, N, N), generator=g, device=dev, dtype=torch.float32) for _ in range(3)]
    k, P, Nm = quiet(PKL.Pk_theta, V[0], V[1], V[2], MC.BOX, 2, "None", 1)
    own = 8
    assert int(Nm.sum()) == (N ** 3 - own) // 2 + own - 1            # every independent mode but the DC one
    k2, P2, Nm2 = quiet(PKL.Pk_theta, 2 * V[0], 2 * V[1], 2 * V[2], MC.BOX, 2, "None", 1)
    assert np.array_equal(Nm, Nm2) and np.max(np.abs(P2 / P - 4.0)) < 1e-12      # exact scaling by a power of two
    # unit white noise in each component: <|k.V_k|^2> = |k|^2 N^3  ->  P_theta / k^2 = L^3/N^3
    lo = (k > 0.05) & (k < 0.4)
    shot = MC.BOX ** 3 / N ** 3
    assert abs(np.mean(P[lo] / k[lo] ** 2) / shot - 1.0) < 0.05
    # Xi of white noise: xi(0) = variance, xi(r>0) ~ 0; the DC bin is dropped, so check the sum rule instead:
    d = V[0]
    x = quiet(PKL.Xi, d, MC.BOX, "None", 2, 1)
    assert int(x.Nmodes3D.sum()) == N ** 3 - 1
    assert np.max(np.abs(x.xi[:, 0])) < 0.05                          # no correlation beyond r = 0
    # image
    M = 2048
    img = torch.randn((M, M), generator=g, device=dev, dtype=torch.float32)
    p = quiet(PKL.Pk_plane, img, MC.BOX, "None", 1, False)
    assert int(p.Nmodes.sum()) == (M * M - 4) // 2 + 4 - 1
    assert abs(np.mean(p.Pk[p.Nmodes > 1000]) / (MC.BOX ** 2 / M ** 2) - 1.0) < 0.02


@pytest.mark.parametrize("N,F", [(20, 6), (15, 5)])
def test_xpk_imag_more_fields_than_one_launch(env, oracle, N, F):
    """More than PYL_MAX_FIELDS fields: block decomposition over launches, pairs in the reference's order, and
    the antisymmetric cross term keeps its sign (i < j)."""
    torch, PKL, PM = env
    from oracle import cpu_more
    rng = np.random.default_rng(N * F)
    base = rng.standard_normal((N, N, N)).astype(np.float32)
    fields = [(np.roll(base, i, axis=i % 3) + 0.5 * rng.standard_normal((N, N, N))).astype(np.float32) for i in range(F)]
    mas = (["CIC", "PCS", "NGP", "TSC", "None", "CIC"])[:F]
    got = quiet(PKL.XPk_imag, fields, MC.BOX, 1, mas, 1)
    ref = quiet(cpu_more.XPk_imag, fields, MC.BOX, 1, mas, 1)
    X = F * (F - 1) // 2
    assert got.XPk.shape == ref.XPk.shape == (ref.k3D.size, 3, X)
    assert np.array_equal(got.Nmodes3D, ref.Nmodes3D) and np.array_equal(got.Nmodes2D, ref.Nmodes2D)
    pairs = [(i, j) for i in range(F) for j in range(i + 1, F)]
    for x, (i, j) in enumerate(pairs):
        scale = np.sqrt(np.abs(ref.Pk[:, 0, i] * ref.Pk[:, 0, j]))
        assert np.max(np.abs(got.XPk[:, 0, x] - ref.XPk[:, 0, x]) / np.maximum(np.abs(ref.XPk[:, 0, x]), 0.1 * scale)) < TOL * 3
    assert np.max(np.abs(got.Pk[:, 0, :] / ref.Pk[:, 0, :] - 1)) < TOL


@pytest.mark.parametrize("N", list(MC.SIZES) + [64, 45])
def test_smoothing_library(env, oracle, golden, N):
    """FT_filter / FT_filter_2D (Top-Hat, Gaussian, Top-Hat-k), field_smoothing and field_smoothing_2D against the
    golden outputs of the compiled reference (small sizes) and the oracle; amplitudes to 1e-5 of the peak
    (two float32 FFT libraries)."""
    torch, PKL, PM = env
    from oracle import cpu_more
    from pylians3_b200 import smoothing_library as SL
    got = MC.run_smoothing(SL, N)
    ref = golden if N in MC.SIZES else MC.run_smoothing(cpu_more, N)
    bad = MC.compare_smoothing(got, ref, tol=1e-5)
    assert not bad, bad
    # device-resident chain: filter stays on the GPU, field is a CUDA tensor, result is a CUDA tensor
    I = MC.inputs(N)
    fk = SL.FT_filter(MC.BOX, 90.0, N, "Gaussian", 1, as_tensor=True)
    d = torch.from_numpy(I["d1"]).cuda()
    sm = SL.field_smoothing(d, fk, 1)
    assert sm.is_cuda and float((sm.cpu() - torch.from_numpy(np.asarray(ref["N%d_sm_s3" % N]))).abs().max()) < \
        1e-5 * float(np.abs(ref["N%d_sm_s3" % N]).max())
    # a normalised filter preserves the mean of the field (its DC mode is 1)
    assert abs(float(sm.double().mean()) - float(d.double().mean())) < 1e-6 * float(d.abs().max())


@pytest.mark.parametrize("N", MC.SIZES)
def test_xxi_multi(env, golden, N):
    """XXi_multi is a composition of the primitives tested above (pyl_modes_power -> pyl_fft_c2r -> pyl_shell_bin);
    its Python layer is also driven on the CPU with the kernel bodies (test_pk_more_host_glue.py).  Added after
    round 1's GPU budget was spent: first run on a GPU happens in the round-end suite."""
    torch, PKL, PM = env
    bad = MC.compare_xxi_multi(MC.run_xxi_multi(PKL, N), golden, TOL, N)
    assert not bad, bad


def test_xxi_projected_vs_golden_and_oracle():
    """XXi_projected on the GPU (2D r2c -> pyl_modes_power_2d -> 2D c2r -> pyl_radial_bin_2d) against the golden
    outputs of the compiled reference (N = 12, 9, 32) and against the oracle on larger images: counts bit-exact,
    r_p 1e-12, xi_p 1e-4 relative plus the float32-FFT floor (1e-5 of the peak)."""
    import contextlib
    import io
    import os

    import numpy as np
    from conftest import BOX, GOLDEN
    from golden.make_golden_more import inputs
    from oracle import cpu_more as OM
    from pylians3_b200 import Pk_library as PKL
    g = np.load(os.path.join(GOLDEN, "xxi_projected_golden.npz"))

    def check(r, Nm, rp, xi):
        assert np.array_equal(r.Nmodes_p, Nm)
        assert np.max(np.abs(r.r_p - rp) / rp) < 1e-12
        assert np.all(np.abs(r.xi_p - xi) <= 1e-4 * np.abs(xi) + 1e-5 * np.max(np.abs(xi)))
    for N in (12, 9, 32):
        I = inputs(N)
        with contextlib.redirect_stdout(io.StringIO()):
            r = PKL.XXi_projected(I["img1"], I["img2"], BOX, ["CIC", "PCS"], 1)
        check(r, g["N%d_Nm" % N], g["N%d_r" % N], g["N%d_xi" % N])
    rng = np.random.default_rng(3)
    for N, mas in ((256, ["TSC", "None"]), (45, ["NGP", "PCS"]), (2048, ["CIC", "CIC"])):
        a = rng.standard_normal((N, N)).astype(np.float32)
        b = (0.5 * a + rng.standard_normal((N, N))).astype(np.float32)
        with contextlib.redirect_stdout(io.StringIO()):
            r = PKL.XXi_projected(a, b, BOX, mas, 1)
            o = OM.XXi_projected(a, b, BOX, mas, 1)
        check(r, o.Nmodes_p, o.r_p, o.xi_p)
        assert int(r.Nmodes_p.sum()) + 1 == N * N          # every cell of the image lands in a bin (bin 0 = the DC cell)
