"""GPU parity tests of the FFT + deconvolution + binning stage against the CPU oracle and the golden
vectors of the compiled reference.

Tolerances (BASELINE.json north_star): Nmodes* and bin indices bit-exact; k3D/k1D 1e-12; Pk, XPk,
Pk1D, Pk2D 1e-4 relative per bin.  Multipoles and cross spectra can be pure round-off in a bin
(e.g. the corner-mode quadrupole, SURVEY 8a note ii), so their floor is 1e-4*(2l+1)*P0(bin)."""
import contextlib
import io

import numpy as np
import pytest

from conftest import BOX, make_particles, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def env():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from pylians3_b200 import MAS_library as MASL, Pk_library as PKL, _lib
    _lib.load()
    return torch, MASL, PKL, _lib


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


EPS_FFT = 1e-5


def _tol(ref, scale, tol):
    """|got-ref| <= tol*|ref| + EPS_FFT*sqrt(scale*max(scale)) + EPS_FFT^2*max(scale).

    The first term is the north_star bound.  The second is the single-precision FFT floor: two float32
    FFTs (cuFFT here, pocketfft in the oracle, FFTW in the reference) agree per mode to ~1e-6 of the
    spectrum's LARGEST amplitudes, not of each mode's own, so a bin whose power is orders of magnitude
    below the peak carries an absolute error ~ eps*sqrt(P_bin*P_peak) (amplitude error x amplitude).
    With tol=TOL and same-delta_k inputs (test_bin_kernel_same_delta_k) the second term is dropped."""
    scale = np.abs(np.where(np.isnan(scale), 0.0, scale))
    peak = float(np.max(scale)) if scale.size else 0.0
    # third term: a mode whose own amplitude is ~0 (the DC mode of a zero-mean field) still carries the
    # amplitude error eps*A_peak, i.e. a power of eps^2 * P_peak
    return tol * np.abs(ref) + EPS_FFT * np.sqrt(scale * peak) + EPS_FFT ** 2 * peak


def check_pk(got, ref, cross=False, tol=TOL, fft_floor=True, phase_min_modes=0):
    """phase_min_modes: Pkphase (a scale-free angle per mode) of a nearly empty shell is dominated by the float32
    FFT noise of its few low-amplitude modes; callers whose transform differs from the oracle's in structure
    (slab FFT) compare it only on shells with at least that many modes."""
    global EPS_FFT
    eps_saved = EPS_FFT
    if not fft_floor:
        EPS_FFT = 1e-11            # float64 accumulation-order noise only
    try:
        _check_pk(got, ref, cross, tol, phase_min_modes)
    finally:
        EPS_FFT = eps_saved


def _check_pk(got, ref, cross, tol, phase_min_modes=0):
    for nm in ("Nmodes3D", "Nmodes1D", "Nmodes2D"):
        assert np.array_equal(np.asarray(getattr(got, nm)), np.asarray(getattr(ref, nm))), nm
    for nm in ("k3D", "k1D", "kpar", "kper"):
        assert rel_err(getattr(got, nm), np.asarray(getattr(ref, nm)), 1e-300) < 1e-12, nm
    P_ref = np.asarray(ref.Pk)
    P_got = np.asarray(got.Pk)
    if P_ref.ndim == 2:
        P_ref, P_got = P_ref[:, :, None], P_got[:, :, None]
    F = P_ref.shape[2]
    ell = np.array([1.0, 5.0, 9.0])[None, :, None]
    P0 = np.abs(P_ref[:, 0:1, :])
    assert np.all(np.abs(P_got - P_ref) <= _tol(P_ref, P0 * ell, tol) + tol * P0 * ell * (ell > 1)), "Pk"
    assert np.all(np.abs(P_got[:, 0] - P_ref[:, 0]) <= _tol(P_ref[:, 0], P0[:, 0], tol)), "P0"

    def auto(nm):
        a, b = np.asarray(getattr(got, nm)), np.asarray(getattr(ref, nm))
        if a.ndim == 1:
            a, b = a[:, None], b[:, None]
        return a, b
    for nm in ("Pk1D", "Pk2D"):
        a, b = auto(nm)
        assert np.array_equal(np.isnan(a), np.isnan(b)), nm
        ok = np.abs(a - b) <= _tol(b, b, tol)
        assert np.all(ok | np.isnan(b)), nm
    if hasattr(ref, "Pkphase") and not cross:
        sel = np.asarray(ref.Nmodes3D) >= phase_min_modes
        assert rel_err(np.asarray(got.Pkphase)[sel], np.asarray(ref.Pkphase)[sel], 1e-300) < max(tol, 2e-6), "Pkphase"
    if cross:
        Xr, Xg = np.asarray(ref.XPk), np.asarray(got.XPk)
        ix = 0
        for i in range(F):
            for j in range(i + 1, F):
                amp = np.sqrt(P0[:, :, i] * P0[:, :, j]) * ell[:, :, 0]      # |cross| <= sqrt(auto_i*auto_j)
                assert np.all(np.abs(Xg[:, :, ix] - Xr[:, :, ix]) <= _tol(Xr[:, :, ix], amp, tol) + tol * amp), "XPk"
                for nm, anm in (("PkX1D", "Pk1D"), ("PkX2D", "Pk2D")):
                    a, b = np.asarray(getattr(got, nm))[:, ix], np.asarray(getattr(ref, nm))[:, ix]
                    au = np.asarray(getattr(ref, anm))
                    amp1 = np.sqrt(np.abs(au[:, i] * au[:, j]))
                    assert np.array_equal(np.isnan(a), np.isnan(b)), nm
                    ok = np.abs(a - b) <= _tol(b, amp1, tol) + tol * np.where(np.isnan(amp1), 0, amp1)
                    assert np.all(ok | np.isnan(b)), nm
                ix += 1


def device_spectrum(PKL, torch, dk_list, mas_list, N, axis, phase):
    """The product's bin kernel + finalisation applied to GIVEN half-spectra (numpy complex64)."""
    dk = [torch.from_numpy(np.ascontiguousarray(d)).cuda() for d in dk_list]
    raw = PKL.bin_fields(dk, [PKL.MAS_function(m) for m in mas_list], N, axis, want_phase=phase)
    o = PKL._finalize(raw, BOX, N)

    class R:
        pass
    r = R()
    r.k3D, r.Nmodes3D, r.k1D, r.Nmodes1D = o["k3D"], o["Nmodes3D"], o["k1D"], o["Nmodes1D"]
    r.kpar, r.kper, r.Nmodes2D = o["kpar"], o["kper"], o["Nmodes2D"]
    if len(dk_list) == 1 and phase:
        r.Pk, r.Pk1D, r.Pk2D, r.Pkphase = o["Pk"][:, :, 0], o["Pk1D"][:, 0], o["Pk2D"][:, 0], o["Pkphase"]
    else:
        r.Pk, r.XPk, r.Pk1D, r.PkX1D, r.Pk2D, r.PkX2D = o["Pk"], o["XPk"], o["Pk1D"], o["PkX1D"], o["Pk2D"], o["PkX2D"]
    return r


@pytest.mark.parametrize("N,F", [(64, 1), (33, 1), (48, 2), (40, 3), (32, 4), (21, 3)])
@pytest.mark.parametrize("axis", [0, 1, 2])
def test_bin_kernel_same_delta_k(env, oracle, N, F, axis):
    """The hand-written deconvolution+binning kernel against the oracle loop on IDENTICAL delta_k
    (the oracle's own FFT output): only float64 summation order differs -> 1e-10, counts exact.
    Strongly clustered PCS/CIC/NGP/TSC fields: ~4 decades of dynamic range in P(k)."""
    torch, MASL, PKL, _ = env
    fields, mas = make_fields(oracle, N, F, 500 + N + F)
    dks = [oracle.fft3d_r2c(f) for f in fields]
    got = device_spectrum(PKL, torch, dks, mas, N, axis, phase=(F == 1))
    if F == 1:
        ref = oracle.Pk(fields[0], BOX, axis, mas[0], 1, False)
    else:
        ref = quiet(oracle.XPk, fields, BOX, axis, mas, 1)
    check_pk(got, ref, cross=(F > 1), tol=1e-10, fft_floor=False)


@pytest.mark.parametrize("N", [16, 15, 64, 33])
def test_fft_r2c_matches_float64_fft(env, N):
    torch, MASL, PKL, _ = env
    d = np.random.default_rng(N).standard_normal((N, N, N)).astype(np.float32)
    got = PKL.FFT3Dr_f(d, 1)
    assert got.dtype == np.complex64 and got.shape == (N, N, N // 2 + 1)
    ref = np.fft.rfftn(d.astype(np.float64))
    assert np.max(np.abs(got - ref)) < 5e-6 * np.max(np.abs(ref))
    assert np.array_equal(d, np.random.default_rng(N).standard_normal((N, N, N)).astype(np.float32))  # input untouched


@pytest.mark.parametrize("N", [16, 15])
@pytest.mark.parametrize("axis", [0, 1, 2])
def test_pk_vs_reference_golden(env, pk_golden, N, axis):
    torch, MASL, PKL, _ = env
    delta = pk_golden["N%d_delta" % N]

    class R:
        pass
    for fi, mas in ((0, "PCS"), (1, "CIC"), (2, "NGP"), (0, "TSC"), (0, None)):
        ref = R()
        t = "N%d_Pk_a%d_f%d_%s_" % (N, axis, fi, mas)
        for nm in ("k3D", "Pk", "Nmodes3D", "Pkphase", "k1D", "Pk1D", "Nmodes1D", "kpar", "kper", "Pk2D", "Nmodes2D"):
            setattr(ref, nm, pk_golden[t + nm])
        got = PKL.Pk(delta[fi], BOX, axis, mas, 1, verbose=False)
        check_pk(got, ref)
    ref = R()
    t = "N%d_XPk_a%d_" % (N, axis)
    for nm in ("k3D", "Pk", "XPk", "Nmodes3D", "k1D", "Pk1D", "PkX1D", "Nmodes1D", "kpar", "kper", "Pk2D", "PkX2D",
               "Nmodes2D"):
        setattr(ref, nm, pk_golden[t + nm])
    got = quiet(PKL.XPk, [delta[0], delta[1], delta[2]], BOX, axis, ["PCS", "CIC", "None"], 1)
    check_pk(got, ref, cross=True)
    assert got.Pk.shape == ref.Pk.shape and got.XPk.shape == ref.XPk.shape and got.PkX2D.shape == ref.PkX2D.shape


def make_fields(oracle, N, nfields, seed):
    pos, W = make_particles(seed, 3 * N ** 3, True)
    specs = [("PCS", None), ("CIC", W), ("NGP", None), ("TSC", W * W), ("CIC", None), ("PCS", W)]
    out, mas = [], []
    for m, w in specs[:nfields]:
        g = np.zeros((N, N, N), np.float32)
        oracle.MA(pos, g, BOX, m, w)
        g /= np.mean(g, dtype=np.float64)
        g -= 1.0
        out.append(g)
        mas.append(m)
    return out, mas


@pytest.mark.parametrize("N", [64, 48, 33, 128])
@pytest.mark.parametrize("axis", [0, 1, 2])
def test_pk_vs_oracle_medium(env, oracle, N, axis):
    torch, MASL, PKL, _ = env
    fields, mas = make_fields(oracle, N, 1, N + axis)
    ref = oracle.Pk(fields[0], BOX, axis, mas[0], 1, False)
    got = PKL.Pk(fields[0], BOX, axis, mas[0], threads=4, verbose=False)
    check_pk(got, ref)
    # device-resident input gives the same answer and is not modified
    d = torch.from_numpy(fields[0]).cuda()
    got2 = PKL.Pk(d, BOX, axis, mas[0], verbose=False)
    assert np.array_equal(got2.Pk, got.Pk) or rel_err(got2.Pk[:, 0], got.Pk[:, 0], 1e-300) < 1e-12
    assert torch.equal(d.cpu(), torch.from_numpy(fields[0]))


@pytest.mark.parametrize("N,F", [(48, 2), (64, 3), (33, 3), (32, 4), (24, 5), (24, 6)])
def test_xpk_vs_oracle(env, oracle, N, F):
    torch, MASL, PKL, _ = env
    fields, mas = make_fields(oracle, N, F, 100 + N + F)
    if F >= 3:
        mas[2] = "None"
    for axis in (0, 2):
        ref = quiet(oracle.XPk, fields, BOX, axis, mas, 1)
        got = quiet(PKL.XPk, fields, BOX, axis, mas, 2)
        check_pk(got, ref, cross=True)


def test_slab_windows_add_up_to_the_full_binning(env, oracle):
    """The ky-window form used after the distributed transpose: partial results over disjoint windows
    sum to the single-GPU result (counts exactly)."""
    torch, MASL, PKL, _ = env
    N = 40
    fields, mas = make_fields(oracle, N, 2, 77)
    dk = [PKL.fft3d_r2c_device(torch.from_numpy(f).cuda()) for f in fields]
    mi = [PKL.MAS_function(m) for m in mas]
    for axis in (0, 1, 2):
        full = PKL.bin_fields(dk, mi, N, axis)
        acc = None
        for (a, b) in ((0, 11), (11, 27), (27, N)):
            part = PKL.bin_fields([t[:, a:b, :].contiguous() for t in dk], mi, N, axis, ky_lo=a, nky=b - a)
            acc = part if acc is None else {k: acc[k] + part[k] for k in acc}
        for k in ("Nm3D", "Nm1D", "Nm2D", "k1D"):
            assert np.array_equal(acc[k], full[k]), k
        for k in ("k3D", "Pk3D", "PkX3D", "Pk1D", "PkX1D", "Pk2D", "PkX2D"):
            scale = float(np.max(np.abs(full[k]))) or 1.0
            assert np.max(np.abs(acc[k] - full[k])) < 1e-10 * scale, k


@pytest.mark.parametrize("N,parts", [(40, 3), (33, 4), (64, 8), (18, 10)])
def test_mirrored_slabs_add_up_to_the_full_binning(env, oracle, N, parts):
    """The mirrored-ky slab form of the multi-GPU path (pyl_pk_bin_mirrored): every rank holds |ky| rows and
    their mirrors; partial results sum to the single-GPU result -- counts exactly, sums to float64 round-off,
    phase included -- for every line of sight, even/odd N, and slabs as thin as one row."""
    torch, MASL, PKL, _ = env
    from pylians3_b200.dist import mirrored_rows, split_sizes
    fields, mas = make_fields(oracle, N, 1, 91 + N)
    dk = [PKL.fft3d_r2c_device(torch.from_numpy(f).cuda()) for f in fields]
    mi = [PKL.MAS_function(m) for m in mas]
    sizes, offs = split_sizes(N // 2 + 1, parts)
    for axis in (0, 1, 2):
        full = PKL.bin_fields(dk, mi, N, axis, want_phase=True)
        acc = None
        for r in range(parts):
            rows = torch.tensor(mirrored_rows(N, offs[r], sizes[r]), device="cuda")
            part = PKL.bin_fields([t[:, rows, :].contiguous() for t in dk], mi, N, axis, want_phase=True,
                                  ky_lo=offs[r], nky=sizes[r], mirrored=True)
            acc = part if acc is None else {k: acc[k] + part[k] for k in acc}
        for k in ("Nm3D", "Nm1D", "Nm2D", "k1D"):
            assert np.array_equal(acc[k], full[k]), (axis, k)
        for k in ("k3D", "Pk3D", "Pk1D", "Pk2D", "phase"):
            scale = float(np.max(np.abs(full[k]))) or 1.0
            assert np.max(np.abs(acc[k] - full[k])) < 1e-10 * scale, (axis, k)


@pytest.mark.parametrize("N,F", [(48, 1), (33, 3), (16, 2)])
def test_device_finalisation_equals_host_finalisation(env, oracle, N, F):
    """pyl_pk_finalize (units, averages, (2l+1), 1D area weight on the device) against the vectorised host
    restatement of Pk_library.pyx:384-418 / :735-791 on the same accumulators: same IEEE operations in the
    same order -> 1e-15, same NaN pattern (empty 2D bins are 0/0 in the reference too)."""
    torch, MASL, PKL, _ = env
    fields, mas = make_fields(oracle, N, F, 300 + N)
    dk = [PKL.fft3d_r2c_device(torch.from_numpy(f).cuda()) for f in fields]
    mi = [PKL.MAS_function(m) for m in mas]
    for axis in (0, 2):
        # ONE run of the bin kernel (its float64 reductions are not ordered), finalised both ways
        out, lay = PKL.bin_device(dk, mi, N, axis, F == 1)
        host = PKL._finalize(PKL.unpack_raw(out.cpu().numpy(), lay), BOX, N)
        dev = PKL.finalize_device(out, lay, BOX, N, k2d_on_device=True)
        assert set(host) == set(dev)
        for k in host:
            a, b = np.asarray(dev[k], dtype=np.float64), np.asarray(host[k], dtype=np.float64)
            assert a.shape == b.shape, k
            assert np.array_equal(np.isnan(a), np.isnan(b)), k
            ok = np.isnan(b) | (np.abs(a - b) <= 1e-15 * np.abs(b))
            assert np.all(ok), (k, float(np.nanmax(np.abs(a - b))))


def test_pk_full_size_properties(env):
    """BASELINE config 2 grid (512^3): exact mode count (Pk_library.pyx:87-99), exact scaling
    Pk(2*delta) = 4*Pk(delta), and shot noise of a uniform-random NGP field ~ L^3/Np."""
    torch, MASL, PKL, _ = env
    N = 512
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    pos = torch.rand((N ** 3 // 8, 3), generator=g, device="cuda", dtype=torch.float32) * BOX
    grid = torch.zeros((N, N, N), dtype=torch.float32, device="cuda")
    MASL.MA(pos, grid, BOX, "CIC")
    from pylians3_b200 import overdensity_
    overdensity_(grid)
    pk = PKL.Pk(grid, BOX, 0, "CIC", verbose=False)
    own = 8
    assert int(pk.Nmodes3D.sum()) + 1 == (N ** 3 - own) // 2 + own       # bin 0 (DC) was dropped
    assert pk.Pk.shape == (443, 3) and pk.Nmodes1D.shape == (256,) and pk.Pk2D.shape == (257 * 363,)
    shot = BOX ** 3 / (N ** 3 // 8)
    sel = (pk.k3D > 0.05) & (pk.k3D < 1.0)
    assert abs(np.mean(pk.Pk[sel, 0]) / shot - 1.0) < 0.02
    grid.mul_(2.0)
    pk2 = PKL.Pk(grid, BOX, 0, "CIC", verbose=False)
    assert rel_err(pk2.Pk[:, 0], 4.0 * pk.Pk[:, 0], 1e-300) < 1e-12
    assert rel_err(pk2.Pk2D, 4.0 * pk.Pk2D, 1e-300) < 1e-12


def test_full_size_512_against_the_compiled_reference(env):
    """one 512^3 CIC + Pk run of BASELINE config 2's shape against the COMPILED, UNMODIFIED
    reference (oracle/_ref) on identical host arrays: deposited grid 1e-5 per cell, mode counts bit-exact, spectra
    inside the north_star bars.  About a minute of host time for the reference's serial loops."""
    torch, MASL, PKL, _ = env
    from oracle import ref_loader
    if not ref_loader.have_ref():
        pytest.skip("oracle/_ref (the compiled reference) is not present")
    RM, RP = ref_loader.ref_MASL(), ref_loader.ref_PKL()
    N = 512
    pos = np.random.default_rng(2).random((N ** 3, 3), dtype=np.float32) * np.float32(BOX)
    ref = np.zeros((N, N, N), np.float32)
    RM.MA(pos, ref, BOX, "CIC")
    grid = torch.zeros((N, N, N), dtype=torch.float32, device="cuda")
    MASL.MA(torch.from_numpy(pos).cuda(), grid, BOX, "CIC")
    got = grid.cpu().numpy()
    assert abs(float(got.sum(dtype=np.float64)) / N ** 3 - 1.0) < 1e-5          # library/tests/test.py:30
    assert rel_err(got, ref, floor=float(ref.mean())) < 1e-5
    del pos, got
    # both sides transform the SAME delta (the reference's grid), so only FFT + binning are compared here
    ref /= np.mean(ref, dtype=np.float64)
    ref -= 1.0
    want = quiet(RP.Pk, ref, BOX, 0, "CIC", 8, False)
    have = PKL.Pk(torch.from_numpy(ref).cuda(), BOX, 0, "CIC", verbose=False)
    check_pk(have, want, phase_min_modes=64)
    # and the fused form of the bench step: the deposit starts from -c (prebias_), the transform sees n - c, and
    # n/<n> - 1 is folded into the scale of the binned sums.  (Without the offset the float32 transform of a grid
    # whose mass sits in the DC mode misses the 1e-4 bar in the single-mode Pk2D bins on the axis through k = 0.)
    from pylians3_b200 import prebias_
    pos_d = torch.from_numpy(np.random.default_rng(2).random((N ** 3, 3), dtype=np.float32) * np.float32(BOX)).cuda()
    c = prebias_(grid, N ** 3)
    assert float(c) == 1.0
    MASL.MA(pos_d, grid, BOX, "CIC")
    fused = PKL.Pk(grid, BOX, 0, "CIC", verbose=False, density=True, offset=c)
    check_pk(fused, want, phase_min_modes=64)


def make_densities(oracle, N, nfields, seed):
    """The fields of make_fields BEFORE `g /= mean; g -= 1` (densities with different means), and after."""
    pos, W = make_particles(seed, 3 * N ** 3, True)
    specs = [("PCS", None), ("CIC", W), ("NGP", None), ("TSC", W * W), ("CIC", None), ("PCS", W)]
    dens, delta, mas = [], [], []
    for m, w in specs[:nfields]:
        g = np.zeros((N, N, N), np.float32)
        oracle.MA(pos, g, BOX, m, w)
        dens.append(g.copy())
        g /= np.mean(g, dtype=np.float64)
        g -= 1.0
        delta.append(g)
        mas.append(m)
    return dens, delta, mas


@pytest.mark.parametrize("N,axis", [(64, 0), (48, 2), (33, 1), (128, 0)])
def test_pk_of_a_density_equals_pk_of_its_overdensity(env, oracle, N, axis):
    """density=True (pyl_pk_take_dc + pyl_pk_density_scale): the caller's `delta /= mean; delta -= 1`
    (Pk_snapshot.py:88-89) folded into the spectrum.  Checked against the oracle's Pk of the normalised field."""
    torch, MASL, PKL, _ = env
    dens, delta, mas = make_densities(oracle, N, 2, 7 * N + axis)
    for fi in (0, 1):
        ref = oracle.Pk(delta[fi], BOX, axis, mas[fi], 1, False)
        d = torch.from_numpy(dens[fi]).cuda()
        with pytest.raises(ValueError):
            PKL.Pk(d, BOX, axis, mas[fi], verbose=False, density=True)      # the offset has to be stated
        got = PKL.Pk(d, BOX, axis, mas[fi], verbose=False, density=True, offset=0.0)
        check_pk(got, ref, phase_min_modes=64)       # (a different field goes through the float32 transform)
        assert torch.equal(d.cpu(), torch.from_numpy(dens[fi]))            # the density is not modified
        assert got.Pk2D[0] == 0.0                                          # the DC mode is dropped, not binned
        # the grid may hold n - c for any constant c (the deposit started from -c): <n> = c + DC/dims^3
        c = float(np.float32(0.97 * dens[fi].mean()))
        got = PKL.Pk(d - c, BOX, axis, mas[fi], verbose=False, density=True, offset=c)
        check_pk(got, ref, phase_min_modes=64)       # (a different field goes through the float32 transform)


def test_prebias_then_deposit_then_density_spectrum(env, oracle):
    """The bench step's recipe end to end at a small size: prebias_ (fills -c, c from a sample of the weights,
    no host sync) -> MA -> Pk(density=True, offset=c) against oracle MA -> delta -> Pk."""
    torch, MASL, PKL, _ = env
    from pylians3_b200 import prebias_
    N = 96
    pos, W = make_particles(41, 3 * N ** 3, True)
    ref = np.zeros((N, N, N), np.float32)
    oracle.MA(pos, ref, BOX, "PCS", W)
    dens = ref.copy()
    ref /= np.mean(ref, dtype=np.float64)
    ref -= 1.0
    want = oracle.Pk(ref, BOX, 0, "PCS", 1, False)
    pos_d, W_d = torch.from_numpy(pos).cuda(), torch.from_numpy(W).cuda()
    for w_arg in (W_d, W):                              # weights on the device, or still on the host
        grid = torch.empty((N, N, N), dtype=torch.float32, device="cuda")
        c = prebias_(grid, len(pos), w_arg)
        assert abs(float(c) / float(dens.mean(dtype=np.float64)) - 1.0) < 1e-6      # the exact sum of the weights
        assert torch.all(grid == -c.float())
        MASL.MA(pos_d, grid, BOX, "PCS", W_d)
        got_dens = (grid.double() + c).float().cpu().numpy()
        assert rel_err(got_dens, dens, floor=float(dens.mean())) < 1e-5
        fused = PKL.Pk(grid, BOX, 0, "PCS", verbose=False, density=True, offset=c)
        check_pk(fused, want, phase_min_modes=1 << 30)              # Pk, Pk1D, Pk2D, counts against the oracle's chain
        # the phase statistic hangs on modes at the float32 noise floor: the two DEPOSITS (accumulation order) move
        # it by 2e-4 in one shell whatever transforms them, so it is held to the oracle's Pk of THIS deposit
        got_dens /= np.mean(got_dens, dtype=np.float64)
        got_dens -= 1.0
        check_pk(fused, oracle.Pk(got_dens, BOX, 0, "PCS", 1, False), phase_min_modes=64)

@pytest.mark.parametrize("N,F", [(48, 2), (40, 3), (32, 4), (24, 6)])
def test_xpk_of_densities(env, oracle, N, F):
    torch, MASL, PKL, _ = env
    dens, delta, mas = make_densities(oracle, N, F, 300 + N + F)
    ref = quiet(oracle.XPk, delta, BOX, 2, mas, 1)
    got = quiet(PKL.XPk, dens, BOX, 2, mas, 1, density=True, offset=0.0)
    check_pk(got, ref, cross=True)
    cs = [float(np.float32(1.02 * d.mean())) for d in dens]
    got = quiet(PKL.XPk, [d - np.float32(c) for d, c in zip(dens, cs)], BOX, 2, mas, 1, density=True, offset=cs)
    check_pk(got, ref, cross=True)
