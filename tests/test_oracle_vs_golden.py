"""CPU: pin the oracle (oracle/*.c restatement) against the golden vectors produced by the
compiled, unmodified reference (tests/golden/make_golden.py)."""
import contextlib
import io

import numpy as np
import pytest

from conftest import BOX, rel_err

MAS = ("NGP", "CIC", "TSC", "PCS")


@pytest.mark.parametrize("N", [16, 9])
@pytest.mark.parametrize("clu", ["uni", "clu"])
@pytest.mark.parametrize("mas", MAS)
@pytest.mark.parametrize("weighted", [False, True])
def test_ma_oracle_matches_reference(oracle, ma_golden, N, clu, mas, weighted):
    tag = "N%d_%s" % (N, clu)
    pos, W = ma_golden[tag + "_pos"], ma_golden[tag + "_W"]
    w = W if weighted else None
    g = np.zeros((N, N, N), np.float32)
    oracle.MA(pos, g, BOX, mas, w)
    ref = ma_golden["%s_%s_%s_3D" % (tag, mas, "W" if weighted else "U")]
    g2 = np.zeros((N, N), np.float32)
    oracle.MA(np.ascontiguousarray(pos[:, :2]), g2, BOX, mas, w)
    ref2 = ma_golden["%s_%s_%s_2D" % (tag, mas, "W" if weighted else "U")]
    if not weighted:
        # same operations in the same (serial) order -> bit-exact
        assert np.array_equal(g, ref)
        assert np.array_equal(g2, ref2)
    else:
        # the reference binary fuses w*W + number into one FMA; allow float32 round-off
        assert rel_err(g, ref, floor=float(ref.mean())) < 2e-6
        assert rel_err(g2, ref2, floor=float(ref2.mean())) < 2e-6


def test_ma_oracle_accumulates(oracle, ma_golden):
    pos, W = ma_golden["accum_pos"], ma_golden["accum_W"]
    g = np.full((12, 12, 12), 0.25, np.float32)
    oracle.MA(pos[:1500], g, BOX, "TSC", W[:1500])
    oracle.MA(pos[1500:], g, BOX, "TSC", W[1500:])
    assert rel_err(g, ma_golden["accum_TSC_W_3D"], floor=0.25) < 2e-6
    g2 = np.full((12, 12), 0.25, np.float32)
    oracle.MA(np.ascontiguousarray(pos[:, :2]), g2, BOX, "PCS", None, False, False)
    assert np.array_equal(g2, ma_golden["accum_PCS_U_2D_norenorm"])


@pytest.mark.parametrize("N", [16, 9])
def test_cic_interp_oracle_matches_reference(oracle, ma_golden, N):
    """MAS_library.pyx:558-599: the reference is compiled with -ffast-math (free to reassociate the 8-term sum),
    so float32 round-off of the sum is allowed: 1e-6 of the field's scale."""
    pos, field, ref = (ma_golden["interp_N%d_%s" % (N, k)] for k in ("pos", "field", "den"))
    den = np.full(len(pos), 7.0, np.float32)          # overwritten, not accumulated
    oracle.CIC_interp(field, BOX, pos, den)
    assert rel_err(den, ref, floor=float(np.abs(field).mean())) < 1e-6


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_redshift_space_oracle_matches_reference(oracle, ma_golden, axis):
    """redshift_space_library.pyx:29-46, bit-exact (fused multiply-add like the reference binary), including
    velocities that cross the box several times."""
    pos = ma_golden["rsd_pos"].copy()
    oracle.pos_redshift_space(pos, ma_golden["rsd_vel"], BOX, 171.5, 0.5, axis)
    assert np.array_equal(pos, ma_golden["rsd_out_a%d" % axis])


PK_ATTRS = ("k3D", "Pk", "Nmodes3D", "Pkphase", "k1D", "Pk1D", "Nmodes1D", "kpar", "kper", "Pk2D", "Nmodes2D")
XPK_ATTRS = ("k3D", "Pk", "XPk", "Nmodes3D", "k1D", "Pk1D", "PkX1D", "Nmodes1D", "kpar", "kper", "Pk2D",
             "PkX2D", "Nmodes2D")


@pytest.mark.parametrize("N", [16, 15])
@pytest.mark.parametrize("axis", [0, 1, 2])
def test_pk_oracle_matches_reference(oracle, pk_golden, N, axis):
    delta = pk_golden["N%d_delta" % N]
    for fi, mas in ((0, "PCS"), (1, "CIC"), (2, "NGP"), (0, "TSC"), (0, None)):
        pk = oracle.Pk(delta[fi], BOX, axis, mas, 1, False)
        t = "N%d_Pk_a%d_f%d_%s_" % (N, axis, fi, mas)
        for nm in PK_ATTRS:
            ref = pk_golden[t + nm]
            got = np.asarray(getattr(pk, nm))
            if nm.startswith("Nmodes"):
                assert np.array_equal(got, ref), nm
            else:
                # identical algorithm and FFT; only the association of a few float64 products differs
                assert rel_err(got, ref, floor=1e-300) < 1e-9 or _small(got, ref, pk_golden[t + "Pk"]), nm


def _small(got, ref, pk):
    # multipoles that are pure round-off (corner-mode quadrupole, SURVEY 8a note ii)
    return float(np.nanmax(np.abs(np.asarray(got) - np.asarray(ref)))) < 1e-9 * float(np.nanmax(np.abs(pk)))


@pytest.mark.parametrize("N", [16, 15])
@pytest.mark.parametrize("axis", [0, 1, 2])
def test_xpk_oracle_matches_reference(oracle, pk_golden, N, axis):
    delta = pk_golden["N%d_delta" % N]
    with contextlib.redirect_stdout(io.StringIO()):
        x = oracle.XPk([delta[0], delta[1], delta[2]], BOX, axis, ["PCS", "CIC", "None"], 1)
    t = "N%d_XPk_a%d_" % (N, axis)
    for nm in XPK_ATTRS:
        ref = pk_golden[t + nm]
        got = np.asarray(getattr(x, nm))
        if nm.startswith("Nmodes"):
            assert np.array_equal(got, ref), nm
        else:
            assert rel_err(got, ref, floor=1e-300) < 1e-9 or _small(got, ref, pk_golden[t + "Pk"]), nm


def test_mode_count_invariant(oracle):
    # Pk_library.pyx:87-99
    for N in (8, 9, 16, 15):
        d = np.random.default_rng(N).standard_normal((N, N, N)).astype(np.float32)
        r = oracle.bin_raw([oracle.fft3d_r2c(d)], N, [0], 2, BOX)
        assert int(r["Nm3D"].sum()) == oracle.expected_modes(N)
