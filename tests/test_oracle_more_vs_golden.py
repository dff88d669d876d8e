"""CPU: pin the oracle's restatement of the sibling estimators (oracle/cpu_more.py over pk_oracle.c) against
tests/golden/pk_more_golden.npz, the outputs of the compiled, unmodified reference."""
import os

import numpy as np
import pytest

import more_cases as MC
from conftest import GOLDEN


@pytest.fixture(scope="module")
def golden():
    return dict(np.load(os.path.join(GOLDEN, "pk_more_golden.npz")))


@pytest.mark.parametrize("N", MC.SIZES)
def test_sibling_oracle_matches_reference(oracle, golden, N):
    from oracle import cpu_more
    got = MC.run_all(cpu_more, N)
    assert set(got) == {k for k in golden if k.startswith("N%d_" % N) and "_sm_" not in k and "_xm_" not in k}
    # same transform (pocketfft) on both sides: only float64 association / float32 FMA contraction differ
    bad = MC.compare(got, golden, tol=2e-6)
    assert not bad, bad


@pytest.mark.parametrize("N", MC.SIZES)
def test_smoothing_oracle_matches_reference(oracle, golden, N):
    """smoothing_library: FT_filter / FT_filter_2D for the three filters, field_smoothing and its 2D form."""
    from oracle import cpu_more
    got = MC.run_smoothing(cpu_more, N)
    assert set(got) == {k for k in golden if k.startswith("N%d_sm_" % N)}
    bad = MC.compare_smoothing(got, golden, tol=1e-6)
    assert not bad, bad


@pytest.mark.parametrize("N", MC.SIZES)
def test_xxi_multi_oracle_matches_reference(oracle, golden, N):
    from oracle import cpu_more
    # XXi_multi forms re*re + im*im in double and rounds once (`np.float64_t[::1] real_part`, :2552-2555); the
    # restatement reuses Xi's float32 mode loop: one float32 ulp per mode
    bad = MC.compare_xxi_multi(MC.run_xxi_multi(cpu_more, N), golden, 1e-5, N)
    assert not bad, bad


def test_inverse_transform_is_normalised(oracle, golden):
    """correct_MAS with MAS=None is FFT -> IFFT: the reference returns the field itself, which pins the
    normalised inverse the oracle (and the product) must use (see oracle/pyfftw_shim)."""
    from oracle import cpu_more
    I = MC.inputs(12)
    back = cpu_more.correct_MAS(I["d1"].copy(), MC.BOX, "None", 1)
    assert np.max(np.abs(back - I["d1"])) < 1e-5 * np.abs(I["d1"]).max()
    # and the reference's deconvolved field is of the same order as the input, not dims^3 larger
    assert np.abs(golden["N12_cmas"]).max() < 10 * np.abs(I["d1"]).max()


def test_xxi_projected_restatement_matches_the_compiled_reference():
    """XXi_projected (Pk_library.pyx:2684-2789): counts and r_p exact, xi_p to float32 round-off of its peak (the
    reference's float32 products may be fused by its compiler)."""
    import contextlib
    import io
    import os

    import numpy as np
    from conftest import BOX, GOLDEN
    from golden.make_golden_more import inputs
    from oracle import cpu_more as OM
    g = np.load(os.path.join(GOLDEN, "xxi_projected_golden.npz"))
    for N in (12, 9, 32):
        I = inputs(N)
        with contextlib.redirect_stdout(io.StringIO()):
            r = OM.XXi_projected(I["img1"], I["img2"], BOX, ["CIC", "PCS"], 1)
        assert np.array_equal(r.Nmodes_p, g["N%d_Nm" % N])
        assert np.max(np.abs(r.r_p - g["N%d_r" % N]) / g["N%d_r" % N]) < 1e-12
        assert np.max(np.abs(r.xi_p - g["N%d_xi" % N])) < 2e-6 * np.max(np.abs(g["N%d_xi" % N]))
