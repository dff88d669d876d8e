"""CPU: when the compiled reference is present (oracle/_ref, authoring container), compare the
oracle restatement with it live on fresh random inputs.  Skipped on boxes without it."""
import contextlib
import io

import numpy as np
import pytest

from conftest import BOX, make_particles, rel_err

ref_loader = pytest.importorskip("oracle.ref_loader")
pytestmark = pytest.mark.skipif(not ref_loader.have_ref(), reason="oracle/_ref not built here")


@pytest.mark.parametrize("N,clustered", [(32, False), (33, True)])
def test_ma_live(oracle, N, clustered):
    M = ref_loader.ref_MASL()
    pos, W = make_particles(N, 2 * N ** 3, clustered)
    for mas in ("NGP", "CIC", "TSC", "PCS"):
        for w in (None, W):
            a = np.zeros((N, N, N), np.float32)
            b = a.copy()
            M.MA(pos, a, BOX, mas, w)
            oracle.MA(pos, b, BOX, mas, w)
            if w is None:
                assert np.array_equal(a, b), mas
            else:
                assert rel_err(b, a, floor=float(a.mean())) < 2e-6, mas


@pytest.mark.parametrize("N", [24, 21])
def test_pk_xpk_live(oracle, N):
    M, P = ref_loader.ref_MASL(), ref_loader.ref_PKL()
    pos, W = make_particles(N + 1, 3 * N ** 3, True)
    fields = []
    for mas, w in (("TSC", None), ("PCS", W)):
        g = np.zeros((N, N, N), np.float32)
        M.MA(pos, g, BOX, mas, w)
        g /= np.mean(g, dtype=np.float64)
        g -= 1.0
        fields.append(g)
    for axis in (0, 1, 2):
        p = P.Pk(fields[0], BOX, axis, "TSC", 1, False)
        o = oracle.Pk(fields[0], BOX, axis, "TSC", 1, False)
        for nm in ("Nmodes3D", "Nmodes1D", "Nmodes2D"):
            assert np.array_equal(np.asarray(getattr(p, nm)), getattr(o, nm))
        for nm in ("k3D", "k1D", "Pk1D", "Pk2D", "Pkphase", "kpar", "kper"):
            assert rel_err(getattr(o, nm), np.asarray(getattr(p, nm)), 1e-300) < 1e-9, nm
        P0 = np.asarray(p.Pk)[:, 0:1]
        assert np.all(np.abs(o.Pk - np.asarray(p.Pk)) <= 1e-9 * np.abs(P0) * [1, 5, 9])
        with contextlib.redirect_stdout(io.StringIO()):
            xp = P.XPk(fields, BOX, axis, ["TSC", "PCS"], 1)
            xo = oracle.XPk(fields, BOX, axis, ["TSC", "PCS"], 1)
        assert rel_err(xo.XPk[:, 0], np.asarray(xp.XPk)[:, 0], 1e-300) < 1e-9
        assert rel_err(xo.PkX2D, np.asarray(xp.PkX2D), 1e-300) < 1e-9
