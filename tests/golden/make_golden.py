#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the COMPILED, UNMODIFIED reference (oracle/_ref).

Run in the authoring container only (needs /root/reference to have been built by
oracle/build_ref.py):   python tests/golden/make_golden.py
The fixtures are small on purpose (a few hundred kB) and are committed; the GPU box has no
/root/reference, so the -m gpu tests and the oracle tests read these files instead.

Reference entry points exercised:
  MAS_library.MA            library/MAS_library/MAS_library.pyx:57-112
  MAS_library.CIC_interp    library/MAS_library/MAS_library.pyx:558-599
  redshift_space_library.pos_redshift_space  library/redshift_space_library/redshift_space_library.pyx:29-46
  Pk_library.Pk             library/Pk_library/Pk_library.pyx:263-420
  Pk_library.XPk            library/Pk_library/Pk_library.pyx:529-793
(the FFT inside Pk/XPk goes through oracle/pyfftw_shim -> scipy pocketfft, float32)
"""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
BOX = 1000.0
EDGE = np.array([[0, 0, 0], [1000, 1000, 1000], [999.99994, 0.5, 500], [1000, 0, 999.99994],
                 [31.25, 31.249998, 15.625], [62.5, 93.75, 968.75], [499.99997, 500.00003, 0.0001]],
                dtype=np.float32)


def particles(seed, n, clustered=False):
    rng = np.random.default_rng(seed)
    pos = rng.random((n, 3), dtype=np.float32) * np.float32(BOX)
    if clustered:  # a few tight clumps -> many particles per cell, as after Zel'dovich collapse
        c = rng.random((8, 3), dtype=np.float32) * np.float32(BOX)
        k = n // 2
        pos[:k] = (c[rng.integers(0, 8, k)] + rng.normal(0, 12.0, (k, 3)).astype(np.float32)) % np.float32(BOX)
    pos[:len(EDGE)] = EDGE
    W = rng.random(n, dtype=np.float32)
    return pos, W


def main():
    M = ref_loader.ref_MASL()
    P = ref_loader.ref_PKL()

    # ---- MA: every scheme x (weights) x (3D, 2D) x (even, odd grid) x (uniform, clustered)
    out = {}
    for N in (16, 9):
        for clustered in (False, True):
            tag0 = "N%d_%s" % (N, "clu" if clustered else "uni")
            pos, W = particles(100 + N + int(clustered), 6000, clustered)
            out[tag0 + "_pos"] = pos
            out[tag0 + "_W"] = W
            for mas in ("NGP", "CIC", "TSC", "PCS"):
                for w in (None, W):
                    wt = "W" if w is not None else "U"
                    g = np.zeros((N, N, N), np.float32)
                    M.MA(pos, g, BOX, mas, w)
                    out["%s_%s_%s_3D" % (tag0, mas, wt)] = g
                    g2 = np.zeros((N, N), np.float32)
                    M.MA(np.ascontiguousarray(pos[:, :2]), g2, BOX, mas, w)
                    out["%s_%s_%s_2D" % (tag0, mas, wt)] = g2
    # accumulation across calls into a non-empty grid (MAS_gadget.py:63-75 relies on it)
    pos, W = particles(7, 3000)
    g = np.full((12, 12, 12), 0.25, np.float32)
    M.MA(pos[:1500], g, BOX, "TSC", W[:1500]); M.MA(pos[1500:], g, BOX, "TSC", W[1500:])
    out["accum_pos"], out["accum_W"], out["accum_TSC_W_3D"] = pos, W, g
    g2 = np.full((12, 12), 0.25, np.float32)          # 2D without the final renormalisation
    M.MA(np.ascontiguousarray(pos[:, :2]), g2, BOX, "PCS", None, False, False)
    out["accum_PCS_U_2D_norenorm"] = g2
    # CIC_interp (grid -> particles), MAS_library.pyx:558-599: a signed field, edge positions included
    for N in (16, 9):
        pos, _ = particles(300 + N, 5000, clustered=True)
        field = np.random.default_rng(400 + N).standard_normal((N, N, N)).astype(np.float32)
        den = np.zeros(len(pos), np.float32)
        M.CIC_interp(field, BOX, pos, den)
        out["interp_N%d_pos" % N], out["interp_N%d_field" % N], out["interp_N%d_den" % N] = pos, field, den
    # pos_redshift_space (redshift_space_library.pyx:29-46): ordinary velocities plus a few that cross the box
    # several times in both directions (the "neutrino" case the reference comments on)
    R = ref_loader.ref_RSL()
    rng = np.random.default_rng(555)
    pos = rng.random((4000, 3), dtype=np.float32) * np.float32(BOX)
    pos[:len(EDGE)] = EDGE
    vel = (rng.standard_normal((4000, 3)) * 2500.0).astype(np.float32)
    vel[:40] *= 400.0
    out["rsd_pos"], out["rsd_vel"] = pos, vel
    for axis in (0, 1, 2):
        p = pos.copy()
        R.pos_redshift_space(p, vel, BOX, 171.5, 0.5, axis)
        out["rsd_out_a%d" % axis] = p
    np.savez_compressed(os.path.join(HERE, "ma_golden.npz"), **out)

    # ---- Pk / XPk
    out = {}
    sink = io.StringIO()
    for N in (16, 15):
        pos, W = particles(200 + N, 4 * N ** 3, clustered=True)
        fields = []
        for mas, w in (("PCS", None), ("CIC", W), ("NGP", None)):
            g = np.zeros((N, N, N), np.float32)
            M.MA(pos, g, BOX, mas, w)
            g /= np.mean(g, dtype=np.float64); g -= 1.0
            fields.append(g)
        out["N%d_delta" % N] = np.stack(fields)
        for axis in (0, 1, 2):
            for fi, mas in ((0, "PCS"), (1, "CIC"), (2, "NGP"), (0, "TSC"), (0, None)):
                pk = P.Pk(fields[fi], BOX, axis, mas, 1, False)
                t = "N%d_Pk_a%d_f%d_%s_" % (N, axis, fi, mas)
                for nm in ("k3D", "Pk", "Nmodes3D", "Pkphase", "k1D", "Pk1D", "Nmodes1D",
                           "kpar", "kper", "Pk2D", "Nmodes2D"):
                    out[t + nm] = np.asarray(getattr(pk, nm))
            with contextlib.redirect_stdout(sink):
                x = P.XPk(fields, BOX, axis, ["PCS", "CIC", "None"], 1)
            t = "N%d_XPk_a%d_" % (N, axis)
            for nm in ("k3D", "Pk", "XPk", "Nmodes3D", "k1D", "Pk1D", "PkX1D", "Nmodes1D",
                       "kpar", "kper", "Pk2D", "PkX2D", "Nmodes2D"):
                out[t + nm] = np.asarray(getattr(x, nm))
    np.savez_compressed(os.path.join(HERE, "pk_golden.npz"), **out)
    for f in ("ma_golden.npz", "pk_golden.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "kB")


if __name__ == "__main__":
    main()
