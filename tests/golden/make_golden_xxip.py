#!/usr/bin/env python
"""Golden vectors of XXi_projected (Pk_library.pyx:2684-2789) from the COMPILED, UNMODIFIED reference (oracle/_ref).

    python tests/golden/make_golden_xxip.py        # where oracle/_ref has been built (oracle/build_ref.py)

Inputs are the seeded images of make_golden_more.inputs(N).  Output: tests/golden/xxi_projected_golden.npz."""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_more import BOX, ROOT, inputs  # noqa: E402

sys.path.insert(0, ROOT)
SIZES = (12, 9, 32)


def main():
    from oracle import ref_loader
    P = ref_loader.ref_PKL()
    out = {}
    for N in SIZES:
        I = inputs(N)
        with contextlib.redirect_stdout(io.StringIO()):
            r = P.XXi_projected(I["img1"], I["img2"], BOX, ["CIC", "PCS"], 1)
        out.update({"N%d_r" % N: np.asarray(r.r_p), "N%d_xi" % N: np.asarray(r.xi_p), "N%d_Nm" % N: np.asarray(r.Nmodes_p)})
    np.savez_compressed(os.path.join(HERE, "xxi_projected_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
