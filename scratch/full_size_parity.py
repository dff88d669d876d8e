"""BASELINE config 3 at FULL size against the compiled, unmodified reference (oracle/_ref) on identical inputs:
1024^3 Zel'dovich particles + W -> 1024^3 grid, PCS, delta, Pk(axis 0).  One-off measurement (several minutes of
host time for the reference's serial loops), result -> gpurun_out/full_size_parity_config3.json.

    python scratch/full_size_parity.py [grid=1024]
"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import bench  # noqa: E402
from oracle import ref_loader  # noqa: E402
from pylians3_b200 import MAS_library as MASL, Pk_library as PKL, prebias_  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
BOX = bench.BOX
dev = torch.device("cuda", 0)
wl = bench.WORKLOADS["config3"]
pos, W = bench.make_inputs(wl, N, 1, 0, dev)
grid = torch.empty((N, N, N), dtype=torch.float32, device=dev)
c = prebias_(grid, pos.shape[0], W)
MASL.MA(pos, grid, BOX, "PCS", W)
pk = PKL.Pk(grid, BOX, 0, "PCS", verbose=False, density=True, offset=c)
dens_gpu = (grid.double() + c).float().cpu().numpy()
del grid
pos_h, W_h = pos.cpu().numpy(), W.cpu().numpy()
del pos, W
torch.cuda.empty_cache()

RM, RP = ref_loader.ref_MASL(), ref_loader.ref_PKL()
out = {"workload": "config 3 at full size: %d^3 Zel'dovich particles + W -> %d^3 grid, PCS, Pk(axis=0)" % (N, N)}
t = time.time()
ref = np.zeros((N, N, N), np.float32)
RM.MA(pos_h, ref, BOX, "PCS", W_h)
out["reference_MA_seconds"] = time.time() - t
mean = float(ref.mean(dtype=np.float64))
err = np.abs(dens_gpu - ref)
out["deposit_max_abs_err_over_mean"] = float(err.max() / mean)
out["deposit_max_rel_err_cells_above_mean"] = float((err / np.maximum(np.abs(ref), mean)).max())
out["mass_rel_diff"] = abs(float(dens_gpu.sum(dtype=np.float64)) / float(ref.sum(dtype=np.float64)) - 1.0)
del dens_gpu, err
ref /= np.mean(ref, dtype=np.float64)
ref -= 1.0
t = time.time()
import contextlib, io
with contextlib.redirect_stdout(io.StringIO()):
    want = RP.Pk(ref, BOX, 0, "PCS", 16, False)
out["reference_Pk_seconds"] = time.time() - t
for nm in ("Nmodes3D", "Nmodes1D", "Nmodes2D"):
    out[nm + "_equal"] = bool(np.array_equal(np.asarray(getattr(pk, nm)), np.asarray(getattr(want, nm))))
for nm in ("k3D", "k1D", "kpar", "kper"):
    a, b = np.asarray(getattr(pk, nm)), np.asarray(getattr(want, nm))
    out[nm + "_max_rel_err"] = float(np.nanmax(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
P, Q = np.asarray(pk.Pk), np.asarray(want.Pk)
out["Pk0_max_rel_err"] = float(np.max(np.abs(P[:, 0] - Q[:, 0]) / np.abs(Q[:, 0])))
out["Pk0_bins_outside_1e-4"] = int(np.sum(np.abs(P[:, 0] - Q[:, 0]) > 1e-4 * np.abs(Q[:, 0])))
out["Pk0_bins"] = int(len(Q))
for l, name in ((1, "Pk2"), (2, "Pk4")):
    out[name + "_max_err_over_P0"] = float(np.max(np.abs(P[:, l] - Q[:, l]) / np.abs(Q[:, 0])))
a, b = np.asarray(pk.Pk1D), np.asarray(want.Pk1D)
out["Pk1D_max_rel_err"] = float(np.nanmax(np.abs(a - b) / np.abs(b)))
a, b = np.asarray(pk.Pk2D), np.asarray(want.Pk2D)
ok = ~np.isnan(b)
ok[0] = False                                  # the DC slot: 0 here, squared rounding residue in the reference
r = np.abs(a[ok] - b[ok]) / np.abs(b[ok])
out["Pk2D_max_rel_err"] = float(r.max())
out["Pk2D_bins_outside_1e-4"] = int(np.sum(r > 1e-4))
out["Pk2D_bins"] = int(ok.sum())
# the test suite's bars (tests/test_gpu_pk.py: counts exact, k 1e-12, spectra 1e-4 relative + the float32 FFT floor of
# low-power bins).  The k3D bar is a small-grid bar: at 1024^3 a shell sums up to 3e6 values of |k| and the two
# summation orders (serial in the reference, tree here) differ by a few 1e-12; it is reported and then set aside so
# that the remaining bars are evaluated too.
from test_gpu_pk import check_pk
notes = []
for attempt in range(3):
    try:
        check_pk(pk, want, phase_min_modes=64)
        notes.append("passed")
        break
    except AssertionError as e:
        notes.append("failed: %s" % e)
        if str(e).startswith("k3D") and attempt == 0:
            pk.k3D = np.asarray(want.k3D)
            continue
        break
out["check_pk"] = notes
sel = np.asarray(want.Nmodes3D) >= 64
out["Pkphase_max_rel_err_shells_ge_64_modes"] = float(np.max(np.abs(np.asarray(pk.Pkphase)[sel] / np.asarray(want.Pkphase)[sel] - 1)))
print(json.dumps(out, indent=1))
open("gpurun_out/full_size_parity_config3.json", "w").write(json.dumps(out, indent=1) + "\n")
