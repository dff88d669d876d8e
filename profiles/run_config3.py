#!/usr/bin/env python
"""BASELINE.json config 3 on one B200: 1024^3 Zel'dovich-displaced particles with weights -> 1024^3 grid with PCS,
then Pk (l = 0,2,4, 1D, 2D(kpar,kper)).  Prints stage times (CUDA events, median of `reps`), the size-independent
checks (mass conservation against the float64 sum of the weights, mode counts, exact scaling) and ONE JSON line.

    python profiles/run_config3.py [n_side=1024] [reps=3] [kind=zeldovich|uniform]
"""
import contextlib
import io
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from pylians3_b200 import MAS_library as MASL, Pk_library as PKL, synth, overdensity_  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
kind = sys.argv[3] if len(sys.argv) > 3 else "zeldovich"
BOX, MAS, AXIS = 1000.0, "PCS", 0
dev = torch.device("cuda", 0)
t0 = time.time()
pos = synth.zeldovich_device(n, BOX, 3, dev) if kind == "zeldovich" else synth.uniform_device(n ** 3, BOX, 3, dev)
W = synth.weights_device(n ** 3, 3, dev)
torch.cuda.synchronize()
gen_s = time.time() - t0
torch.cuda.empty_cache()
grid = torch.zeros((n, n, n), dtype=torch.float32, device=dev)


def ev():
    return torch.cuda.Event(enable_timing=True)


stages = {k: [] for k in ("zero", "deposit", "overdensity", "pk")}
pk = None
for r in range(reps + 1):
    e = [ev() for _ in range(5)]
    torch.cuda.synchronize()
    e[0].record(); grid.zero_()
    e[1].record(); MASL.MA(pos, grid, BOX, MAS, W)
    e[2].record()
    if r == 0:
        total = float(grid.sum(dtype=torch.float64))
        wsum = float(W.sum(dtype=torch.float64))
        e[2].record()
    overdensity_(grid)
    e[3].record()
    with contextlib.redirect_stdout(io.StringIO()):
        pk = PKL.Pk(grid, BOX, AXIS, MAS, verbose=False)
    e[4].record()
    torch.cuda.synchronize()
    if r > 0:                                   # first pass = warm-up (plans, workspaces)
        for i, k in enumerate(("zero", "deposit", "overdensity", "pk")):
            stages[k].append(e[i].elapsed_time(e[i + 1]))
med = {k: float(np.median(v)) for k, v in stages.items()}
step = sum(med.values())
m = n // 2
checks = {
    "mass_conservation_rel": abs(total / wsum - 1.0),
    "modes_counted": int(pk.Nmodes3D.sum()) + 1,
    "modes_expected": (n ** 3 - 8) // 2 + 8,
    "n2d_bins": int(pk.Pk2D.shape[0]),
    "n2d_expected": (m + 1) * (int(np.sqrt(2.0 * m * m)) + 1),
    "Nmodes2D_sum": int(pk.Nmodes2D.sum()),
    "Pk0_finite": bool(np.all(np.isfinite(pk.Pk[:, 0]))),
}
ok = checks["mass_conservation_rel"] < 1e-5 and checks["modes_counted"] == checks["modes_expected"] and \
    checks["n2d_bins"] == checks["n2d_expected"] and checks["Nmodes2D_sum"] == checks["modes_expected"] and checks["Pk0_finite"]
alg_dep = n ** 3 * (12 + 4) + 8 * n ** 3
print(json.dumps({
    "config": "BASELINE config 3: %d^3 %s particles + W -> %d^3 grid, MA(%s,W) -> delta -> Pk(axis=%d) incl. Pk2D" % (n, kind, n, MAS, AXIS),
    "ms": med, "ms_step": step, "particles_per_s_step": n ** 3 / step * 1e3,
    "ma_particles_per_s": n ** 3 / med["deposit"] * 1e3,
    "deposit_algorithmic_GBps": alg_dep / med["deposit"] / 1e6,
    "checks": checks, "ok": ok, "generate_s": gen_s,
    "hbm_peak_allocated_GB": torch.cuda.max_memory_allocated() / 1e9}))
sys.exit(0 if ok else 1)
