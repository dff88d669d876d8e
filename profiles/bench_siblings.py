#!/usr/bin/env python
"""Kernel timings of the sibling estimators at one grid size (CUDA events, after warm-up), with the algorithmic
bytes each launch must move and the fraction of the measured copy bandwidth it reaches.

    python profiles/bench_siblings.py [N=512] [reps=5]  > gpurun_out/siblings.md

Algorithmic bytes: shell kernels read every stored complex mode of every field once (8 B * N^2 (N/2+1) * fields;
the Xi binning reads 4 B * N^3); the mode passes read and write the half-spectrum once (16 B per mode, 24 B for the
cross form).  Bins are negligible.  Inputs are resident in HBM; transforms are excluded (cuFFT, library stage)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from pylians3_b200 import _pk_more as PM, Pk_library as PKL  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda", 0)
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    peak = 6650.0
g = torch.Generator(device=dev); g.manual_seed(1)
nz = N // 2 + 1
modes = N * N * nz
real = [torch.randn((N, N, N), generator=g, device=dev, dtype=torch.float32) for _ in range(2)]
cplx = [PKL.fft3d_r2c_device(real[i % 2]) for i in range(6)]
img = [PM.fft2d_r2c_device(torch.randn((8 * N, 8 * N), generator=g, device=dev, dtype=torch.float32)) for _ in range(2)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > L2 (126 MB)


def timed(fn):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


rows = []


def add(name, nbytes, fn):
    ms = timed(fn)
    rows.append((name, nbytes / 1e9, ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / peak))


add("shell theta  (3 fields)", 8 * modes * 3, lambda: PM.shell_bin("theta", cplx[:3], [2], N))
add("shell dv     (4 fields)", 8 * modes * 4, lambda: PM.shell_bin("dv", cplx[:4], [2], N))
add("shell vv     (6 fields)", 8 * modes * 6, lambda: PM.shell_bin("vv", cplx[:6], [2], N))
add("shell xi     (real grid)", 4 * N ** 3, lambda: PM.shell_bin("xi", real[:1], [], N, axis=2, scale=1.0 / N ** 3))
M = 8 * N
add("shell plane  (%d^2 image)" % M, 8 * M * (M // 2 + 1), lambda: PM.shell_bin("plane", img[:1], [2], M))
add("shell xplane (2 images)", 16 * M * (M // 2 + 1), lambda: PM.shell_bin("xplane", img, [2, 4], M))
work = cplx[0].clone()
add("modes deconvolve", 16 * modes, lambda: PM._modes("deconvolve", work, None, N, 2, 0))
add("modes power (auto)", 16 * modes, lambda: PM._modes("power", work, None, N, 2, 0))
add("modes power (cross)", 24 * modes, lambda: PM._modes("power", work, cplx[1], N, 2, 4))
add("Pk bin, F=1 + phase (for scale)", 8 * modes, lambda: PKL.bin_device(cplx[:1], [2], N, 0, True))
add("XPk bin, F=2", 16 * modes, lambda: PKL.bin_device(cplx[:2], [2, 2], N, 0))

print("# Sibling-estimator kernels at N = %d (B200, CUDA events, median of %d, L2 flushed between launches)\n" % (N, reps))
print("Times include the D2H of the O(bins) accumulator block for the shell kernels (`shell_bin` returns host sums).\n")
print("| launch | algorithmic GB | ms | GB/s | fraction of %.0f GB/s |" % peak)
print("|---|---|---|---|---|")
for r in rows:
    print("| %s | %.3f | %.3f | %.0f | %.3f |" % r)
