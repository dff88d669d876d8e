#!/usr/bin/env python
"""Slab FFT timings at the config-5 shapes for one setting of PYL_FFT_X_BATCH / PYL_FFT_YZ_BATCH (env)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pylians3_b200 import dist as PD
N, P = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda", 0)
ops = PD.DeviceOps(dev)
nx, nz = N // P, N // 2 + 1
nky = len(PD.mirrored_rows(N, 0, PD.split_sizes(N // 2 + 1, P)[0][0]))
def t(name, fn, reps=2):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    print("%s X=%s YZ=%s: %.2f ms" % (name, os.environ.get("PYL_FFT_X_BATCH", "auto"), os.environ.get("PYL_FFT_YZ_BATCH", "auto"), a.elapsed_time(b) / reps), flush=True)
if sys.argv[3] == "x":
    cols = torch.view_as_complex(torch.randn((N, nky, nz, 2), dtype=torch.float32, device=dev))
    t("fft_x", lambda: ops.fft_x_(cols, N))
else:
    slab = torch.randn((nx, N, N), dtype=torch.float32, device=dev)
    ring = torch.empty((16, N, nz), dtype=torch.complex64, device=dev)
    def yz():
        for b0 in range(0, nx, 16):
            ops.fft_yz(slab[b0:b0 + 16], N, out=ring)
    t("fft_yz", yz)
