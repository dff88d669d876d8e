#!/usr/bin/env python
"""Tiny driver for ncu captures of one stage (kept small so ncu replays stay cheap).

    python profiles/run_stage.py deposit CIC auto 512 [reps] [uniform|zeldovich] [W]
    python profiles/run_stage.py pk 512 [axis] [reps]
    python profiles/run_stage.py step 512 [reps]          # config 2 step: zero, MA(CIC), delta, Pk
    python profiles/run_stage.py step3 1024 [reps]        # the bench.py step (config 3): Zel'dovich + W, PCS, Pk
    python profiles/run_stage.py shell 512 [reps]         # shell kernels (theta, dv, vv, xi) + mode passes
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from pylians3_b200 import MAS_library as MASL, Pk_library as PKL, synth, overdensity_  # noqa: E402

BOX = 1000.0
what = sys.argv[1]
dev = torch.device("cuda", 0)
if what == "deposit":
    mas, mode, N = sys.argv[2], sys.argv[3], int(sys.argv[4])
    reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
    kind = sys.argv[6] if len(sys.argv) > 6 else "uniform"
    pos = synth.uniform_device(N ** 3, BOX, 1, dev) if kind == "uniform" else synth.zeldovich_device(N, BOX, 1, dev)
    W = synth.weights_device(N ** 3, 1, dev) if len(sys.argv) > 7 and sys.argv[7] == "W" else None
    grid = torch.zeros((N, N, N), dtype=torch.float32, device=dev)
    for _ in range(reps):
        MASL.MA(pos, grid, BOX, mas, W, mode=mode)
    torch.cuda.synchronize()
    print("sum/N^3 =", float(grid.sum(dtype=torch.float64)) / N ** 3 / reps)
elif what == "shell":
    from pylians3_b200 import _pk_more as PM
    N = int(sys.argv[2])
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    g = torch.Generator(device=dev); g.manual_seed(1)
    real = [torch.randn((N, N, N), generator=g, device=dev, dtype=torch.float32) for _ in range(2)]
    cplx = [PKL.fft3d_r2c_device(real[i % 2]) for i in range(6)]
    for _ in range(reps):
        PM.shell_bin("theta", cplx[:3], [2], N)
        PM.shell_bin("dv", cplx[:4], [2], N)
        PM.shell_bin("vv", cplx[:6], [2], N)
        r = PM.shell_bin("xi", real[:1], [], N, axis=2, scale=1.0 / N ** 3)
        PM._modes("deconvolve", cplx[0], None, N, 2, 0)
        PM._modes("power", cplx[1], cplx[2], N, 2, 4)
    torch.cuda.synchronize()
    print("xi Nm sum =", r["Nm"].sum())
elif what == "step3":
    # the bench.py step on BASELINE config 3: Zel'dovich particles + W, PCS, Pk incl. Pk2D
    N = int(sys.argv[2])
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    pos = synth.zeldovich_device(N, BOX, 3, dev)
    W = synth.weights_device(N ** 3, 3, dev)
    torch.cuda.empty_cache()
    grid = torch.zeros((N, N, N), dtype=torch.float32, device=dev)
    from pylians3_b200 import prebias_
    for _ in range(reps):
        c = prebias_(grid, pos.shape[0], W)
        MASL.MA(pos, grid, BOX, "PCS", W)
        pk = PKL.Pk(grid, BOX, 0, "PCS", verbose=False, density=True, offset=c)
    torch.cuda.synchronize()
    print("Pk0[:3] =", pk.Pk[:3, 0])
elif what == "step":
    N = int(sys.argv[2])
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    pos = synth.uniform_device(N ** 3, BOX, 1, dev)
    grid = torch.zeros((N, N, N), dtype=torch.float32, device=dev)
    for _ in range(reps):
        grid.zero_()
        MASL.MA(pos, grid, BOX, "CIC")
        overdensity_(grid)
        pk = PKL.Pk(grid, BOX, 0, "CIC", verbose=False)
    torch.cuda.synchronize()
    print("Pk0[:3] =", pk.Pk[:3, 0])
else:
    N = int(sys.argv[2])
    axis = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    pos = synth.uniform_device(N ** 3, BOX, 1, dev)
    grid = torch.zeros((N, N, N), dtype=torch.float32, device=dev)
    MASL.MA(pos, grid, BOX, "CIC")
    del pos
    overdensity_(grid)
    for _ in range(reps):
        pk = PKL.Pk(grid, BOX, axis, "CIC", verbose=False)
    print("Pk0[:3] =", pk.Pk[:3, 0])
