#!/usr/bin/env python
"""Times MA per scheme/mode with CUDA events: python scratch/time_deposit.py N [modes] [schemes] [kind]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pylians3_b200 import MAS_library as MASL, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["tiled"]
schemes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["NGP", "CIC", "TSC", "PCS"]
kind = sys.argv[4] if len(sys.argv) > 4 else "uniform"
dev = torch.device("cuda", 0)
BOX = 1000.0
pos = synth.uniform_device(N ** 3, BOX, 1, dev) if kind == "uniform" else synth.zeldovich_device(N, BOX, 1, dev)
grid = torch.zeros((N, N, N), dtype=torch.float32, device=dev)
for mode in modes:
    for mas in schemes:
        MASL.MA(pos, grid, BOX, mas, mode=mode)
        reps = 5
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            MASL.MA(pos, grid, BOX, mas, mode=mode)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        print("%s %s %s N=%d fill_shift=%s: %.3f ms  %.2f Gpart/s" % (kind, mode, mas, N, os.environ.get("PYL_FILL_SHIFT", "default"), ms, N ** 3 / ms / 1e6), flush=True)
