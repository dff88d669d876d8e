#!/usr/bin/env python
"""The per-GPU pieces of BASELINE config 5 (4096^3 grid over 8 ranks) timed on ONE GPU with the shapes of rank 0:
slab deposit of 2048^3/8 particles into 512 planes, delta, 2D FFTs of the slab, 1D FFT along x of the rank's ky
rows, mirrored bin kernel.  (The transposes need the peers and are timed by bench.py on 8 GPUs.)

    python scratch/config5_pieces.py [N=4096] [P=8] [reps=2]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pylians3_b200 import Pk_library as PKL, _device as D, dist as PD, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
P = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
BOX = 1000.0
dev = torch.device("cuda", 0)
ops = PD.DeviceOps(dev)
nx, nz = N // P, N // 2 + 1
n_local = (N // 2) ** 3 // P
ylo_sizes, ylo_offs = PD.split_sizes(N // 2 + 1, P)
nky = len(PD.mirrored_rows(N, ylo_offs[0], ylo_sizes[0]))


def timeit(name, fn, nbytes=None):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print("%-28s %9.3f ms%s" % (name, ms, "" if nbytes is None else "   %.0f GB/s of %.1f GB" % (nbytes / ms / 1e6, nbytes / 1e9)), flush=True)


pos = synth.uniform_device(n_local, BOX, 5000, dev, x_range=synth.slab_x_bounds(0, nx, N, BOX))
store = torch.zeros((nx + 3, N, N), dtype=torch.float32, device=dev)
work = store[:nx + 1]
dropped = torch.zeros(1, dtype=torch.int64, device=dev)
timeit("zero slab", lambda: store.zero_(), 4.0 * store.numel())
timeit("deposit_slab CIC", lambda: ops.deposit_slab("CIC", pos, work, None, N, BOX, 0, nx, dropped),
       n_local * 12 + 8.0 * nx * N * N)
print("dropped", int(dropped.item()))
slab = store[:nx]
tot = ops.sum_f64(slab)
timeit("sum_f64", lambda: ops.sum_f64(slab), 4.0 * slab.numel())
timeit("overdensity", lambda: ops.overdensity_(slab, tot, float(N) ** 3), 8.0 * slab.numel())
del pos
D.release_workspaces(); torch.cuda.empty_cache()
ring = torch.empty((16, N, nz), dtype=torch.complex64, device=dev)


def yz():
    for b0 in range(0, nx, 16):
        ops.fft_yz(slab[b0:b0 + 16], N, out=ring)


timeit("fft_yz (all planes, ring)", yz, 4.0 * slab.numel() + 8.0 * nx * N * nz)
del store, work, slab, ring
torch.cuda.empty_cache()
cols = torch.randn((N, nky, nz, 2), dtype=torch.float32, device=dev)
cols = torch.view_as_complex(cols)
timeit("fft_x (in place, nky=%d)" % nky, lambda: ops.fft_x_(cols, N), 16.0 * cols.numel())
for axis in (0, 1, 2):
    timeit("bin mirrored axis=%d" % axis, lambda: ops.bin([cols], [PKL.MAS_function("CIC")], N, axis, True, ylo_offs[0], ylo_sizes[0]),
           8.0 * cols.numel())
cols = cols.view(nky, N, nz)            # the layout of the peer-memory transpose: ky rows outermost, x in the middle
timeit("fft_x ky-major (per ky plane)", lambda: ops.fft_x_kymajor_(cols, N), 16.0 * cols.numel())
for axis in (0, 1, 2):
    timeit("bin mirrored ky-major axis=%d" % axis, lambda: ops.bin([cols], [PKL.MAS_function("CIC")], N, axis, True, ylo_offs[0], ylo_sizes[0]),
           8.0 * cols.numel())
print("peak GB", torch.cuda.max_memory_allocated() / 1e9)
