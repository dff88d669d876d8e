"""torchrun script: fine-grained timing of the distributed Pk stage (bin kernel / all-reduce / D2H / finalise)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from pylians3_b200 import dist as PD, Pk_library as PKL, synth
N = int(sys.argv[1]); BOX = 1000.0
ctx = PD.SlabContext(N, BOX)
x0, x1 = ctx.x_range; cell = BOX / N
pos = synth.uniform_device(N ** 3 // world, BOX, 1000 + rank, dev, x_range=(x0 * cell, x1 * cell))
slab = ctx.new_slab()
def sync(): torch.cuda.synchronize()
def T(fn, reps=3):
    fn(); sync(); dist.barrier(); sync()
    t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    sync(); return (time.perf_counter() - t0) / reps * 1e3, r
t_ma, _ = T(lambda: (slab.zero_(), ctx.MA(pos, slab, "CIC", routed=True)))
t_od, _ = T(lambda: ctx.overdensity_(slab))
t_sum, tot = T(lambda: ctx.ops.sum_f64(slab))
t_ar1, _ = T(lambda: dist.all_reduce(tot))
t_fft, dk = T(lambda: ctx.fft(slab))
t_yz, a = T(lambda: ctx.ops.fft_yz(slab, N))
t_x, _ = T(lambda: ctx.ops.fft_x_(dk, N))
t_bin, (out, lay) = T(lambda: ctx.ops.bin([dk], [2], N, 0, True, ctx.ky_lo, ctx.ny_lo))
t_red, f64 = T(lambda: ctx._reduce(out.clone(), lay))
t_d2h, h = T(lambda: f64.cpu().numpy())
t_raw, raw = T(lambda: ctx._raw([dk], [2], 0, True))
t_fin, _ = T(lambda: PKL._finalize(raw, BOX, N))
t_spec, _ = T(lambda: ctx._spectra([dk], [2], 0, True))
out2, lay2 = ctx.ops.bin([dk], [2], N, 0, True, ctx.ky_lo, ctx.ny_lo)
red2 = ctx._reduce(out2, lay2)
t_find, _ = T(lambda: PKL.finalize_device(red2.clone().view(torch.int64), lay2, BOX, N, counts_are_f64=True))
def pin_only():
    h = torch.empty(red2.shape, dtype=torch.float64, pin_memory=True); h.copy_(red2, non_blocking=True); torch.cuda.current_stream().synchronize(); return h
t_pin, _ = T(pin_only)
if rank == 0:
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    for _ in range(5): PKL.finalize_device(red2.clone().view(torch.int64), lay2, BOX, N, counts_are_f64=True)
    pr.disable(); pstats.Stats(pr).sort_stats("tottime").print_stats(8)
if rank == 0: print("finalize_device alone %.2f ms; pinned alloc+D2H alone %.2f ms" % (t_find, t_pin), flush=True)
if rank == 0:
    print("N=%d world=%d  MA %.2f | overdensity %.2f (sum %.2f allreduce %.2f) | fft %.2f (yz %.2f, x %.2f -> transpose ~%.2f) | "
          "bin %.2f reduce %.2f d2h %.2f raw-total %.2f finalize %.2f | NEW spectra total %.2f [ms]" % (N, world, t_ma, t_od, t_sum, t_ar1, t_fft, t_yz, t_x,
          t_fft - t_yz - t_x, t_bin, t_red, t_d2h, t_raw, t_fin, t_spec), flush=True)
dist.destroy_process_group()
