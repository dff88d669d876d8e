#!/usr/bin/env python
"""bench.py -- MA -> delta -> Pk throughput of the B200-native path (and of the reference's CPU path).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic input (BASELINE.json config 2 at
N=1): zero the grid, MAS_library.MA (CIC) of dims^3 uniform-random particles onto a dims^3 grid,
delta = n/<n> - 1, Pk_library.Pk (FFT + deconvolution + binning of l=0,2,4, 1D and 2D spectra),
results on the host.  `value` = particles per second through that whole step with the positions
resident in HBM; `e2e` = the same step with the positions starting in pinned HOST memory (H2D copy
inside the timed region).  Rank 0 prints ONE JSON line.

N>1 (torchrun, one rank per GPU): weak scaling -- each rank owns an x-slab of a larger grid and the
same number of particles; halo exchange, distributed FFT and bin all-reduce are inside the step.
"""
import argparse
import contextlib
import io
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BOX = 1000.0
MAS = "CIC"
AXIS = 0
# weak-scaling ladder: ~512^3 particles and cells per GPU, FFT-friendly sizes (2^a 5^b)
GRID_FOR_GPUS = {1: 512, 2: 640, 4: 800, 8: 1024}
# DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of one MA(CIC) call at 512^3/512^3 from the
# `ncu --set full` capture summarised in profiles/r1_tiled_deposit_cic_final.md: tile_count 0.54 GB +
# tile_scatter 3.90 GB + tile_deposit 3.21 GB
NCU_DEPOSIT_TRAFFIC = {("CIC", 512): 7.65e9}


_RESULT = None


def claim_stdout():
    """Keep stdout for the ONE JSON line: a duplicate of the original fd 1 is kept for the result, and fd 1 itself is
    pointed at stderr, so that anything a library prints to stdout (NCCL's "NCCL version ..." banner arrives on the
    C-level stdout even with NCCL_DEBUG_FILE set) lands on stderr instead."""
    global _RESULT
    if _RESULT is None:
        sys.stdout.flush()
        _RESULT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _RESULT if _RESULT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def measured_peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# reference / CPU baseline legs (the only places that may execute oracle/)
# ------------------------------------------------------------------------------------------------
def cpu_path():
    """(MA, Pk, kind): the compiled unmodified reference if oracle/_ref travelled here, else the port."""
    from oracle import ref_loader
    if ref_loader.have_ref():
        M, P = ref_loader.ref_MASL(), ref_loader.ref_PKL()
        return M.MA, (lambda d, box, axis, mas, thr: P.Pk(d, box, axis, mas, thr, False)), "reference"
    from oracle import build as obuild
    obuild.build()
    from oracle import cpu as O
    return O.MA, (lambda d, box, axis, mas, thr: O.Pk(d, box, axis, mas, thr, False)), "port"


def cpu_step(MA, Pk, pos, dims, threads):
    grid = np.zeros((dims, dims, dims), np.float32)
    MA(pos, grid, BOX, MAS)
    grid /= np.mean(grid, dtype=np.float64)
    grid -= 1.0
    return Pk(grid, BOX, AXIS, MAS, threads)


CPU_SAMPLE_DIMS = 256


def cpu_baseline(reps=2):
    """Reference CPU path on a bounded sample of the workload: 256^3 particles -> 256^3 grid + Pk
    (1/8 of the 512^3 step in particles and in modes), best of `reps`."""
    MA, Pk, kind = cpu_path()
    dims = CPU_SAMPLE_DIMS
    pos = np.random.default_rng(1).random((dims ** 3, 3), dtype=np.float32) * np.float32(BOX)
    threads = os.cpu_count() or 1
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_step(MA, Pk, pos, dims, threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": dims ** 3 / best, "unit": "particles/s", "cores": 1 if kind == "port" else threads,
            "kind": kind, "seconds_per_sample": best,
            "sample": "%d^3 uniform particles -> %d^3 grid, MA(%s) + delta + Pk(axis=%d): 1/8 of one step; "
                      "MA is serial in the reference, threads=%d reach only its FFT" % (dims, dims, MAS, AXIS, threads)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    MA, Pk, kind = cpu_path()
    dims = CPU_SAMPLE_DIMS
    pos = np.random.default_rng(1).random((dims ** 3, 3), dtype=np.float32) * np.float32(BOX)
    threads = os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_step(MA, Pk, pos, dims, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(MA, Pk, pos, dims, threads)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    value = dims ** 3 / dt
    grid = GRID_FOR_GPUS.get(args.gpus, 512)
    sample = ("each step = %d^3 uniform particles -> %d^3 grid, MA(%s)+delta+Pk(axis=%d) on the host CPU "
              "(bounded sample of the %d^3 workload; particles/s is size-normalised)" % (dims, dims, MAS, AXIS, grid))
    line = {"impl": "reference", "metric": metric_name(grid), "value": value, "unit": "particles/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(grid, args.gpus),
            "cpu_baseline": {"value": value, "unit": "particles/s", "cores": threads if kind == "reference" else 1,
                             "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def metric_name(grid):
    return "MA+Pk particles/sec (%s, %d^3 uniform particles -> %d^3 grid, Pk l=0,2,4 + 1D + 2D, axis=%d)" % (
        MAS, grid, grid, AXIS)


def workload_config(grid, gpus):
    return {"workload": "BASELINE config 2 shape: %d^3 uniform-random float32 particles, %d^3 grid, BoxSize=%g, "
                        "MA(%s) -> delta=n/<n>-1 -> Pk(axis=%d)" % (grid, grid, BOX, MAS, AXIS),
            "particles": grid ** 3, "grid": grid, "mas": MAS, "axis": AXIS, "gpus": gpus,
            "decomposition": "single GPU" if gpus == 1 else "x-slabs, halo exchange + slab FFT all-to-all + bin all-reduce",
            "l2": "inputs larger than L2 (positions %.2f GB + grid %.2f GB per GPU >> 126 MB)" % (
                grid ** 3 * 12 / gpus / 1e9, grid ** 3 * 4 / gpus / 1e9)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # rank 0 prints ONE JSON line on stdout: NCCL's own version/debug lines go to a file instead
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/pyl_b200_nccl_%h_%p.log")
        dist.init_process_group("nccl", device_id=dev)

    from pylians3_b200 import MAS_library as MASL, Pk_library as PKL, _lib, overdensity_, synth
    lib = _lib.load()
    grid_n = args.grid or GRID_FOR_GPUS.get(world, 512)
    npart = grid_n ** 3

    if world > 1:
        from pylians3_b200 import dist as PD
        ctx = PD.SlabContext(grid_n, BOX)
        x0, x1 = ctx.x_range
        cell = BOX / grid_n
        n_local = npart // world
        pos = synth.uniform_device(n_local, BOX, 1000 + rank, dev, x_range=(x0 * cell, x1 * cell))
        slab = ctx.new_slab()

        def step(p):
            slab.zero_()
            ctx.MA(p, slab, MAS, routed=True)
            ctx.overdensity_(slab)
            return ctx.Pk(slab, AXIS, MAS)
        n_step_particles = n_local * world
    else:
        pos = synth.uniform_device(npart, BOX, 1, dev)
        grid = torch.zeros((grid_n, grid_n, grid_n), dtype=torch.float32, device=dev)

        def step(p):
            grid.zero_()
            MASL.MA(p, grid, BOX, MAS)
            overdensity_(grid)
            return PKL.Pk(grid, BOX, AXIS, MAS, verbose=False)
        n_step_particles = npart

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, with_ma_events=False):
        for _ in range(warmup):
            fn()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    # ---- timed region 1: device-resident inputs --------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    n0 = lib.pyl_kernel_launches()
    pk_holder = {}

    def fn_dev():
        pk_holder["pk"] = step(pos)

    for _ in range(args.warmup):
        fn_dev()
    barrier()
    launches0 = lib.pyl_kernel_launches()
    if sampler:
        sampler.start()
    ms_step = timed(fn_dev, args.steps, 0)
    clocks = sampler.stop() if sampler else None
    gpu_launches = lib.pyl_kernel_launches() - launches0
    value = n_step_particles / (ms_step * 1e-3)

    # ---- dominant kernel (the deposit) timed live with CUDA events -------------------------------
    ma_ms = None
    if world == 1:
        evs = []
        for _ in range(max(3, args.steps)):
            grid.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            MASL.MA(pos, grid, BOX, MAS)
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        ma_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))

    # ---- stage breakdown and per-scheme deposit rates (outside the timed region) -----------------
    stages, ma_rates = {}, {}
    if world == 1:
        def ev():
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            return e
        reps = 3
        acc = {k: 0.0 for k in ("zero", "deposit", "overdensity", "fft", "bin+finalise+d2h")}
        for _ in range(reps):
            e0 = ev(); grid.zero_()
            e1 = ev(); MASL.MA(pos, grid, BOX, MAS)
            e2 = ev(); overdensity_(grid)
            e3 = ev(); dk = PKL.fft3d_r2c_device(grid)
            e4 = ev(); PKL.spectra([dk], [PKL.MAS_function(MAS)], grid_n, AXIS, BOX, want_phase=True)
            e5 = ev(); torch.cuda.synchronize()
            for k, (a, b) in zip(acc, ((e0, e1), (e1, e2), (e2, e3), (e3, e4), (e4, e5))):
                acc[k] += a.elapsed_time(b) / reps
            del dk
        stages = {k: round(v, 4) for k, v in acc.items()}
        # binning kernel alone (device time, no D2H)
        dk = PKL.fft3d_r2c_device(grid)
        a = ev()
        for _ in range(reps):
            PKL.bin_device([dk], [2], grid_n, AXIS, True)
        b = ev(); torch.cuda.synchronize()
        stages["bin_kernels_only"] = round(a.elapsed_time(b) / reps, 4)
        del dk
        W = synth.weights_device(npart, 1, dev)
        for mode in ("auto", "atomic"):
            for mas in ("NGP", "CIC", "TSC", "PCS"):
                for w, tag in ((None, ""), (W, "+W")):
                    if mode == "atomic" and w is not None:
                        continue
                    MASL.MA(pos, grid, BOX, mas, w, mode=mode)
                    a = ev()
                    for _ in range(reps):
                        MASL.MA(pos, grid, BOX, mas, w, mode=mode)
                    b = ev(); torch.cuda.synchronize()
                    ma_rates[mas + tag + ("" if mode == "auto" else "[atomic]")] = npart / (a.elapsed_time(b) / reps * 1e-3)
        del W

    if world > 1:
        # stage breakdown of the distributed step (device time on this rank, max over ranks)
        def ev():
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            return e
        reps = 3
        acc = {k: 0.0 for k in ("zero", "deposit+halo", "overdensity", "fft(yz,transpose,x)", "bin+allreduce+finalise+d2h")}
        for _ in range(reps):
            barrier()                      # ranks start each repetition together: no inter-rank skew in the stage times
            e0 = ev(); slab.zero_()
            e1 = ev(); ctx.MA(pos, slab, MAS, routed=True)
            e2 = ev(); ctx.overdensity_(slab)
            e3 = ev(); dk = ctx.fft(slab)
            e4 = ev(); ctx._spectra([dk], [PKL.MAS_function(MAS)], AXIS, True)
            e5 = ev(); torch.cuda.synchronize()
            for k, (a, b) in zip(acc, ((e0, e1), (e1, e2), (e2, e3), (e3, e4), (e4, e5))):
                acc[k] += a.elapsed_time(b) / reps
            del dk
        t = torch.tensor(list(acc.values()), dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        stages = {k: round(float(v), 4) for k, v in zip(acc, t.tolist())}

    # ---- timed region 2: end to end from pinned host memory --------------------------------------
    pos_host = torch.empty(pos.shape, dtype=torch.float32, pin_memory=True)
    pos_host.copy_(pos)
    torch.cuda.synchronize()
    h2d = pos_host.numel() * 4
    d2h = _lib.pk_layout(grid_n, 1).total_words * 8
    del pos
    torch.cuda.empty_cache()

    def fn_e2e():
        pk_holder["pk"] = step(pos_host)

    e2e_steps = max(2, min(args.steps, 5))
    ms_e2e = timed(fn_e2e, e2e_steps, 1)
    e2e_value = n_step_particles / (ms_e2e * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_hbm()
    roofline = None
    if ma_ms is not None:
        alg_bytes = npart * 12 + 8 * grid_n ** 3          # positions once + grid RMW once
        achieved = alg_bytes / (ma_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "deposit (%s, %s)" % (MAS, "pyl_deposit"), "achieved": achieved,
                    "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": NCU_DEPOSIT_TRAFFIC.get((MAS, grid_n)),
                    "traffic_source": "profiles/r1_tiled_deposit_cic_final.md (ncu --set full, per MA call)",
                    "kernels": "tile_count + tile_scatter + tile_deposit (one pyl_deposit call)",
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                    "avg_launch_ms": ma_ms, "share_of_step": ma_ms / ms_step}
    if roofline is None and world > 1:
        # per-GPU deposit roofline of the sharded step, from the stage breakdown (max over ranks; the stage
        # includes the ghost-plane exchange).  Guarded: a missing stage leaves the key null, never breaks the line.
        try:
            dep_ms = float(stages["deposit+halo"])
            alg_bytes = (npart * 12 + 8 * grid_n ** 3) / world
            achieved = alg_bytes / (dep_ms * 1e-3) / 1e9
            roofline = {"bound": "hbm", "kernel": "deposit + halo exchange, per GPU (%s, pyl_deposit_slab)" % MAS,
                        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                        "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": dep_ms,
                        "share_of_step": dep_ms / ms_step}
        except Exception:
            roofline = None
    pk = pk_holder["pk"]
    line = {"metric": metric_name(grid_n), "value": value, "unit": "particles/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(grid_n, world),
            "e2e": {"value": e2e_value, "unit": "particles/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e, "steps": e2e_steps},
            "gpu_launches": int(gpu_launches), "clocks": clocks, "roofline": roofline,
            "stages_ms": stages, "ma_particles_per_s": ma_rates,
            "check": {"Pk0_mean_over_shot_noise": float(np.mean(pk.Pk[10:200, 0]) / (BOX ** 3 / npart)),
                      "modes_counted": int(pk.Nmodes3D.sum()) + 1}}
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline()
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=0, help="override the grid side (default: by --gpus)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
