"""Micro-benchmark of the peer-memory transpose ALONE (no concurrent FFT) on P GPUs, against a plain peer copy.

    torchrun --nproc-per-node 2 scratch/transpose_bench.py 2048 [planes]

Prints per layout the kernel time, the bytes that left the GPU per second, and the rate of a peer-to-peer
tensor copy (copy engine over NVLink) of the same number of bytes."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylians3_b200 import dist as PD  # noqa: E402


def main():
    N = int(sys.argv[1])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    ctx = PD.SlabContext(N, 1000.0)
    nx = min(ctx.nx, int(sys.argv[2]) if len(sys.argv) > 2 else 64)
    nz = ctx.nz
    a = torch.randn((nx, N, nz, 2), dtype=torch.float32, device=dev).view(torch.float32)
    a = torch.view_as_complex(a.view(nx, N, nz, 2))
    buf, hdl, ptrs = ctx._peer_slot(0)
    sent = a.numel() * 8 * (world - 1) / world
    out = {}
    for name, ky_major, order in (("normal", False, True), ("ky-major", True, True), ("ky-major, plain order", True, False)):
        ts = []
        for rep in range(6):
            hdl.barrier(channel=0)
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ctx.ops.transpose_scatter(a, ptrs, ctx._peer["nky"], ctx._peer["owner"], ctx._peer["row"], N, ctx.x_range[0],
                                      ky_major=ky_major, ky_order=ctx._peer["order"] if order else None)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = torch.tensor([sorted(ts[1:])[len(ts[1:]) // 2]], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name] = (float(t), sent / float(t) / 1e6)
    # reference: copy-engine peer copy of the same remote bytes (to the next rank)
    peer = hdl.get_buffer((rank + 1) % world, (buf.numel(),), torch.complex64)
    n = int(sent // 8)
    n = min(n, buf.numel(), a.numel())
    src = a.view(-1)[:n]
    ts = []
    for rep in range(5):
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        peer[:n].copy_(src)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts[1:])[len(ts[1:]) // 2]
    if rank == 0:
        print("grid %d, %d ranks, %d planes per rank, %.2f GB leave each GPU; lib %s" % (
            N, world, nx, sent / 1e9, os.environ.get("PYL_B200_SO", "default")))
        for k, (ms, gbs) in out.items():
            print("  transpose kernel, %-22s %8.3f ms  %7.1f GB/s sent" % (k, ms, gbs))
        print("  peer tensor copy (copy engine)           %8.3f ms  %7.1f GB/s" % (t, n * 8 / t / 1e6))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
