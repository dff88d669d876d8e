"""GPU, 2/4/8 ranks over NCCL: the slab-decomposed path (pylians3_b200.dist with the real CUDA kernels)
against the CPU oracle.  Cases needing more GPUs than the box has are skipped (the CPU/gloo twin is test_dist_cpu.py)."""
import os
import socket
import sys
import traceback

import numpy as np
import pytest

from conftest import BOX, ROOT, make_particles, rel_err

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, q, transpose="peer"):
    try:
        ky_major = transpose == "peer-kymajor"
        transpose = "peer" if ky_major else transpose
        os.environ["PYL_TRANSPOSE"] = transpose
        import torch
        import torch.distributed as dist
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        from pylians3_b200 import dist as PD
        from oracle import cpu as O
        ctx = PD.SlabContext(N, BOX)
        if ky_major:
            ctx.KY_MAJOR_MIN_DIMS = 0           # the (nky, N, nz) spectrum layout of the large grids, forced here
        assert (ctx._peer is not None) == (transpose == "peer"), "peer-memory transpose not active"
        pos, W = make_particles(77, 4 * N ** 3, True)
        mine = slice(rank, None, world)
        x0, x1 = ctx.x_range
        res = {}
        slabs, refs = {}, {}
        for mas, w in (("NGP", None), ("CIC", W), ("TSC", None), ("PCS", W)):
            slab = ctx.new_slab()
            ctx.MA(torch.from_numpy(pos[mine].copy()).to(dev), slab, mas,
                   None if w is None else torch.from_numpy(w[mine].copy()).to(dev))
            ref = np.zeros((N, N, N), np.float32)
            O.MA(pos, ref, BOX, mas, w)
            got = slab.cpu().numpy()
            res["ma_" + mas] = rel_err(got, ref[x0:x1], floor=float(np.mean(np.abs(ref))))
            slabs[mas], refs[mas] = slab, ref
        ctx.check_dropped()
        dens = {mas: slabs[mas].clone() for mas in ("CIC", "PCS")}
        for mas in ("CIC", "PCS"):
            ref = refs[mas]
            ref /= np.mean(ref, dtype=np.float64)
            ref -= 1.0
            ctx.overdensity_(slabs[mas])
            res["delta_" + mas] = float(np.max(np.abs(slabs[mas].cpu().numpy() - ref[x0:x1])) / np.max(np.abs(ref)))
            # the spectra are compared on the delta the slabs actually hold (gathered), so that the
            # float32 accumulation-order noise of the deposit (checked above) is not mistaken for an
            # FFT/binning error: low-power bins amplify it by P_peak/P_bin
            nmax = max(ctx.x_sizes)
            mine_p = torch.zeros((nmax, N, N), dtype=torch.float32, device=dev)
            mine_p[:ctx.nx] = slabs[mas]
            parts = [torch.empty_like(mine_p) for _ in range(world)]
            dist.all_gather(parts, mine_p)
            refs[mas] = torch.cat([parts[r][:ctx.x_sizes[r]] for r in range(world)]).cpu().numpy()
        from test_gpu_pk import check_pk, quiet
        for axis in (0, 1, 2):
            got = ctx.Pk(slabs["PCS"], axis, "PCS")
            want = O.Pk(refs["PCS"], BOX, axis, "PCS", 1, False)
            check_pk(got, want)
        gx = ctx.XPk([slabs["PCS"], slabs["CIC"]], 0, ["PCS", "CIC"])
        wx = quiet(O.XPk, [refs["PCS"], refs["CIC"]], BOX, 0, ["PCS", "CIC"], 1)
        check_pk(gx, wx, cross=True)
        # density=True: the slabs keep n, the normalisation is the scale of the binned sums (DC mode all-reduced)
        check_pk(ctx.Pk(dens["PCS"], 1, "PCS", density=True, offset=0.0), O.Pk(refs["PCS"], BOX, 1, "PCS", 1, False),
                 phase_min_modes=64)
        check_pk(ctx.XPk([dens["PCS"], dens["CIC"]], 0, ["PCS", "CIC"], density=True, offset=0.0), wx, cross=True)
        # the bench step's recipe: every rank fills its slab with the same -c, deposits, and passes c on
        slab = ctx.new_slab()
        w_mine = torch.from_numpy(W[mine].copy()).to(dev)
        c = ctx.prebias_(slab, len(w_mine), w_mine)
        ctx.MA(torch.from_numpy(pos[mine].copy()).to(dev), slab, "PCS", w_mine)
        check_pk(ctx.Pk(slab, 1, "PCS", density=True, offset=c), O.Pk(refs["PCS"], BOX, 1, "PCS", 1, False),
                 phase_min_modes=64)
        q.put((rank, "ok", res))
        dist.destroy_process_group()
    except Exception:
        q.put((rank, "fail", traceback.format_exc()))


@pytest.mark.parametrize("world,N,transpose", [(2, 64, "peer"), (2, 45, "peer"), (2, 45, "nccl"), (2, 64, "peer-kymajor"),
                                               (4, 64, "peer"), (4, 45, "peer-kymajor"), (8, 64, "peer"),
                                               (8, 72, "peer-kymajor")])
def test_multi_gpu_slab_pipeline(oracle, world, N, transpose):
    """transpose="peer": the slab-FFT transpose is one kernel storing into the owners' symmetric receive buffers
    over NVLink (pyl_transpose_scatter); "peer-kymajor": the same with the (nky, N, nz) layout used from 2048^3 on;
    "nccl": pack + all_to_all_single."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, q, transpose)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, payload in out:
        assert status == "ok", "rank %d:\n%s" % (rank, payload)
        assert payload["ma_NGP"] == 0.0
        for k in ("ma_CIC", "ma_TSC", "ma_PCS"):
            assert payload[k] < 1e-5, (k, payload[k])
        for k in ("delta_CIC", "delta_PCS"):
            assert payload[k] < 1e-5, (k, payload[k])
