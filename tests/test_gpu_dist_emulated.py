"""GPU, ONE device: the slab-decomposed pipeline of pylians3_b200.dist driven rank by rank in a single process.

The driver's GPU box has one B200, where the NCCL tests of test_gpu_dist.py skip.  This test runs every compute
kernel of the distributed path with the layouts of P = 2, 4 and 8 ranks -- slab-window deposits with ghost planes
(pyl_deposit_slab), per-slab 2D FFTs, the mirrored-ky column sets of every rank and the 1D FFT along x
(pyl_fft_slab_yz / pyl_fft_slab_x), the mirrored bin kernel, the count conversion and the device finalisation -- and
replaces only the communication (ghost-plane send/recv, transpose, all-reduce) by local copies and sums.  Results
are held to the CPU oracle with the bars of the NCCL test."""
import numpy as np
import pytest

from conftest import BOX, make_particles, rel_err

pytestmark = pytest.mark.gpu
S = {"NGP": 1, "CIC": 2, "TSC": 3, "PCS": 4}


class _Res:
    pass


@pytest.mark.parametrize("P,N,mas", [(2, 64, "PCS"), (4, 64, "CIC"), (8, 64, "PCS"), (8, 72, "TSC"), (3, 45, "PCS")])
def test_rank_by_rank_emulation(oracle, P, N, mas):
    import torch
    from pylians3_b200 import Pk_library as PKL, _device as D, _lib as L, dist as PD
    from test_gpu_pk import check_pk
    dev = torch.device("cuda", 0)
    ops = PD.DeviceOps(dev)
    lib = L.load()
    pos, W = make_particles(5 * N + P, 4 * N ** 3, True)
    pos_d, W_d = torch.from_numpy(pos).to(dev), torch.from_numpy(W).to(dev)
    x_sizes, x_offs = PD.split_sizes(N, P)
    ylo_sizes, ylo_offs = PD.split_sizes(N // 2 + 1, P)
    ky_rows = [PD.mirrored_rows(N, ylo_offs[r], ylo_sizes[r]) for r in range(P)]
    ghosts = S[mas] - 1

    # route: owner of the x-plane of the first stencil cell
    plane = ops.base_plane(mas, pos_d, N, BOX).cpu().numpy()
    owner = np.searchsorted(np.asarray(x_offs[1:]), plane, side="right")
    dropped = torch.zeros(1, dtype=torch.int64, device=dev)
    bufs = []
    for r in range(P):
        sel = torch.from_numpy(np.nonzero(owner == r)[0]).to(dev)
        buf = torch.zeros((x_sizes[r] + ghosts, N, N), dtype=torch.float32, device=dev)
        ops.deposit_slab(mas, pos_d[sel].contiguous(), buf, W_d[sel].contiguous(), N, BOX, x_offs[r], x_sizes[r], dropped)
        bufs.append(buf)
    assert int(dropped.item()) == 0
    slabs = []
    for r in range(P):
        slab = bufs[r][:x_sizes[r]].clone()
        if ghosts:
            slab[:ghosts] += bufs[(r - 1) % P][x_sizes[(r - 1) % P]:]       # the halo exchange
        slabs.append(slab)
    ref = np.zeros((N, N, N), np.float32)
    oracle.MA(pos, ref, BOX, mas, W)
    got = torch.cat(slabs).cpu().numpy()
    assert rel_err(got, ref, floor=float(np.mean(np.abs(ref)))) < 1e-5

    # delta with the global float64 sum (the all-reduce)
    dens = [s.clone() for s in slabs]
    total = sum(ops.sum_f64(s) for s in slabs)
    for s in slabs:
        ops.overdensity_(s, total, float(N) ** 3)
    delta = torch.cat(slabs).cpu().numpy()

    # slab FFT: 2D per slab, "transpose" = gather the rank's mirrored ky rows of every plane, 1D along x
    a = torch.cat([ops.fft_yz(s, N) for s in slabs])                        # (N, N, nz)
    cols = [ops.fft_x_(a[:, torch.tensor(ky_rows[r], device=dev), :].contiguous(), N) for r in range(P)]
    full = PKL.fft3d_r2c_device(torch.from_numpy(delta).to(dev))
    for r in range(P):                                                      # same transform as the 3D plan
        want = full[:, torch.tensor(ky_rows[r], device=dev), :]
        assert float((cols[r] - want).abs().max()) <= 2e-5 * float(full.abs().max())

    # the peer-memory transpose delivers (nky, N, nz) -- x in the middle -- and the x transforms run per ky plane
    cols_ky = [ops.fft_x_kymajor_(a[:, torch.tensor(ky_rows[r], device=dev), :].permute(1, 0, 2).contiguous(), N)
               for r in range(P)]
    for r in range(P):
        assert float((cols_ky[r] - cols[r].permute(1, 0, 2)).abs().max()) <= 2e-5 * float(full.abs().max())

    for axis, use in ((0, cols), (1, cols), (2, cols), (0, cols_ky), (1, cols_ky), (2, cols_ky)):
        acc, lay = None, None
        for r in range(P):
            out, lay = ops.bin([use[r]], [PKL.MAS_function(mas)], N, axis, True, ylo_offs[r], ylo_sizes[r])
            L.check(lib.pyl_pk_counts_to_f64(D.ptr(out), N, 1, D.stream_ptr(dev)), "pyl_pk_counts_to_f64")
            f = out.view(torch.float64)
            acc = f.clone() if acc is None else acc + f                     # the all-reduce
        o = PKL.finalize_device(acc.view(torch.int64), lay, BOX, N, counts_are_f64=True)
        g = _Res()
        g.k1D, g.Pk1D, g.Nmodes1D = o["k1D"], o["Pk1D"][:, 0], o["Nmodes1D"]
        g.kpar, g.kper, g.Pk2D, g.Nmodes2D = o["kpar"], o["kper"], o["Pk2D"][:, 0], o["Nmodes2D"]
        g.k3D, g.Nmodes3D = o["k3D"], o["Nmodes3D"]
        g.Pk, g.Pkphase = np.ascontiguousarray(o["Pk"][:, :, 0]), o["Pkphase"]
        check_pk(g, oracle.Pk(delta, BOX, axis, mas, 1, False), phase_min_modes=64)

    # density=True of SlabContext._spectra: the slabs keep the density, the rank with ky = 0 takes the DC mode, the
    # value rides the all-reduce behind the accumulators, the sums are scaled before the finalisation
    a = torch.cat([ops.fft_yz(s, N) for s in dens])
    acc, lay = None, None
    for r in range(P):
        col = ops.fft_x_(a[:, torch.tensor(ky_rows[r], device=dev), :].contiguous(), N)
        dc = PKL.take_dc([col], holds_dc=(ylo_offs[r] == 0 and ylo_sizes[r] > 0))
        assert (float(dc[0]) != 0.0) == (r == 0)
        out, lay = ops.bin([col], [PKL.MAS_function(mas)], N, 1, True, ylo_offs[r], ylo_sizes[r])
        L.check(lib.pyl_pk_counts_to_f64(D.ptr(out), N, 1, D.stream_ptr(dev)), "pyl_pk_counts_to_f64")
        f = out.view(torch.float64)
        f[lay.total_words:lay.total_words + 1] = dc
        acc = f.clone() if acc is None else acc + f
    assert abs(float(acc[lay.total_words]) / float(total) - 1.0) < 1e-6          # dims^3 <n> = sum of the cells
    PKL.density_scale_(acc.view(torch.int64), lay, N, acc[lay.total_words:lay.total_words + 1].clone())
    o = PKL.finalize_device(acc.view(torch.int64), lay, BOX, N, counts_are_f64=True)
    g = _Res()
    g.k1D, g.Pk1D, g.Nmodes1D = o["k1D"], o["Pk1D"][:, 0], o["Nmodes1D"]
    g.kpar, g.kper, g.Pk2D, g.Nmodes2D = o["kpar"], o["kper"], o["Pk2D"][:, 0], o["Nmodes2D"]
    g.k3D, g.Nmodes3D = o["k3D"], o["Nmodes3D"]
    g.Pk, g.Pkphase = np.ascontiguousarray(o["Pk"][:, :, 0]), o["Pkphase"]
    check_pk(g, oracle.Pk(delta, BOX, 1, mas, 1, False), phase_min_modes=64)


def test_single_rank_group_deposits_into_the_whole_grid():
    """A one-rank 'slab' is the whole periodic grid: pyl_deposit_slab with x_own = x_planes = dims takes the tiled
    periodic path (the advisor's round-1 finding: this used to be rejected as a bad plane window)."""
    import torch
    from pylians3_b200 import MAS_library as MASL, dist as PD
    dev = torch.device("cuda", 0)
    ops = PD.DeviceOps(dev)
    N = 64
    pos, W = make_particles(3, 4 * N ** 3, True)
    pos_d, W_d = torch.from_numpy(pos).to(dev), torch.from_numpy(W).to(dev)
    for mas in ("NGP", "CIC", "TSC", "PCS"):
        a = torch.zeros((N, N, N), dtype=torch.float32, device=dev)
        b = torch.zeros((N, N, N), dtype=torch.float32, device=dev)
        dropped = torch.zeros(1, dtype=torch.int64, device=dev)
        ops.deposit_slab(mas, pos_d, a, W_d, N, BOX, 0, N, dropped)
        MASL.MA(pos_d, b, BOX, mas, W_d, mode="tiled")
        assert int(dropped.item()) == 0
        # same kernels either way; only the order in which neighbouring tiles add their halo cells may differ
        assert float((a - b).abs().max()) <= 2e-6 * float(b.abs().max())
