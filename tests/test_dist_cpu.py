"""CPU, world_size 2 over gloo: the multi-GPU host logic of pylians3_b200.dist (routing, halo exchange,
slab-FFT transpose, bin all-reduce, finalisation) with the CUDA kernels replaced by numpy / oracle
stand-ins (TEST INFRASTRUCTURE -- the product's DeviceOps binds libpyl_b200.so instead)."""
import os
import socket
import sys
import traceback

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import BOX, ROOT, make_particles, rel_err


def base_cell(mas, dist_):
    d64 = dist_.astype(np.float64)
    if mas == "NGP":
        return np.trunc(d64 + 0.5).astype(np.int64)
    if mas == "CIC":
        return np.trunc(dist_).astype(np.int64)
    if mas == "TSC":
        return np.floor(d64 - 1.5).astype(np.int64) + 1
    return np.floor(d64 - 2.0).astype(np.int64) + 1


def numpy_bin(dk_list, mas_index, dims, axis, want_phase, ky_lo, nky, lay, ky_rows=None):
    """Vectorised restatement of Pk_library.pyx:311-378 / :623-732 on (dims, nky, nz) arrays holding the
    global ky rows `ky_rows` (default: the window ky_lo..ky_lo+nky-1), written into the accumulator layout
    of pyl_pk_layout_t."""
    N, m = dims, dims // 2
    nz = m + 1
    even = (N % 2 == 0)
    F = len(dk_list)
    kxx = np.arange(N)[:, None, None]
    kyy = (np.arange(nky) + ky_lo if ky_rows is None else np.asarray(ky_rows))[None, :, None]
    kz = np.arange(nz)[None, None, :]
    kx = np.where(kxx > m, kxx - N, kxx)
    ky = np.where(kyy > m, kyy - N, kyy)
    kx, ky, kz = np.broadcast_arrays(kx, ky, kz)
    zspec = (kz == 0) | ((kz == m) & even)
    skip = zspec & ((kx < 0) | (((kx == 0) | ((kx == m) & even)) & (ky < 0)))
    keep = ~skip
    k2 = kx * kx + ky * ky + kz * kz
    k = np.sqrt(k2.astype(np.float64))
    kidx = k.astype(np.int64)
    comp = (kx, ky, kz)
    kpar = comp[axis]
    others = [c for i, c in enumerate(comp) if i != axis]
    kper = np.sqrt((others[0] ** 2 + others[1] ** 2).astype(np.float64)).astype(np.int64)
    with np.errstate(invalid="ignore", divide="ignore"):
        mu = np.where(k == 0, 0.0, kpar / k)
    mu2 = mu * mu
    v1 = (3.0 * mu2 - 1.0) / 2.0
    v2 = (35.0 * mu2 * mu2 - 30.0 * mu2 + 3.0) / 8.0
    kpa = np.abs(kpar)
    i2 = (lay.kmax_par + 1) * kper + kpa
    in1 = keep & (k <= m)
    out = np.zeros(lay.total_words, dtype=np.int64)
    f64 = out.view(np.float64)
    n3, n1, n2 = lay.kmax + 1, lay.kmax_par + 1, lay.n2d

    def win(i, p):
        x = np.pi * i / N
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.where(i == 0, 1.0, (x / np.sin(x)) ** p)
    out[lay.Nm3D:lay.Nm3D + n3] = np.bincount(kidx[keep], minlength=n3)
    out[lay.Nm1D:lay.Nm1D + n1] = np.bincount(kpa[in1], minlength=n1)
    out[lay.Nm2D:lay.Nm2D + n2] = np.bincount(i2[keep], minlength=n2)
    f64[lay.k3D:lay.k3D + n3] = np.bincount(kidx[keep], weights=k[keep], minlength=n3)
    re, im = [], []
    for f in range(F):
        fac = (win(kx, mas_index[f]) * win(ky, mas_index[f]) * win(kz, mas_index[f])).astype(np.float32)
        d = np.asarray(dk_list[f]).astype(np.complex64)
        re.append((d.real.astype(np.float32) * fac).astype(np.float64))
        im.append((d.imag.astype(np.float32) * fac).astype(np.float64))
    X = F * (F - 1) // 2

    def put(base3, base1, base2, width, col, d2):
        for l, wgt in enumerate((None, v1, v2)):
            val = d2 if wgt is None else d2 * wgt
            f64[base3 + (np.arange(n3) * 3 + l) * width + col] += np.bincount(kidx[keep], weights=val[keep], minlength=n3)
        f64[base1 + np.arange(n1) * width + col] += np.bincount(kpa[in1], weights=d2[in1], minlength=n1)
        f64[base2 + np.arange(n2) * width + col] += np.bincount(i2[keep], weights=d2[keep], minlength=n2)
    for f in range(F):
        put(lay.Pk3D, lay.Pk1D, lay.Pk2D, F, f, re[f] ** 2 + im[f] ** 2)
    ix = 0
    for i in range(F):
        for j in range(i + 1, F):
            put(lay.PkX3D, lay.PkX1D, lay.PkX2D, X, ix, re[i] * re[j] + im[i] * im[j])
            ix += 1
    if want_phase:
        d2 = re[0] ** 2 + im[0] ** 2
        ph = np.arctan2(re[0], np.sqrt(d2)) ** 2
        f64[lay.phase:lay.phase + n3] = np.bincount(kidx[keep], weights=ph[keep], minlength=n3)
    return out


class NumpyOps:
    """Stand-ins for DeviceOps on CPU tensors (oracle deposit, numpy FFT, numpy binning)."""

    def __init__(self):
        from oracle import cpu
        from pylians3_b200 import _lib
        self.O, self._lib = cpu, _lib

    def base_plane(self, mas, pos, dims, BoxSize):
        inv = np.float32(dims) / np.float32(BoxSize)
        b = base_cell(mas, (pos[:, 0].numpy() * inv).astype(np.float32))
        return torch.from_numpy((b % dims).astype(np.int32))

    def deposit_slab(self, mas, pos, work, W, dims, BoxSize, x_origin, x_own, dropped):
        full = np.zeros((dims, dims, dims), np.float32)
        self.O.MA(pos.numpy(), full, BoxSize, mas, None if W is None else W.numpy())
        # like the kernel: global plane p lands in local plane (p - x_origin) mod dims, if that exists
        local = (np.arange(dims) - x_origin) % dims
        inside = local < work.shape[0]
        w = work.numpy()
        w[local[inside]] += full[inside]
        dropped += int(np.count_nonzero(full[~inside]))

    def add_inplace(self, out, inp):
        out += inp

    def sum_f64(self, x):
        return torch.tensor([float(np.sum(x.numpy(), dtype=np.float64))], dtype=torch.float64)

    def overdensity_(self, x, total, cells):
        a = x.numpy()
        a[...] = (a.astype(np.float64) / (float(total[0]) / cells)).astype(np.float32) - np.float32(1.0)

    def fft_yz(self, slab, dims):
        # complex128 here: the stand-in keeps full precision between the two FFT stages so that the
        # result equals one float64 3D transform; the rounding to complex64 happens in numpy_bin
        return torch.from_numpy(np.fft.rfftn(slab.numpy().astype(np.float64), axes=(1, 2)))

    def fft_x_(self, cols, dims):
        c = cols.numpy()
        c[...] = np.fft.fft(c, axis=0)
        return cols

    def bin(self, dk_list, mas_index, dims, axis, want_phase, ky_lo, ny_lo):
        from pylians3_b200.dist import mirrored_rows
        lay = self._lib.pk_layout(dims, len(dk_list))
        rows = mirrored_rows(dims, ky_lo, ny_lo)
        out = numpy_bin([d.numpy() for d in dk_list], mas_index, dims, axis, want_phase, 0, len(rows), lay, rows)
        return torch.from_numpy(out), lay


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, q):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from pylians3_b200 import dist as PD
        ops = NumpyOps()
        O = ops.O
        ctx = PD.SlabContext(N, BOX, device="cpu", ops=ops)
        pos, W = make_particles(42, 6 * N ** 3, True)
        W2 = (W * W).astype(np.float32)
        # every rank starts with an arbitrary half of the particles: routing must sort them out
        mine = slice(rank, None, world)
        res = {}
        slabs = {}
        for mas, w in (("PCS", None), ("CIC", W), ("NGP", None), ("TSC", W2)):
            slab = ctx.new_slab()
            ctx.MA(torch.from_numpy(pos[mine].copy()), slab, mas, None if w is None else torch.from_numpy(w[mine].copy()))
            slabs[mas] = slab
            ref = np.zeros((N, N, N), np.float32)
            O.MA(pos, ref, BOX, mas, w)
            x0, x1 = ctx.x_range
            res["ma_" + mas] = rel_err(slab.numpy(), ref[x0:x1], floor=float(ref.mean()))
        ctx.check_dropped()
        # delta and spectra
        refs = {}
        for mas, w in (("PCS", None), ("CIC", W)):
            ref = np.zeros((N, N, N), np.float32)
            O.MA(pos, ref, BOX, mas, w)
            ref /= np.mean(ref, dtype=np.float64)
            ref -= 1.0
            refs[mas] = ref
            ctx.overdensity_(slabs[mas])
            x0, x1 = ctx.x_range
            res["delta_" + mas] = float(np.max(np.abs(slabs[mas].numpy() - ref[x0:x1])) / np.max(np.abs(ref)))
        # reference spectra: the oracle loop on the SAME transform the stand-ins compute (float64 FFT
        # rounded to complex64), so that only the decomposition differs -> tight tolerances
        f64fft = lambda d: np.fft.rfftn(d.astype(np.float64)).astype(np.complex64)

        def close(a, b, floor=1e-5):
            # the slab path sums halo contributions in another order than the serial oracle, so delta
            # differs by float32 round-off (~1e-7 of its peak); deconvolved low-power bins see that as
            # an error relative to the spectrum's PEAK, hence the peak-relative floor
            pk = float(np.nanmax(np.abs(b)))
            ok = np.abs(a - b) <= 1e-4 * np.abs(b) + floor * pk
            return bool(np.all(ok | np.isnan(b))) and np.array_equal(np.isnan(a), np.isnan(b))
        for axis in (0, 1, 2):
            got = ctx.Pk(slabs["PCS"], axis, "PCS")
            want = O.Pk(None, BOX, axis, "PCS", 1, False, delta_k=f64fft(refs["PCS"]))
            for nm in ("Nmodes3D", "Nmodes1D", "Nmodes2D"):
                assert np.array_equal(getattr(got, nm), getattr(want, nm)), (axis, nm)
            for nm in ("k3D", "k1D", "kpar", "kper"):
                assert rel_err(getattr(got, nm), getattr(want, nm), 1e-300) < 1e-12, (axis, nm)
            for nm in ("Pk", "Pk1D", "Pk2D"):
                assert close(getattr(got, nm), getattr(want, nm)), (axis, nm)
            assert close(got.Pkphase, want.Pkphase, floor=3e-4), axis
        import contextlib
        import io
        gx = ctx.XPk([slabs["PCS"], slabs["CIC"]], 2, ["PCS", "CIC"])
        with contextlib.redirect_stdout(io.StringIO()):
            wx = O.XPk(None, BOX, 2, ["PCS", "CIC"], 1, delta_k=[f64fft(refs["PCS"]), f64fft(refs["CIC"])])
        for nm in ("Pk", "XPk", "Pk1D", "PkX1D", "Pk2D", "PkX2D"):
            assert getattr(gx, nm).shape == getattr(wx, nm).shape, nm
            assert close(getattr(gx, nm), getattr(wx, nm)), nm
        q.put((rank, "ok", res))
    except Exception:
        q.put((rank, "fail", traceback.format_exc()))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


@pytest.mark.parametrize("N", [16, 15])
def test_two_rank_pipeline_matches_single_process_oracle(oracle, N):
    from pylians3_b200 import build
    build.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, N, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, payload in out:
        assert status == "ok", "rank %d:\n%s" % (rank, payload)
        for k, v in payload.items():
            if k.startswith("ma_NGP"):
                assert v == 0.0, (k, v)
            elif k.startswith("ma_"):
                assert v < 1e-5, (k, v)
            else:
                assert v < 5e-6, (k, v)      # delta relative to its peak: float32 round-off of the halo sums


def test_numpy_bin_stand_in_agrees_with_the_c_oracle(oracle):
    """The numpy stand-in used above is itself pinned to the C restatement on a full window."""
    from pylians3_b200 import _lib, Pk_library as PKL
    N = 12
    rng = np.random.default_rng(3)
    fields = [rng.standard_normal((N, N, N)).astype(np.float32) for _ in range(2)]
    dks = [oracle.fft3d_r2c(f) for f in fields]
    for axis in (0, 1, 2):
        lay = _lib.pk_layout(N, 2)
        words = numpy_bin(dks, [2, 4], N, axis, False, 0, N, lay)
        raw = PKL.unpack_raw(words, lay)
        ref = oracle.bin_raw(dks, N, [2, 4], axis, BOX)
        for a, b in (("Nm3D", "Nm3D"), ("Nm1D", "Nm1D"), ("Nm2D", "Nm2D"), ("k1D", "k1D")):
            assert np.array_equal(raw[a], ref[b]), a
        for a in ("k3D", "Pk3D", "PkX3D", "Pk1D", "PkX1D", "Pk2D", "PkX2D"):
            assert np.max(np.abs(raw[a] - ref[a])) <= 1e-11 * max(1.0, float(np.max(np.abs(ref[a])))), a


def test_split_sizes():
    from pylians3_b200.dist import split_sizes
    assert split_sizes(10, 3) == ([4, 3, 3], [0, 4, 7, 10])
    assert split_sizes(4096, 8)[0] == [512] * 8


@pytest.mark.parametrize("dims,world", [(16, 2), (15, 2), (64, 8), (45, 4), (9, 5)])
def test_mirrored_ky_rows_partition_every_row_once(dims, world):
    """The mirrored ky distribution covers 0..dims-1 exactly once, keeps +-ky on one rank, and matches the
    row count the C ABI reports."""
    import ctypes
    from pylians3_b200 import _lib
    from pylians3_b200.dist import mirrored_rows, split_sizes, _row_runs
    lib = _lib.load()
    sizes, offs = split_sizes(dims // 2 + 1, world)
    seen = []
    for r in range(world):
        rows = mirrored_rows(dims, offs[r], sizes[r])
        first = ctypes.c_int(-1)
        assert lib.pyl_pk_mirrored_rows(dims, offs[r], sizes[r], ctypes.byref(first)) == len(rows)
        if len(rows) > sizes[r]:
            assert first.value == rows[sizes[r]]
        for ky in rows[:sizes[r]]:
            if ky != 0 and not (dims % 2 == 0 and ky == dims // 2):
                assert dims - ky in rows
        assert len(_row_runs(rows)) <= 2
        seen += rows
    assert sorted(seen) == list(range(dims))
