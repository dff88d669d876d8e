"""Slab-decomposed multi-GPU MA -> delta -> Pk/XPk (one process per GPU, torch.distributed).

The reference has no distributed path (one process, whole grid in host RAM).
This module shards the SAME computation over P ranks by x-slabs of the grid:

  route      each particle goes to the rank owning the x-plane of its FIRST stencil cell
             (all-to-all of positions/weights; skipped when the caller already holds its own)
  deposit    into [own planes + S-1 upward ghost planes]  (pyl_deposit_slab)
  halo       ghost planes travel to the next rank (ring send/recv, periodic) and are added
             into its first planes (pyl_add_inplace)
  delta      local float64 sum -> all-reduce -> n/<n> - 1 (pyl_overdensity_inplace)
  FFT        batched 2D r2c over (y,z) of the local planes (pyl_fft_slab_yz) -> all-to-all
             transpose (ky chunks) -> 1D c2c along x (pyl_fft_slab_x); no transpose back: the bin
             kernel only needs to know which ky rows it holds
  bin        pyl_pk_bin on the (dims, nky_local, dims/2+1) block -> all-reduce of the raw
             accumulators -> host finalisation (identical on every rank)

torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the plumbing; every arithmetic step is
a kernel of libpyl_b200.so reached through `DeviceOps`.  The CPU tests inject numpy stand-ins for
those kernels to exercise the communication skeleton with world_size 2 and no GPU.
"""
import ctypes
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _device as D
from . import _lib as L
from . import MAS_library as MASL
from . import Pk_library as PKL

_S = {"NGP": 1, "CIC": 2, "TSC": 3, "PCS": 4}


def split_sizes(n, parts):
    """Contiguous split of range(n) into `parts` pieces (first n%parts pieces get one more)."""
    base, extra = divmod(n, parts)
    sizes = [base + (1 if r < extra else 0) for r in range(parts)]
    offs = [0]
    for s in sizes:
        offs.append(offs[-1] + s)
    return sizes, offs


def mirrored_rows(dims, ky_lo, ny_lo):
    """Global ky index of the rows a rank stores: the lower rows ky_lo..ky_lo+ny_lo-1, then their mirrors
    dims-ky (those that are separate modes: ky != 0, ky != Nyquist) in ascending order -- the layout
    pyl_pk_bin_mirrored expects (include/pyl_b200.h)."""
    m = dims // 2
    lower = list(range(ky_lo, ky_lo + ny_lo))
    upper = sorted(dims - ky for ky in lower if ky != 0 and not (dims % 2 == 0 and ky == m))
    return lower + upper


def interleaved_row_order(ky_rows, rank):
    """Issue order of the ky rows in the peer-memory transpose of `rank`: row i of owner rank+1, row i of owner
    rank+2, ..., row i of `rank` itself, then rows i+1.  Every wave of CTAs then writes to ALL ranks, and no two
    senders start on the same receiver (an all-to-all in lock step puts every sender on one receiver's link)."""
    P = len(ky_rows)
    order = []
    for i in range(max(len(r) for r in ky_rows)):
        for d in range(1, P + 1):
            rows = ky_rows[(rank + d) % P]
            if i < len(rows):
                order.append(rows[i])
    return order


def _row_runs(rows):
    """Contiguous runs [a,b) of an ascending-by-pieces index list (at most two for mirrored rows)."""
    runs, start = [], 0
    for i in range(1, len(rows) + 1):
        if i == len(rows) or rows[i] != rows[i - 1] + 1:
            runs.append((rows[start], rows[i - 1] + 1))
            start = i
    return runs


class DeviceOps:
    """The compute kernels of the path, bound to libpyl_b200.so (device pointers, current stream)."""

    def __init__(self, device):
        self.lib = L.load()
        self.device = device

    def _s(self):
        return D.stream_ptr(self.device)

    def base_plane(self, mas, pos, dims, BoxSize):
        out = torch.empty(pos.shape[0], dtype=torch.int32, device=pos.device)
        L.check(self.lib.pyl_stencil_base_plane(L.MAS_IDS[mas], D.ptr(pos), pos.shape[0], dims,
                                                float(np.float32(BoxSize)), D.ptr(out), self._s()),
                "pyl_stencil_base_plane")
        return out

    def deposit_slab(self, mas, pos, work, W, dims, BoxSize, x_origin, x_own, dropped, count=None):
        """`count` (device uint32/int32 tensor of one element): the number of particles in pos/W is known on the
        device only (routed particles); pos.shape[0] is then the capacity of the arrays."""
        need = self.lib.pyl_deposit_slab_workspace_bytes(L.MAS_IDS[mas], pos.shape[0], dims, int(x_own))
        ws = D.workspace(need, work.device, "deposit")
        if count is None:
            L.check(self.lib.pyl_deposit_slab(L.MAS_IDS[mas], D.ptr(pos), D.ptr(work), D.ptr(W), pos.shape[0], dims,
                                              float(np.float32(BoxSize)), int(x_origin), int(x_own), work.shape[0],
                                              D.ptr(dropped), D.ptr(ws), need, self._s()), "pyl_deposit_slab")
        else:
            L.check(self.lib.pyl_deposit_slab_counted(L.MAS_IDS[mas], D.ptr(pos), D.ptr(work), D.ptr(W), pos.shape[0],
                                                      D.ptr(count), dims, float(np.float32(BoxSize)), int(x_origin),
                                                      int(x_own), work.shape[0], D.ptr(dropped), D.ptr(ws), need,
                                                      self._s()), "pyl_deposit_slab_counted")

    def route_scatter(self, mas, pos, W, dims, BoxSize, x_offs, peer_ptrs, capacity, lost):
        """Particles -> receive blocks of the ranks owning their first stencil plane (peer stores over NVLink)."""
        P = len(peer_ptrs)
        offs = (ctypes.c_int * (P + 1))(*[int(o) for o in x_offs])
        ptrs = (ctypes.c_void_p * P)(*[int(p) for p in peer_ptrs])
        L.check(self.lib.pyl_route_scatter(L.MAS_IDS[mas], D.ptr(pos), D.ptr(W), pos.shape[0], dims,
                                           float(np.float32(BoxSize)), P, offs, ptrs, int(capacity), D.ptr(lost),
                                           self._s()), "pyl_route_scatter")

    def add_inplace(self, out, inp):
        L.check(self.lib.pyl_add_inplace(D.ptr(out), D.ptr(inp), out.numel(), self._s()), "pyl_add_inplace")

    def sum_f64(self, x):
        out = torch.empty(1, dtype=torch.float64, device=x.device)
        L.check(self.lib.pyl_sum_f64(D.ptr(x), x.numel(), D.ptr(out), self._s()), "pyl_sum_f64")
        return out

    def overdensity_(self, x, total, cells):
        L.check(self.lib.pyl_overdensity_inplace(D.ptr(x), x.numel(), D.ptr(total), float(cells), self._s()),
                "pyl_overdensity_inplace")

    def fft_yz(self, slab, dims, out=None):
        """Batched 2D r2c over (y,z) of the planes of `slab`; `out` (nx, dims, nz) complex64 or a fresh tensor."""
        nx = slab.shape[0]
        if out is None:
            out = torch.empty((nx, dims, dims // 2 + 1), dtype=torch.complex64, device=slab.device)
        need = self.lib.pyl_fft_slab_workspace_bytes(dims, nx, 0)
        if need == ctypes.c_size_t(-1).value:
            L.check(-3, "pyl_fft_slab_workspace_bytes")
        ws = D.workspace(need, slab.device, "fft")
        L.check(self.lib.pyl_fft_slab_yz(D.ptr(slab), D.ptr(out), dims, nx, D.ptr(ws), need, self._s()),
                "pyl_fft_slab_yz")
        return out

    def fft_x_(self, cols, dims):
        nky = cols.shape[1]
        need = self.lib.pyl_fft_slab_workspace_bytes(dims, 0, nky)
        if need == ctypes.c_size_t(-1).value:
            L.check(-3, "pyl_fft_slab_workspace_bytes")
        ws = D.workspace(need, cols.device, "fft")
        L.check(self.lib.pyl_fft_slab_x(D.ptr(cols), dims, nky, D.ptr(ws), need, self._s()), "pyl_fft_slab_x")
        return cols

    def transpose_scatter(self, a, peer_ptrs, nky_of_rank, ky_owner, ky_row, dims, x0, ky_major=False, ky_order=None):
        """Rows of the local (nx, dims, nz) stage-1 output -> receive buffers of their owner ranks (peer stores).
        ky_major: the receive buffers are (nky, dims, nz) instead of (dims, nky, nz).  ky_order (int32 CUDA tensor
        [dims], a permutation): the order in which the rows are issued (see interleaved_row_order)."""
        P = len(peer_ptrs)
        ptrs = (ctypes.c_void_p * P)(*[int(p) for p in peer_ptrs])
        nky = (ctypes.c_int * P)(*[int(n) for n in nky_of_rank])
        fn = self.lib.pyl_transpose_scatter_kymajor if ky_major else self.lib.pyl_transpose_scatter
        L.check(fn(D.ptr(a), ptrs, nky, D.ptr(ky_owner), D.ptr(ky_row), D.ptr(ky_order) if ky_order is not None else None,
                   dims, a.shape[0], int(x0), P, self._s()), "pyl_transpose_scatter")

    def fft_x_kymajor_(self, cols, dims):
        """In-place 1D transforms along x of a (nky, dims, nz) complex64 array (x is the MIDDLE axis)."""
        need = self.lib.pyl_fft_slab_x_kymajor_workspace_bytes(dims)
        if need == ctypes.c_size_t(-1).value:
            L.check(-3, "pyl_fft_slab_x_kymajor_workspace_bytes")
        ws = D.workspace(need, cols.device, "fft")
        L.check(self.lib.pyl_fft_slab_x_kymajor(D.ptr(cols), dims, cols.shape[0], D.ptr(ws), need, self._s()),
                "pyl_fft_slab_x_kymajor")
        return cols

    def bin(self, dk_list, mas_index, dims, axis, want_phase, ky_lo, ny_lo):
        """Bins the mirrored slab (rows as in mirrored_rows(dims, ky_lo, ny_lo)).  A field whose FIRST axis is not
        dims long is laid out (nky, dims, nz) -- what the peer-memory transpose produces."""
        ky_major = dk_list[0].shape[0] != dims or getattr(dk_list[0], "_pyl_ky_major", False)
        return PKL.bin_device(dk_list, mas_index, dims, axis, want_phase, ky_lo, ny_lo, mirrored=True,
                              flags=L.PK_KY_MAJOR if ky_major else 0)


class _Result(PKL.K2D):
    pass


class SlabContext:
    """Per-rank state of one slab-decomposed grid: plane ranges, neighbours, scratch buffers."""

    def __init__(self, dims, BoxSize, group=None, device=None, ops=None):
        if not dist.is_initialized():
            raise RuntimeError("SlabContext needs torch.distributed to be initialised")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.dims = int(dims)
        self.BoxSize = float(BoxSize)
        self.nz = self.dims // 2 + 1
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.ops = ops if ops is not None else DeviceOps(self.device)
        self.x_sizes, self.x_offs = split_sizes(self.dims, self.world)
        # ky is distributed in MIRRORED pairs: rank r owns |ky| in [ky_lo, ky_lo+ny_lo) of the half range
        # 0..N/2 together with the rows N-ky, so that the bin kernel can share geometry between +-ky like on a
        # whole grid (pyl_pk_bin_mirrored).  ky_rows[r] = global ky index of every row rank r stores, in order.
        self.ylo_sizes, self.ylo_offs = split_sizes(self.dims // 2 + 1, self.world)
        self.ky_rows = [mirrored_rows(self.dims, self.ylo_offs[r], self.ylo_sizes[r]) for r in range(self.world)]
        self.x_range = (self.x_offs[self.rank], self.x_offs[self.rank + 1])
        self.ky_lo, self.ny_lo = self.ylo_offs[self.rank], self.ylo_sizes[self.rank]
        self.nx = self.x_sizes[self.rank]
        self.nky = len(self.ky_rows[self.rank])
        if min(self.ylo_sizes) < 1:
            raise ValueError("more ranks than ky rows (dims=%d over %d ranks)" % (dims, self.world))
        if min(self.x_sizes) < 3:
            raise ValueError("slabs thinner than the PCS ghost width (dims=%d over %d ranks)" % (dims, self.world))
        self.dropped = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._plane_owner = None
        self._scratch = {}
        # transpose over peer memory (one kernel storing into the owners' receive buffers) where symmetric
        # memory is available: CUDA, one node.  PYL_TRANSPOSE=nccl keeps the pack + all-to-all path.
        self._peer = None
        self._peer_slots = {}
        self._route = None
        self._route_mode = os.environ.get("PYL_ROUTE", "peer")      # "nccl": argsort + all_to_all_single
        if self.device.type == "cuda" and self.world > 1 and os.environ.get("PYL_TRANSPOSE", "peer") == "peer":
            self._peer = self._setup_peer()

    def _setup_peer(self):
        """Peer-memory plumbing (symmetric allocations mapped into every rank), or None -> NCCL all-to-all.  The
        probe allocates and maps a small symmetric buffer right here, so that hosts without P2P, multi-node groups or
        more ranks than the transpose kernel takes fall back now instead of failing inside fft(); the ranks agree
        on the outcome (a rendezvous is collective)."""
        state, why = None, ""
        try:
            import warnings
            import torch.distributed._symmetric_memory as symm
            grp = self.group if self.group is not None else dist.group.WORLD
            if self.world > 16:
                raise RuntimeError("pyl_transpose_scatter takes at most 16 ranks")
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                try:
                    symm.enable_symm_mem_for_group(grp.group_name)
                except Exception:
                    pass
            probe = symm.empty(1024, dtype=torch.uint8, device=self.device)
            symm.rendezvous(probe, grp)
            owner = torch.empty(self.dims, dtype=torch.int32)
            row = torch.empty(self.dims, dtype=torch.int32)
            for r, rows in enumerate(self.ky_rows):
                idx = torch.tensor(rows, dtype=torch.long)
                owner[idx] = r
                row[idx] = torch.arange(len(rows), dtype=torch.int32)
            order = torch.tensor(interleaved_row_order(self.ky_rows, self.rank), dtype=torch.int32)
            state = {"symm": symm, "group": grp, "owner": owner.to(self.device), "row": row.to(self.device),
                     "order": order.to(self.device), "nky": [len(r) for r in self.ky_rows]}
        except Exception as e:                                   # symmetric memory unusable: NCCL all-to-all
            why = str(e)
        ok = torch.tensor([1 if state is not None else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 0:
            if self.rank == 0:
                print("pylians3_b200.dist: peer-memory transpose unavailable (%s); using NCCL all-to-all" % (why or "on another rank"))
            return None
        return state

    def _peer_slot(self, slot):
        """Symmetric receive buffer number `slot` (one per field that must stay alive at the same time)."""
        s = self._peer_slots.get(slot)
        if s is None:
            symm = self._peer["symm"]
            n = self.dims * max(self._peer["nky"]) * self.nz
            buf = symm.empty(n, dtype=torch.complex64, device=self.device)
            hdl = symm.rendezvous(buf, self._peer["group"])
            s = self._peer_slots[slot] = (buf, hdl, [int(p) for p in hdl.buffer_ptrs])
        return s

    def _side_stream(self):
        """Stream of the transposes.  HIGH priority: its CTAs take the SM slots the 2D-FFT kernels of the main
        stream free up, so the NVLink traffic runs at its own pace instead of waiting behind the next FFT batch
        (alone the kernel reaches 0.8 of the link rate; queued behind cuFFT at equal priority it got 0.45)."""
        if getattr(self, "_side", None) is None:
            prio = int(os.environ.get("PYL_SIDE_PRIORITY", "-1"))
            self._side = torch.cuda.Stream(device=self.device, priority=prio)
        return self._side

    def _buf(self, name, shape, dtype):
        """Persistent scratch buffer per (name, shape, dtype): the multi-GB work/transpose buffers are
        allocated once per context instead of once per step (allocator churn showed up as several ms of
        run-to-run noise in the transpose)."""
        key = (name, tuple(shape), dtype)
        b = self._scratch.get(key)
        if b is None:
            for k in [k for k in self._scratch if k[0] == name]:
                del self._scratch[k]
            b = self._scratch[key] = torch.empty(shape, dtype=dtype, device=self.device)
        return b

    # ---- helpers ----------------------------------------------------------------------------------
    FFT_BATCH_PLANES = 16       # x-planes per 2D-FFT / transpose batch of the pipelined slab FFT
    # From this grid size on the transposed spectrum is laid out (nky, N, nz), x in the middle: the x transforms then
    # run per ky plane instead of striding across the whole buffer.  Measured per GPU (profiles/r2_config5_pieces.md):
    # 4096^3 over 8 ranks 31 vs 83 ms, 2048^3 4.2 vs 5.0 ms; at 1024^3 the per-plane calls are launch-bound and the
    # x-outermost layout wins (0.33 vs 1.06 ms).
    KY_MAJOR_MIN_DIMS = 2048
    GHOST_PLANES = 3            # PCS: the stencil reaches three planes above the particle's first plane

    def new_slab(self):
        """(nx_local, dims, dims) float32 zeros.  The allocation carries GHOST_PLANES more planes behind the slab:
        MA deposits straight into [own planes | ghost planes] and ships the ghost planes to the next rank, so no
        slab-sized work buffer exists (at 4096^3 over 8 GPUs a slab is 34 GB)."""
        store = torch.zeros((self.nx + self.GHOST_PLANES, self.dims, self.dims), dtype=torch.float32,
                            device=self.device)
        return store[:self.nx]

    def _with_ghosts(self, slab, ghosts):
        """The (nx + ghosts, dims, dims) view over a slab made by new_slab(), or None for foreign tensors."""
        if slab.device.type != "cuda" or not slab.is_contiguous() or slab.storage_offset() != 0:
            return None
        n2 = self.dims * self.dims
        if slab.untyped_storage().nbytes() < (self.nx + ghosts) * n2 * 4:
            return None
        return torch.empty(0, dtype=torch.float32, device=slab.device).set_(
            slab.untyped_storage(), 0, (self.nx + ghosts, self.dims, self.dims), (n2, self.dims, 1))

    def plane_owner(self):
        if self._plane_owner is None:
            own = torch.empty(self.dims, dtype=torch.int64)
            for r in range(self.world):
                own[self.x_offs[r]:self.x_offs[r + 1]] = r
            self._plane_owner = own.to(self.device)
        return self._plane_owner

    def _all_to_all(self, recv, send, recv_sizes, send_sizes):
        dist.all_to_all_single(recv, send, output_split_sizes=recv_sizes, input_split_sizes=send_sizes,
                               group=self.group)

    # ---- particles -> owner ranks ----------------------------------------------------------------
    def route(self, pos, MAS, W=None):
        """Send every particle to the rank that owns the x-plane of its first stencil cell."""
        plane = self.ops.base_plane(MAS, pos, self.dims, self.BoxSize)
        owner = self.plane_owner()[plane.long()]
        order = torch.argsort(owner, stable=True)
        send_counts = torch.bincount(owner, minlength=self.world)
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=self.group)
        sc = [int(c) for c in send_counts.cpu()]
        rc = [int(c) for c in recv_counts.cpu()]
        pos_s = pos[order].contiguous()
        pos_r = torch.empty((sum(rc), 3), dtype=torch.float32, device=pos.device)
        self._all_to_all(pos_r.view(-1), pos_s.view(-1), [3 * c for c in rc], [3 * c for c in sc])
        W_r = None
        if W is not None:
            W_s = W[order].contiguous()
            W_r = torch.empty(sum(rc), dtype=torch.float32, device=pos.device)
            self._all_to_all(W_r, W_s, rc, sc)
        return pos_r, W_r

    def route_peer(self, pos, MAS, W=None):
        """route() as ONE kernel over peer memory (pyl_route_scatter): returns (pos_r, W_r, count) where pos_r / W_r
        are views of this rank's symmetric receive block (capacity rows) and `count` is the device word holding the
        number of particles received.  No host synchronisation after the first call (which agrees on the capacity)."""
        symm = self._peer["symm"]
        n = int(pos.shape[0])
        R = getattr(self, "_route", None)
        if R is None:
            t = torch.tensor([n], dtype=torch.int64, device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            cap = int(1.25 * int(t.item())) + 65536
            nbytes = int(self.ops.lib.pyl_route_block_bytes(cap))
            buf = symm.empty(nbytes, dtype=torch.uint8, device=self.device)
            hdl = symm.rendezvous(buf, self._peer["group"])
            R = self._route = {"cap": cap, "buf": buf, "hdl": hdl, "ptrs": [int(p) for p in hdl.buffer_ptrs],
                               "lost": torch.zeros(1, dtype=torch.int64, device=self.device),
                               "w_off": (256 + cap * 12 + 255) // 256 * 256}
        cap, buf, hdl = R["cap"], R["buf"], R["hdl"]
        if n > cap:
            raise ValueError("route_peer: %d particles exceed the receive capacity %d agreed on the first call; "
                             "make a new SlabContext for larger inputs" % (n, cap))
        buf[:256].zero_()                                   # this rank's cursor
        hdl.barrier(channel=1)                              # every cursor is zero, every block is free
        self.ops.route_scatter(MAS, pos, W, self.dims, self.BoxSize, self.x_offs, R["ptrs"], cap, R["lost"])
        hdl.barrier(channel=1)                              # every particle has landed
        pos_r = buf[256:256 + cap * 12].view(torch.float32).view(cap, 3)
        W_r = None if W is None else buf[R["w_off"]:R["w_off"] + cap * 4].view(torch.float32)
        return pos_r, W_r, buf[:4].view(torch.int32)

    # ---- mass assignment -------------------------------------------------------------------------
    def MA(self, pos, slab, MAS="CIC", W=None, routed=False, count=None):
        """slab (nx_local, dims, dims) += deposit of this rank's share; ghost planes go to the next rank.
        routed=True: the particles are already on their owner ranks (`count`: their number as a device word, when
        they are the output of route_peer)."""
        if MAS not in _S:
            raise ValueError("option not valid!!!")
        host = isinstance(pos, torch.Tensor) and not pos.is_cuda
        if not routed and self.world > 1:
            if host:
                pos, W = pos.to(self.device, non_blocking=True), (None if W is None else W.to(self.device, non_blocking=True))
                host = False
            if self._peer is not None and self._route_mode == "peer":
                pos, W, count = self.route_peer(pos, MAS, W)
            else:
                pos, W = self.route(pos, MAS, W)
        ghosts = _S[MAS] - 1
        x0 = self.x_range[0]

        chunk = MASL.stream_chunk(self.dims) // self.world

        def deposit(target):
            if host and pos.shape[0] >= 2 * chunk:
                # routed particles still in (pinned) host memory: stream them, one slab deposit per chunk
                MASL.stream_host_chunks(pos, W, self.device, chunk,
                                        lambda p, w: self.ops.deposit_slab(MAS, p, target, w, self.dims, self.BoxSize,
                                                                           x0, self.nx, self.dropped))
            elif host:
                self.ops.deposit_slab(MAS, pos.to(self.device, non_blocking=True), target,
                                      None if W is None else W.to(self.device, non_blocking=True), self.dims,
                                      self.BoxSize, x0, self.nx, self.dropped)
            elif count is not None:
                self.ops.deposit_slab(MAS, pos, target, W, self.dims, self.BoxSize, x0, self.nx, self.dropped,
                                      count=count)
            else:
                self.ops.deposit_slab(MAS, pos, target, W, self.dims, self.BoxSize, x0, self.nx, self.dropped)

        if ghosts == 0 or self.world == 1:
            # one rank: the window is the whole periodic grid (x_own = x_planes = dims), nothing to exchange
            deposit(slab)
            return
        work = self._with_ghosts(slab, ghosts)
        own_slab = work is not None
        if own_slab:
            work[self.nx:].zero_()                       # the ghost planes behind the slab
        else:                                            # a slab not made by new_slab(): side buffer, then add
            work = self._buf("work", (self.nx + ghosts, self.dims, self.dims), torch.float32)
            work.zero_()
        deposit(work)
        halo_out = work[self.nx:]                        # planes x1 .. x1+ghosts-1 belong to the next rank
        halo_in = self._buf("halo_in", halo_out.shape, torch.float32)
        nxt, prv = (self.rank + 1) % self.world, (self.rank - 1) % self.world
        ops = [dist.P2POp(dist.isend, halo_out, nxt, self.group), dist.P2POp(dist.irecv, halo_in, prv, self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        if not own_slab:
            self.ops.add_inplace(slab, work[:self.nx])
        self.ops.add_inplace(slab[:ghosts], halo_in)

    def check_dropped(self):
        """Raise if any stencil contribution fell outside a rank's planes (mis-routed particles)."""
        d = self.dropped.clone()
        if getattr(self, "_route", None) is not None:
            d += self._route["lost"]
        dist.all_reduce(d, group=self.group)
        n = int(d.item())
        if n:
            raise RuntimeError("%d stencil contributions fell outside their slab: particles were not routed" % n)

    # ---- delta = n/<n> - 1 -----------------------------------------------------------------------
    def overdensity_(self, slab):
        total = self.ops.sum_f64(slab)
        dist.all_reduce(total, group=self.group)
        self.ops.overdensity_(slab, total, float(self.dims) ** 3)
        return slab

    def prebias_(self, slab, particles, W=None):
        """field.prebias_ for a slab: every rank fills its planes with the SAME -c (a rank-dependent constant
        would be a step function along x), c = all-reduced sum of the weights / dims^3.  `particles`
        and `W` are this rank's share before routing.  Returns c (CUDA float64[1]) for Pk(..., offset=c).
        CPU stand-ins (tests) zero the slab and return 0."""
        if self.device.type != "cuda":
            slab.zero_()
            return torch.zeros(1, dtype=torch.float64)
        from . import field
        scale = field.weight_total(particles, W, self.device)
        dist.all_reduce(scale, group=self.group)
        return field.prebias_(slab, None, cells=float(self.dims) ** 3, scale=scale)

    # ---- distributed r2c: (nx_local, N, N) real -> (N, nky_local, nz) complex ---------------------
    def fft(self, slab, slot=0, marks=None):
        """(nx_local, N, N) real -> (N, nky_local, nz) complex.  With the peer-memory transpose the result lives
        in symmetric receive buffer `slot` and stays valid until the next fft() with the same slot.  `marks`
        (a dict) receives CUDA events for bench.py: "t0"/"t1" bracket the pipelined 2D-FFT + transpose region on the
        main stream, "pairs" brackets every transpose kernel on its side stream (the NVLink figure)."""
        N, nz, P = self.dims, self.nz, self.world

        def mark(key=None, stream=None):
            if marks is None or self.device.type != "cuda":
                return None
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream) if stream is not None else e.record()
            if key is not None:
                marks[key] = e
            return e

        if P == 1:
            return self.ops.fft_x_(self.ops.fft_yz(slab, N), N)
        if self._peer is not None:
            # Batches of x-planes: the 2D transforms of batch b+1 (main stream) run while batch b travels to the
            # owners of its ky rows (side stream, one kernel storing over NVLink).  Two batch buffers replace the
            # slab-sized stage-1 array.
            buf, hdl, ptrs = self._peer_slot(slot)
            ky_major = N >= self.KY_MAJOR_MIN_DIMS
            nb = max(1, min(self.nx, self.FFT_BATCH_PLANES))
            ring = self._buf("fft_ring", (2, nb, N, nz), torch.complex64)
            main = torch.cuda.current_stream(self.device)
            side = self._side_stream()
            hdl.barrier(channel=0)                                     # every rank is done with this slot
            mark("t0")
            free = [None, None]
            for k, b0 in enumerate(range(0, self.nx, nb)):
                b1, j = min(b0 + nb, self.nx), k & 1
                if free[j] is not None:
                    main.wait_event(free[j])                           # the scatter that read this buffer is done
                a = self.ops.fft_yz(slab[b0:b1], N, out=ring[j, :b1 - b0])
                ready = torch.cuda.Event()
                ready.record(main)
                with torch.cuda.stream(side):
                    side.wait_event(ready)
                    ea = mark(stream=side)
                    self.ops.transpose_scatter(a, ptrs, self._peer["nky"], self._peer["owner"], self._peer["row"], N,
                                               self.x_range[0] + b0, ky_major=ky_major,
                                               ky_order=None if os.environ.get("PYL_TRANSPOSE_ORDER") == "plain"
                                               else self._peer["order"])
                    eb = mark(stream=side)
                    if ea is not None:
                        marks.setdefault("pairs", []).append((ea, eb))
                    free[j] = torch.cuda.Event()
                    free[j].record(side)
            main.wait_stream(side)
            hdl.barrier(channel=0)                                     # every row has landed
            mark("t1")
            if not ky_major:
                return self.ops.fft_x_(buf[:N * self.nky * nz].view(N, self.nky, nz), N)
            # (nky, N, nz): x in the middle, so that the x transforms stay inside one plane per ky row
            out = self.ops.fft_x_kymajor_(buf[:N * self.nky * nz].view(self.nky, N, nz), N)
            out._pyl_ky_major = True                                   # nky == N cannot happen for P > 1, but be explicit
            return out
        a = self.ops.fft_yz(slab, N)                                   # (nx, N, nz)
        send = self._buf("send", (self.nx * N * nz,), a.dtype)
        send_sizes, off = [], 0                                        # (complex64 on the GPU path)
        for r in range(P):                                             # pack: the ky rows of rank r, every local plane
            rows = self.ky_rows[r]
            n = self.nx * len(rows) * nz
            dst = send[off:off + n].view(self.nx, len(rows), nz)
            j = 0
            for k0, k1 in _row_runs(rows):
                dst[:, j:j + (k1 - k0), :].copy_(a[:, k0:k1, :])
                j += k1 - k0
            send_sizes.append(2 * n)
            off += n
        del a
        recv_sizes = [2 * self.x_sizes[s] * self.nky * nz for s in range(P)]
        # a fresh buffer: the caller keeps the spectrum (XPk holds several at once)
        recv = torch.empty((N, self.nky, nz), dtype=send.dtype, device=slab.device)
        # blocks arrive ordered by source rank = ascending x: the receive buffer IS (N, nky, nz)
        self._all_to_all(torch.view_as_real(recv).view(-1), torch.view_as_real(send).view(-1), recv_sizes, send_sizes)
        return self.ops.fft_x_(recv, N)

    # ---- spectra -----------------------------------------------------------------------------------
    def _reduce(self, out, lay, extra=0):
        """All-reduce the raw accumulators.  Counts (uint64 words) become float64 first -- exact below
        2^53 -- so that one float64 SUM covers the whole buffer."""
        f64 = out.view(torch.float64)
        if out.is_cuda:
            with torch.cuda.device(out.device):
                L.check(L.load().pyl_pk_counts_to_f64(D.ptr(out), self.dims, int(lay.fields), D.stream_ptr(out.device)),
                        "pyl_pk_counts_to_f64")
        else:
            for off, n in ((lay.Nm3D, lay.kmax + 1), (lay.Nm1D, lay.kmax_par + 1), (lay.Nm2D, lay.n2d)):
                f64[off:off + n] = out[off:off + n].to(torch.float64)
        # (the kpar/kper scratch behind the accumulators is not reduced, but for `extra` words the caller put there)
        dist.all_reduce(f64[:lay.total_words + extra], group=self.group)
        return f64

    def _raw(self, dk_list, mas_index, axis, want_phase):
        out, lay = self.ops.bin(dk_list, mas_index, self.dims, axis, want_phase, self.ky_lo, self.ny_lo)
        f64 = D.to_host_numpy(self._reduce(out, lay))
        words = f64.view(np.int64).copy()
        for off, n in ((lay.Nm3D, lay.kmax + 1), (lay.Nm1D, lay.kmax_par + 1), (lay.Nm2D, lay.n2d)):
            words[off:off + n] = np.rint(f64[off:off + n]).astype(np.int64)
        return PKL.unpack_raw(words, lay)

    def _spectra(self, dk_list, mas_index, axis, want_phase, density=False, offset=None, marks=None):
        """bin -> all-reduce -> finalisation.  On GPUs the finalisation runs on the device (pyl_pk_finalize)
        and only the finished arrays cross PCIe; the CPU stand-ins of the tests finalise on the host.
        density=True: the slabs held densities n; the rank that holds k = 0 takes dims^3 <n> from the DC modes,
        the values ride the all-reduce behind the accumulators, and the sums are scaled to those of n/<n> - 1.
        marks (a dict, bench.py): CUDA events "binned" (bin kernels done) and "reduced" (all-reduce done)."""
        if self.device.type != "cuda":
            if density:
                raise NotImplementedError("density=True needs the CUDA path")
            return PKL._finalize(self._raw(dk_list, mas_index, axis, want_phase), self.BoxSize, self.dims)

        def mark(key):
            if marks is not None:
                marks[key] = torch.cuda.Event(enable_timing=True)
                marks[key].record()
        if density and offset is None:
            raise ValueError("density=True needs offset= (c = prebias_(slab, particles, W) before the deposit)")
        dc = PKL.take_dc(dk_list, holds_dc=(self.ky_lo == 0 and self.ny_lo > 0)) if density else None
        out, lay = self.ops.bin(dk_list, mas_index, self.dims, axis, want_phase, self.ky_lo, self.ny_lo)
        extra = 0
        if density:
            extra = len(dk_list)                   # (kpar/kper scratch: finalisation overwrites it afterwards)
            out.view(torch.float64)[lay.total_words:lay.total_words + extra].copy_(dc)
        mark("binned")
        f64 = self._reduce(out, lay, extra)
        mark("reduced")
        if density:
            PKL.density_scale_(out, lay, self.dims, f64[lay.total_words:lay.total_words + extra].clone(),
                               PKL._offsets(offset, extra, out.device))
        return PKL.finalize_device(f64.view(torch.int64), lay, self.BoxSize, self.dims, counts_are_f64=True)

    def Pk(self, slab, axis=2, MAS="CIC", density=False, offset=None):
        """Pk_library.Pk of the slab-distributed field; every rank gets the full result.
        density=True: the slab holds the density n, the spectrum is that of n/<n> - 1 (no overdensity_ pass);
        offset=c: it holds n - c (c = prebias_(slab, ...) before the deposit)."""
        dk = self.fft(slab)
        o = self._spectra([dk], [PKL.MAS_function(MAS)], axis, True, density, offset)
        r = _Result()
        r.k1D, r.Pk1D, r.Nmodes1D = o["k1D"], o["Pk1D"][:, 0], o["Nmodes1D"]
        r._kgrid, r.Pk2D, r.Nmodes2D = o["kgrid"], o["Pk2D"][:, 0], o["Nmodes2D"]
        r.k3D, r.Nmodes3D = o["k3D"], o["Nmodes3D"]
        r.Pk, r.Pkphase = np.ascontiguousarray(o["Pk"][:, :, 0]), o["Pkphase"]
        return r

    def XPk(self, slabs, axis=2, MAS=None, density=False, offset=None):
        """Pk_library.XPk of several slab-distributed fields (<= L.MAX_FIELDS per launch)."""
        if MAS is None or len(MAS) != len(slabs):
            raise TypeError("MAS must be a list with one scheme per field")
        if len(slabs) > L.MAX_FIELDS:
            raise ValueError("the distributed XPk bins at most %d fields per call" % L.MAX_FIELDS)
        dk = [self.fft(s, slot=i) for i, s in enumerate(slabs)]
        o = self._spectra(dk, [PKL.MAS_function(m) for m in MAS], axis, False, density, offset)
        r = _Result()
        r.k1D, r.Nmodes1D, r.Pk1D, r.PkX1D = o["k1D"], o["Nmodes1D"], o["Pk1D"], o["PkX1D"]
        r._kgrid, r.Nmodes2D, r.Pk2D, r.PkX2D = o["kgrid"], o["Nmodes2D"], o["Pk2D"], o["PkX2D"]
        r.k3D, r.Nmodes3D, r.Pk, r.XPk = o["k3D"], o["Nmodes3D"], o["Pk"], o["XPk"]
        return r
