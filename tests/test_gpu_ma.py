"""GPU parity tests of the mass-assignment stage: CUDA path (through the C ABI) vs the CPU oracle
and vs the golden vectors of the compiled reference.

Tolerances (BASELINE.json north_star): NGP cell counts bit-exact; CIC/TSC/PCS float32 grids within
1e-5 relative per cell, relative to max(|cell|, mean)."""
import ctypes

import numpy as np
import pytest

from conftest import BOX, make_particles, rel_err

pytestmark = pytest.mark.gpu
MAS = ("NGP", "CIC", "TSC", "PCS")
MODES = ("atomic", "tiled", "deterministic")
TOL = 1e-5


@pytest.fixture(scope="module")
def env():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from pylians3_b200 import MAS_library as MASL, _lib
    _lib.load()
    return torch, MASL, _lib


def cell_err(got, ref):
    return rel_err(got, ref, floor=max(float(np.mean(np.abs(ref))), 1e-30))


@pytest.mark.parametrize("N", [16, 9])
@pytest.mark.parametrize("clu", ["uni", "clu"])
@pytest.mark.parametrize("mas", MAS)
@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("mode", MODES)
def test_ma_vs_reference_golden(env, ma_golden, N, clu, mas, weighted, mode):
    torch, MASL, _ = env
    tag = "N%d_%s" % (N, clu)
    pos, W = ma_golden[tag + "_pos"], ma_golden[tag + "_W"]
    w = W if weighted else None
    for nd in (3, 2):
        g = np.zeros((N,) * nd, np.float32)
        p = pos if nd == 3 else np.ascontiguousarray(pos[:, :2])
        MASL.MA(p, g, BOX, mas, w, mode=mode)
        ref = ma_golden["%s_%s_%s_%dD" % (tag, mas, "W" if weighted else "U", nd)]
        if mas == "NGP" and not weighted:
            assert np.array_equal(g, ref), "NGP counts must be bit-exact"
        else:
            assert cell_err(g, ref) < TOL


@pytest.mark.parametrize("N,npart,clustered", [(64, 300000, False), (64, 300000, True), (128, 1500000, True),
                                               (33, 100000, True), (256, 2000000, False)])
@pytest.mark.parametrize("mode", MODES)
def test_ma_vs_oracle_medium(env, oracle, N, npart, clustered, mode):
    torch, MASL, _ = env
    pos, W = make_particles(N * 7 + int(clustered), npart, clustered)
    for mas in MAS:
        for w in (None, W):
            ref = np.zeros((N, N, N), np.float32)
            oracle.MA(pos, ref, BOX, mas, w)
            got = np.zeros((N, N, N), np.float32)
            MASL.MA(pos, got, BOX, mas, w, mode=mode)
            if mas == "NGP" and w is None:
                assert np.array_equal(got, ref)
            else:
                assert cell_err(got, ref) < TOL, (mas, w is not None)
            # mass conservation, library/tests/test.py:30
            tot = float(np.sum(W, dtype=np.float64)) if w is not None else float(npart)
            assert abs(np.sum(got, dtype=np.float64) / tot - 1.0) < 1e-5


@pytest.mark.parametrize("mode", MODES)
def test_ma_extreme_clustering_accumulation_noise(env, oracle, mode):
    """Thousands of particles per cell: a float32 sum of n terms depends on the summation order at the
    level sqrt(n)*2^-24 (the reference itself changes by that much if the particles are shuffled), so
    the per-cell bound is 1e-5 * max(1, sqrt(n_cell/64)); mass is still conserved to 1e-5."""
    torch, MASL, _ = env
    N = 128
    pos, W = make_particles(901, 1500000, True, sigma=0.01)
    counts = np.zeros((N, N, N), np.float32)
    oracle.MA(pos, counts, BOX, "NGP")
    for mas in ("CIC", "PCS"):
        ref = np.zeros((N, N, N), np.float32)
        oracle.MA(pos, ref, BOX, mas)
        got = np.zeros((N, N, N), np.float32)
        MASL.MA(pos, got, BOX, mas, mode=mode)
        from scipy.ndimage import maximum_filter
        n_cell = maximum_filter(counts, size=5, mode="wrap")
        bound = TOL * np.maximum(1.0, np.sqrt(n_cell / 64.0)) * np.maximum(np.abs(ref), ref.mean())
        assert np.all(np.abs(got - ref) <= bound), mas
        assert abs(np.sum(got, dtype=np.float64) / len(pos) - 1.0) < 1e-5


@pytest.mark.parametrize("mas", MAS)
def test_ma_deterministic_mode_is_bit_reproducible(env, oracle, mas):
    """north_star: "a deterministic sorted-segment mode also offered".  Same input twice -> identical bits
    (3D and 2D, weighted, strongly clustered so that many particles share a cell), and the result accumulates
    onto an existing grid like every other mode."""
    torch, MASL, _ = env
    N = 96
    pos, W = make_particles(77, 700001, True, sigma=0.01)
    pos_d, W_d = torch.from_numpy(pos).cuda(), torch.from_numpy(W).cuda()
    runs = []
    for _ in range(3):
        g = torch.full((N, N, N), 0.5, dtype=torch.float32, device="cuda")
        MASL.MA(pos_d, g, BOX, mas, W_d, mode="deterministic")
        runs.append(g.cpu().numpy())
    assert np.array_equal(runs[0], runs[1]) and np.array_equal(runs[0], runs[2])
    # parity on a moderately clustered set (the extreme set above is summation-order noise limited, see
    # test_ma_extreme_clustering_accumulation_noise)
    pos_m, W_m = make_particles(78, 400000, True)
    ref = np.full((N, N, N), 0.5, np.float32)
    oracle.MA(pos_m, ref, BOX, mas, W_m)
    got = np.full((N, N, N), 0.5, np.float32)
    MASL.MA(pos_m, got, BOX, mas, W_m, mode="deterministic")
    assert cell_err(got, ref) < TOL
    p2 = np.ascontiguousarray(pos[:, :2])
    a, b = np.zeros((N, N), np.float32), np.zeros((N, N), np.float32)
    MASL.MA(p2, a, BOX, mas, W, mode="deterministic")
    MASL.MA(p2, b, BOX, mas, W, mode="deterministic")
    assert np.array_equal(a, b)


@pytest.mark.parametrize("pinned", [False, True])
def test_ma_streams_host_particles_in_chunks(env, oracle, monkeypatch, pinned):
    """Host-resident particles cross PCIe chunk by chunk while the previous chunk is deposited; the result is
    the same accumulate as one call (odd chunk count, ragged last chunk, weights, 3D and 2D)."""
    torch, MASL, _ = env
    monkeypatch.setattr(MASL, "STREAM_CHUNK", 70001)
    N = 64
    pos, W = make_particles(11, 400003, True)
    for mas in ("NGP", "PCS"):
        for w in (None, W):
            ref = np.full((N, N, N), 0.125, np.float32)
            oracle.MA(pos, ref, BOX, mas, w)
            got = torch.full((N, N, N), 0.125, dtype=torch.float32, device="cuda")
            p_in = torch.from_numpy(pos).pin_memory() if pinned else pos
            w_in = None if w is None else (torch.from_numpy(w).pin_memory() if pinned else w)
            MASL.MA(p_in, got, BOX, mas, w_in)
            got = got.cpu().numpy()
            if mas == "NGP" and w is None:
                assert np.array_equal(got, ref)
            else:
                assert cell_err(got, ref) < TOL
    p2 = np.ascontiguousarray(pos[:, :2])
    ref2, got2 = np.zeros((N, N), np.float32), np.zeros((N, N), np.float32)
    oracle.MA(p2, ref2, BOX, "TSC", W)
    MASL.MA(p2, got2, BOX, "TSC", W)
    assert cell_err(got2, ref2) < TOL


def test_ma_accumulates_like_reference(env, ma_golden):
    torch, MASL, _ = env
    pos, W = ma_golden["accum_pos"], ma_golden["accum_W"]
    g = np.full((12, 12, 12), 0.25, np.float32)
    MASL.MA(pos[:1500], g, BOX, "TSC", W[:1500])
    MASL.MA(pos[1500:], g, BOX, "TSC", W[1500:])
    assert cell_err(g, ma_golden["accum_TSC_W_3D"]) < TOL
    g2 = np.full((12, 12), 0.25, np.float32)
    MASL.MA(np.ascontiguousarray(pos[:, :2]), g2, BOX, "PCS", None, False, False)
    assert cell_err(g2, ma_golden["accum_PCS_U_2D_norenorm"]) < TOL


def test_ma_device_tensors_zero_copy_and_strided_inputs(env, oracle):
    torch, MASL, _ = env
    N = 48
    pos, W = make_particles(5, 200001, True)           # odd count: exercises the scalar tail
    ref = np.zeros((N, N, N), np.float32)
    oracle.MA(pos, ref, BOX, "CIC", W)
    pos_d = torch.from_numpy(pos).cuda()
    W_d = torch.from_numpy(W).cuda()
    grid = torch.zeros((N, N, N), dtype=torch.float32, device="cuda")
    ptr0 = grid.data_ptr()
    assert MASL.MA(pos_d, grid, BOX, "CIC", W_d) is None
    assert grid.data_ptr() == ptr0
    assert cell_err(grid.cpu().numpy(), ref) < TOL
    # unaligned view of the positions (drops the float4 fast path) and a strided host array
    pos_shift = torch.empty(pos.shape[0] * 3 + 1, dtype=torch.float32, device="cuda")[1:].view(-1, 3)
    pos_shift.copy_(pos_d)
    g2 = torch.zeros_like(grid)
    MASL.MA(pos_shift, g2, BOX, "CIC", W_d)
    assert cell_err(g2.cpu().numpy(), ref) < TOL
    big = np.zeros((pos.shape[0], 6), np.float32)
    big[:, ::2] = pos
    g3 = np.zeros((N, N, N), np.float32)
    MASL.MA(big[:, ::2], g3, BOX, "CIC", W)
    assert cell_err(g3, ref) < TOL


def test_ma_rejects_wrong_dtype(env):
    torch, MASL, _ = env
    with pytest.raises(ValueError):
        MASL.MA(np.zeros((10, 3), np.float64), np.zeros((8, 8, 8), np.float32), BOX)
    with pytest.raises(ValueError):
        MASL.MA(np.zeros((10, 3), np.float32), np.zeros((8, 8, 8), np.float64), BOX)


@pytest.mark.parametrize("mas", MAS)
def test_host_pointer_entry_points_match_mas_c_signature(env, oracle, mas):
    """pyl_NGP/CIC/TSC/PCS take the argument list of MAS_c.h:3-10 on HOST arrays."""
    torch, MASL, _lib = env
    lib = _lib.load()
    N = 40
    pos, W = make_particles(11, 150000, True)
    fp = ctypes.c_void_p
    for axes in (3, 2):
        # ~2-10 particles per cell: float32 sums of O(100) terms differ by >1e-5 between summation
        # orders (serial vs atomics), which is accumulation round-off, not a kernel property
        npart = 150000 if axes == 3 else 4000
        p = np.ascontiguousarray(pos[:npart, :axes])
        for w in (None, W[:npart]):
            # MAS_c.c semantics: accumulate; planes get ONE add per cell (n_max = 1, MAS_c.c:29-34),
            # i.e. the renormalised Cython plane added onto the existing content
            dep = np.zeros((N,) * axes, np.float32)
            oracle.MA(p, dep, BOX, mas, w)
            ref = np.full((N,) * axes, 0.5, np.float32) + dep
            got = np.full((N,) * axes, 0.5, np.float32)
            st = getattr(lib, "pyl_" + mas)(p.ctypes.data, got.ctypes.data, None if w is None else w.ctypes.data,
                                            p.shape[0], N, axes, np.float32(BOX), 8)
            assert st == 0, lib.pyl_last_error()
            assert cell_err(got, ref) < TOL
    assert lib.pyl_host_arena_release() == 0


@pytest.mark.parametrize("mas", MAS)
def test_c_core_wrappers(env, oracle, mas):
    torch, MASL, _ = env
    N = 32
    pos, W = make_particles(3, 50000, False)
    ref = np.zeros((N, N, N), np.float32)
    oracle.MA(pos, ref, BOX, mas, W)
    got = np.zeros((N, N, N), np.float32)
    getattr(MASL, mas + "Wc3D")(pos, got, W, BOX, 4)
    assert cell_err(got, ref) < TOL
    # 2D through the C core: proper plane deposit accumulated onto existing content (MAS_c.c n_max=1)
    p2 = np.ascontiguousarray(pos[:, :2])
    ref2 = np.zeros((N, N), np.float32)
    oracle.MA(p2, ref2, BOX, mas, None)
    got2 = np.full((N, N), 2.0, np.float32)
    getattr(MASL, mas + "c2D")(p2, got2, BOX, 4)
    assert cell_err(got2 - 2.0, ref2) < 1e-4    # the +2 offset costs float32 digits


@pytest.mark.parametrize("mas", MAS)
def test_slab_deposit_with_ghost_planes(env, oracle, mas):
    """Two x-slabs with ghost planes, merged the way the halo exchange does, equal the full grid."""
    torch, MASL, _lib = env
    lib = _lib.load()
    N = 32
    pos, W = make_particles(21, 120000, True)
    ref = np.zeros((N, N, N), np.float32)
    oracle.MA(pos, ref, BOX, mas, W)
    inv = np.float32(N) / np.float32(BOX)
    owner_plane = np.floor((pos[:, 0] * inv).astype(np.float32)).astype(np.int64) % N
    if mas == "NGP":
        owner_plane = (np.floor(pos[:, 0] * inv + 0.5).astype(np.int64)) % N
    lo, hi = 1, 2                                    # ghost planes below / above (enough for PCS)
    full = torch.zeros((N, N, N), dtype=torch.float32, device="cuda")
    bounds = [(0, 13), (13, N)]
    for (a, b) in bounds:
        sel = (owner_plane >= a) & (owner_plane < b)
        p = torch.from_numpy(pos[sel]).cuda()
        w = torch.from_numpy(W[sel]).cuda()
        planes = (b - a) + lo + hi
        slab = torch.zeros((planes, N, N), dtype=torch.float32, device="cuda")
        dropped = torch.zeros(1, dtype=torch.int64, device="cuda")
        st = lib.pyl_deposit_slab(_lib.MAS_IDS[mas], p.data_ptr(), slab.data_ptr(), w.data_ptr(), p.shape[0], N,
                                  np.float32(BOX), (a - lo) % N, planes, planes, dropped.data_ptr(), None, 0,
                                  torch.cuda.current_stream().cuda_stream)
        assert st == 0, lib.pyl_last_error()
        assert int(dropped.item()) == 0
        idx = (torch.arange(planes, device="cuda") + (a - lo)) % N
        full.index_add_(0, idx, slab)
    assert cell_err(full.cpu().numpy(), ref) < TOL


@pytest.mark.parametrize("mas", MAS)
def test_tiled_slab_deposit_routed_by_base_plane(env, oracle, mas):
    """The multi-GPU form of the tiled kernel: particles routed by pyl_stencil_base_plane, every slab has
    S-1 upward ghost planes, the workspace enables the tiled path; slabs + ghosts add up to the grid."""
    torch, MASL, _lib = env
    lib = _lib.load()
    N = 128
    S = MAS.index(mas) + 1
    pos, W = make_particles(31, 1200000, True)
    ref = np.zeros((N, N, N), np.float32)
    oracle.MA(pos, ref, BOX, mas, W)
    pos_d, W_d = torch.from_numpy(pos).cuda(), torch.from_numpy(W).cuda()
    plane = torch.empty(len(pos), dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    assert lib.pyl_stencil_base_plane(_lib.MAS_IDS[mas], pos_d.data_ptr(), len(pos), N, np.float32(BOX),
                                      plane.data_ptr(), stream) == 0
    full = torch.zeros((N, N, N), dtype=torch.float32, device="cuda")
    dropped = torch.zeros(1, dtype=torch.int64, device="cuda")
    for (a, b) in ((0, 41), (41, 90), (90, N)):
        sel = (plane >= a) & (plane < b)
        p, w = pos_d[sel].contiguous(), W_d[sel].contiguous()
        planes = (b - a) + S - 1
        slab = torch.zeros((planes, N, N), dtype=torch.float32, device="cuda")
        need = lib.pyl_deposit_slab_workspace_bytes(_lib.MAS_IDS[mas], p.shape[0], N, b - a)
        assert need > 0
        ws = torch.empty(need, dtype=torch.uint8, device="cuda")
        st = lib.pyl_deposit_slab(_lib.MAS_IDS[mas], p.data_ptr(), slab.data_ptr(), w.data_ptr(), p.shape[0], N,
                                  np.float32(BOX), a, b - a, planes, dropped.data_ptr(), ws.data_ptr(), need, stream)
        assert st == 0, lib.pyl_last_error()
        idx = (torch.arange(planes, device="cuda") + a) % N
        full.index_add_(0, idx, slab)
    assert int(dropped.item()) == 0
    assert cell_err(full.cpu().numpy(), ref) < TOL
    # mis-routed particles are reported, not silently lost
    slab = torch.zeros((20 + S - 1, N, N), dtype=torch.float32, device="cuda")
    need = lib.pyl_deposit_slab_workspace_bytes(_lib.MAS_IDS[mas], len(pos), N, 20)
    ws = torch.empty(max(need, 1), dtype=torch.uint8, device="cuda")
    st = lib.pyl_deposit_slab(_lib.MAS_IDS[mas], pos_d.data_ptr(), slab.data_ptr(), W_d.data_ptr(), len(pos), N,
                              np.float32(BOX), 0, 20, 20 + S - 1, dropped.data_ptr(), ws.data_ptr(), need, stream)
    assert st == 0
    assert int(dropped.item()) > 0


@pytest.mark.parametrize("mas", ["CIC", "PCS"])
def test_mass_conservation_full_size(env, mas):
    """BASELINE config 2 shape (512^3 particles on 512^3 cells), library/tests/test.py:30, plus the
    size-independent properties: linearity in W and agreement between deposit algorithms."""
    torch, MASL, _ = env
    N = 512
    g = torch.Generator(device="cuda"); g.manual_seed(2)
    pos = torch.rand((N ** 3, 3), generator=g, device="cuda", dtype=torch.float32) * BOX
    grid = torch.zeros((N, N, N), dtype=torch.float32, device="cuda")
    MASL.MA(pos, grid, BOX, mas)
    total = float(grid.sum(dtype=torch.float64).item())
    assert abs(total / N ** 3 - 1.0) < 1e-5
    W = torch.full((N ** 3,), 2.0, dtype=torch.float32, device="cuda")
    grid2 = torch.zeros_like(grid)
    MASL.MA(pos, grid2, BOX, mas, W, mode="atomic")
    scale = float(grid.abs().mean().item())
    assert float((grid2 - 2.0 * grid).abs().max().item()) < 2e-5 * max(scale, float(grid.max().item()))


@pytest.mark.parametrize("N", [16, 9])
def test_cic_interp_vs_reference_golden(env, ma_golden, N):
    """Grid -> particle CIC interpolation (MAS_library.pyx:558-599) against the compiled reference."""
    torch, MASL, _ = env
    pos, field, ref = (ma_golden["interp_N%d_%s" % (N, k)] for k in ("pos", "field", "den"))
    den = np.full(len(pos), 7.0, np.float32)
    assert MASL.CIC_interp(field, BOX, pos, den) is None
    assert rel_err(den, ref, floor=float(np.abs(field).mean())) < TOL


def test_cic_interp_vs_oracle_medium_and_device_tensors(env, oracle):
    torch, MASL, _ = env
    N = 128
    pos, _ = make_particles(21, 1000003, True)
    field = np.random.default_rng(5).standard_normal((N, N, N)).astype(np.float32)
    ref = np.zeros(len(pos), np.float32)
    oracle.CIC_interp(field, BOX, pos, ref)
    den = torch.full((len(pos),), -3.0, dtype=torch.float32, device="cuda")
    ptr0 = den.data_ptr()
    MASL.CIC_interp(torch.from_numpy(field).cuda(), BOX, torch.from_numpy(pos).cuda(), den)
    assert den.data_ptr() == ptr0
    assert rel_err(den.cpu().numpy(), ref, floor=float(np.abs(field).mean())) < TOL
    # interpolating a constant field returns the constant (weights sum to 1), the transpose of mass conservation
    one = np.ones((N, N, N), np.float32)
    out = np.zeros(len(pos), np.float32)
    MASL.CIC_interp(one, BOX, pos, out)
    assert np.max(np.abs(out - 1.0)) < 1e-6


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_redshift_space_vs_reference_golden(env, ma_golden, axis):
    """pos_redshift_space (redshift_space_library.pyx:29-46) bit-exact against the compiled reference, on NumPy
    arrays (in place) and on device tensors (zero-copy, in place)."""
    torch, MASL, _ = env
    from pylians3_b200 import redshift_space_library as RSL
    ref = ma_golden["rsd_out_a%d" % axis]
    pos = ma_golden["rsd_pos"].copy()
    assert RSL.pos_redshift_space(pos, ma_golden["rsd_vel"], BOX, 171.5, 0.5, axis) is None
    assert np.array_equal(pos, ref)
    pos_d = torch.from_numpy(ma_golden["rsd_pos"].copy()).cuda()
    ptr0 = pos_d.data_ptr()
    RSL.pos_redshift_space(pos_d, torch.from_numpy(ma_golden["rsd_vel"]).cuda(), BOX, 171.5, 0.5, axis)
    assert pos_d.data_ptr() == ptr0 and np.array_equal(pos_d.cpu().numpy(), ref)


def test_device_pipeline_redshift_space_to_pk(env, oracle):
    """positions -> redshift space -> MA -> delta -> Pk entirely on the device (the Pk_snapshot.Pk_comp recipe,
    Pk_snapshot.py:57-91) against the same chain on the CPU oracle."""
    torch, MASL, _ = env
    from pylians3_b200 import Pk_library as PKL, overdensity_, redshift_space_library as RSL
    from test_gpu_pk import check_pk
    N = 64
    pos, _ = make_particles(31, 500000, True)
    vel = (np.random.default_rng(32).standard_normal(pos.shape) * 400.0).astype(np.float32)
    pos_d, vel_d = torch.from_numpy(pos.copy()).cuda(), torch.from_numpy(vel).cuda()
    RSL.pos_redshift_space(pos_d, vel_d, BOX, 100.0, 0.0, 2)
    grid = torch.zeros((N, N, N), dtype=torch.float32, device="cuda")
    MASL.MA(pos_d, grid, BOX, "CIC")
    overdensity_(grid)
    got = PKL.Pk(grid, BOX, 2, "CIC", verbose=False)
    oracle.pos_redshift_space(pos, vel, BOX, 100.0, 0.0, 2)
    ref = np.zeros((N, N, N), np.float32)
    oracle.MA(pos, ref, BOX, "CIC")
    ref /= np.mean(ref, dtype=np.float64)
    ref -= 1.0
    # spectra compared on the delta the device holds (deposit order noise is checked elsewhere)
    assert np.max(np.abs(grid.cpu().numpy() - ref)) < 1e-5 * np.max(np.abs(ref))
    check_pk(got, oracle.Pk(grid.cpu().numpy(), BOX, 2, "CIC", 1, False))


def test_per_scheme_entry_points(env, oracle):
    """MASL.CIC(pos, number, BoxSize) / MASL.PCSW(pos, number, BoxSize, W) ...: the reference's cpdef functions
    (MAS_library.pyx:123-545), which existing scripts call directly."""
    torch, MASL, _ = env
    N = 32
    pos, W = make_particles(5, 20000)
    for mas in MAS:
        for w in (None, W):
            ref = np.zeros((N, N, N), np.float32)
            oracle.MA(pos, ref, BOX, mas, w)
            got = np.zeros((N, N, N), np.float32)
            if w is None:
                getattr(MASL, mas)(pos, got, BOX)
            else:
                getattr(MASL, mas + "W")(pos, got, BOX, w)
            assert cell_err(got, ref) < TOL, (mas, w is not None)
    # the (dims, dims, 1) plane view of the reference's 2D calls: no renormalisation outside MA
    p2 = np.ascontiguousarray(pos[:, :2])
    got = np.zeros((N, N, 1), np.float32)
    MASL.CIC(p2, got, BOX)
    ref = np.zeros((N, N), np.float32)
    oracle.MA(p2, ref, BOX, "CIC", None, renormalize_2D=False)
    assert cell_err(got[:, :, 0], ref) < TOL
    assert abs(float(got.sum()) / (2.0 * len(p2)) - 1.0) < 1e-5          # every particle lands twice (:138-139)


@pytest.mark.parametrize("N,n_side", [(128, 128), (64, 160), (96, 128)])
def test_tiled_deposit_of_lattice_ordered_particles(env, oracle, N, n_side):
    """Spatially ordered input (a displaced lattice in lattice order, like initial conditions and most snapshots)
    takes the partition's direct route: the first pass writes straight into the tile buckets.  Random-order tests
    never reach that code; this one does (and also covers a grid that is not a multiple of the tile size)."""
    torch, MASL, _ = env
    from pylians3_b200 import synth
    pos = synth.zeldovich_host(n_side, BOX, 11)
    W = np.random.default_rng(4).random(len(pos), dtype=np.float32)
    for mas in MAS:
        for w in (None, W):
            ref = np.zeros((N, N, N), np.float32)
            oracle.MA(pos, ref, BOX, mas, w)
            got = np.zeros((N, N, N), np.float32)
            MASL.MA(pos, got, BOX, mas, w, mode="tiled")
            if mas == "NGP" and w is None:
                assert np.array_equal(got, ref)
            else:
                assert cell_err(got, ref) < TOL, (mas, w is not None)
            tot = float(np.sum(W, dtype=np.float64)) if w is not None else float(len(pos))
            assert abs(np.sum(got, dtype=np.float64) / tot - 1.0) < 1e-5
