"""Host-side logic of the slab pipeline that needs neither a GPU nor a process group."""


def test_interleaved_row_order_is_a_rotated_round_robin_permutation():
    """Issue order of the peer-memory transpose: a permutation of the rows, owners interleaved row by row, every
    rank starting on a different owner (no two senders on one receiver in lock step)."""
    from pylians3_b200 import dist as PD
    for N, P in ((64, 8), (45, 2), (72, 8), (33, 3)):
        sizes, offs = PD.split_sizes(N // 2 + 1, P)
        ky_rows = [PD.mirrored_rows(N, offs[r], sizes[r]) for r in range(P)]
        owner = {ky: r for r, rows in enumerate(ky_rows) for ky in rows}
        firsts = set()
        for rank in range(P):
            order = PD.interleaved_row_order(ky_rows, rank)
            assert sorted(order) == list(range(N))
            head = [owner[ky] for ky in order[:P]]
            assert head == [(rank + d) % P for d in range(1, P + 1)]
            firsts.add(head[0])
        assert len(firsts) == P


def test_offsets_of_the_density_spectra():
    """`offset=` of Pk/XPk(density=True): a number, a sequence of numbers/tensors, or one tensor -> float64 [fields]."""
    import numpy as np
    import pytest
    import torch
    from pylians3_b200 import Pk_library as PKL
    cpu = torch.device("cpu")
    assert PKL._offsets(None, 3, cpu) is None
    assert PKL._offsets(0.5, 3, cpu).tolist() == [0.5, 0.5, 0.5]
    assert PKL._offsets(np.float32(0.25), 2, cpu).dtype == torch.float64
    t = PKL._offsets([0.5, torch.tensor([1.5], dtype=torch.float32), 2], 3, cpu)
    assert t.dtype == torch.float64 and t.tolist() == [0.5, 1.5, 2.0] and t.is_contiguous()
    assert PKL._offsets(torch.tensor([3.0]), 1, cpu).tolist() == [3.0]
    assert PKL._offsets(torch.tensor([3.0]), 2, cpu).tolist() == [3.0, 3.0]
    with pytest.raises(ValueError):
        PKL._offsets([1.0, 2.0], 3, cpu)


def test_kpar_kper_on_first_access():
    """K2D / _Spectra: kpar and kper are index arithmetic (Pk_library.pyx:394-399), computed on the host when first
    asked for and kept; equal to the arrays of the host finalisation."""
    import numpy as np
    from pylians3_b200 import Pk_library as PKL
    kF, kN, kmax_par, kmax_per, kmax = PKL.frequencies(1000.0, 24)
    want = PKL._kpar_kper(kmax_par, kmax_per, kF)

    class R(PKL.K2D):
        pass
    r = R()
    r._kgrid = (kmax_par, kmax_per, kF)
    assert "_k2d_v" not in r.__dict__
    assert np.array_equal(r.kpar, want[0]) and np.array_equal(r.kper, want[1])
    assert r.kpar is r.kpar                                  # computed once
    assert r.kpar.shape == ((kmax_par + 1) * (kmax_per + 1),)
    # bin i2 = k_per_index * (kmax_par + 1) + k_par_index; centres at (index + 0.5) kF
    assert r.kpar[1] == 1.5 * kF and r.kper[kmax_par + 1] == 1.5 * kF and r.kpar[kmax_par + 1] == 0.5 * kF
    o = PKL._Spectra(kgrid=(kmax_par, kmax_per, kF))
    assert "kpar" not in o
    assert np.array_equal(o["kper"], want[1]) and "kpar" in o
    try:
        o["nothing"]
        raise AssertionError("missing keys must still raise")
    except KeyError:
        pass


def test_weight_total_on_the_host():
    """prebias_'s numerator: the particle count, or the exact float64 sum of host weights up to HOST_SUM_LIMIT."""
    import numpy as np
    import torch
    from pylians3_b200 import field
    cpu = torch.device("cpu")
    assert field.weight_total(12345, None, cpu).tolist() == [12345.0]
    W = np.random.default_rng(1).random(100000, dtype=np.float32)
    got = field.weight_total(len(W), W, cpu)
    assert got.dtype == torch.float64 and abs(float(got) - float(W.sum(dtype=np.float64))) < 1e-9 * float(W.sum())
    got_t = field.weight_total(len(W), torch.from_numpy(W), cpu)
    assert float(got_t) == float(got)
    # beyond the limit: an estimate from a strided sample (documented as such)
    saved = field.HOST_SUM_LIMIT
    try:
        field.HOST_SUM_LIMIT = 1000
        est = field.weight_total(len(W), W, cpu)
        assert abs(float(est) / float(got) - 1.0) < 0.02
    finally:
        field.HOST_SUM_LIMIT = saved
