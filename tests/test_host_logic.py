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
