"""Host-side logic of the slab decomposition (no GPU): the partition and the neighbour exchange
protocol over torch.distributed with the gloo backend, world_size 2 and 3."""
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp


def test_slab_cuts_balance_and_cover(fsg):
    rng = np.random.default_rng(0)
    for G, world in [(17, 2), (64, 8), (40, 3), (16, 8)]:
        hist = np.zeros(G, np.int64)
        lo, hi = G // 4, max(G // 4 + 1, 3 * G // 4)
        hist[lo:hi] = rng.integers(50, 100, hi - lo)
        if 2 * world == G:
            hist[:] = 1
        cuts = fsg.slab_cuts(hist, world)
        assert cuts[0][0] == 0 and cuts[-1][1] == G and len(cuts) == world
        assert all(b - a >= 2 for a, b in cuts) and all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
        per = [int(hist[a:b].sum()) for a, b in cuts]
        if world < hi - lo:
            assert max(per) - min(per) <= 2 * hist.max()      # balanced to within a layer or two
    with pytest.raises(ValueError):
        fsg.slab_cuts(np.ones(9), 5)


def test_plume_layer_hist_matches_scene(fsg):
    cfg = fsg.scenes.plume_config(24)
    hist = fsg.slab.plume_layer_hist(cfg)
    state = fsg.scenes.plume_scene(cfg, jitter=0.0)
    assert hist.sum() == state["pos"].shape[0]
    assert np.array_equal(hist, fsg.slab.layer_hist_from_positions(cfg, state["pos"]))
    assert fsg.slab.message_bytes(3, 5) == fsg._lib.load().fsg_slab_message_bytes(3, 5) == 64 + (4 * 3 + 2 * 5) * 16 + 64
    # unidyn messages carry the volume fractions too
    assert fsg.slab.message_bytes(3, 5, 1) == fsg._lib.load().fsg_slab_message_bytes_model(1, 3, 5) == 64 + (5 * 3 + 3 * 5) * 16 + 64
    assert fsg._lib.load().fsg_slab_message_bytes_model(0, 3, 5) == fsg.slab.message_bytes(3, 5)
    cap_m, cap_g = fsg.slab.message_caps(hist, fsg.slab_cuts(hist, 3))
    assert cap_g >= hist.max() and cap_m >= 4096


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _exchange_worker(rank, world, port, rounds):
    import torch.distributed as dist
    from fluidsolvergpu_b200.slab import DistExchange, message_bytes
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        ex = DistExchange()
        nbytes = message_bytes(37, 211)

        def payload(r, side, rnd):
            g = np.random.default_rng(7 + 10 * r + side + 100 * rnd)
            return torch.from_numpy(g.integers(0, 256, nbytes, dtype=np.uint8))
        for rnd in range(rounds):
            to_left, to_right = payload(rank, 0, rnd), payload(rank, 1, rnd)
            from_left, from_right = torch.full((nbytes,), 255, dtype=torch.uint8), torch.full((nbytes,), 255, dtype=torch.uint8)
            ex.exchange(to_left, to_right, from_left, from_right, nbytes)
            if rank > 0:
                assert torch.equal(from_left, payload(rank - 1, 1, rnd))        # the left neighbour's message to its right
            else:
                assert bool((from_left == 255).all())                           # untouched: no left neighbour
            if rank < world - 1:
                assert torch.equal(from_right, payload(rank + 1, 0, rnd))
            else:
                assert bool((from_right == 255).all())
            tot = ex.global_sum([rank + 1, 1], "cpu")
            assert tot[1] == world and tot[0] == world * (world + 1) // 2
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_dist_exchange_gloo(world):
    mp.spawn(_exchange_worker, args=(world, _free_port(), 3), nprocs=world, join=True)


def test_slab_cuts_with_a_ghost_layer_weight():
    """ghost_weight adds a share of the layer below each slab to its load (the symmetric pair kernel walks that ghost layer as home
    bins): the cuts stay a partition into slabs of at least two layers, and the weighted maximum is never worse than what the
    unweighted cuts give under the same measure."""
    import numpy as np
    import fluidsolvergpu_b200 as fsg
    rng = np.random.default_rng(5)
    for G, world in ((40, 3), (64, 8), (17, 2)):
        hist = (rng.integers(0, 1000, size=G) * (rng.random(G) < 0.7)).astype(np.int64)
        hist[G // 3:2 * G // 3] += 500

        def load(cuts, w):
            return max(int(hist[a:b].sum()) + (int(w * hist[a - 1]) if a > 0 else 0) for a, b in cuts)
        plain, weighted = fsg.slab_cuts(hist, world), fsg.slab_cuts(hist, world, ghost_weight=0.6)
        for cuts in (plain, weighted):
            assert cuts[0][0] == 0 and cuts[-1][1] == G and all(b - a >= 2 for a, b in cuts)
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
        assert load(weighted, 0.6) <= load(plain, 0.6)
