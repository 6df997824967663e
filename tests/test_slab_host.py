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
    assert fsg.slab.message_bytes(3, 5) == fsg._lib.load().fsg_slab_message_bytes(3, 5) == (4 * 3 + 2 * 5) * 16


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
        cap = 1 << 16
        for rnd in range(rounds):
            # counts every rank can recompute for every other rank
            def counts_of(r):
                g = np.random.default_rng(1000 * rnd + r)
                c = [int(v) for v in g.integers(0, 40, 4)]
                if rnd == 1:
                    c = [0, 0, 0, 0]                    # a round with nothing to move
                if r == 0:
                    c[0] = c[1] = 0                     # no left neighbour
                if r == world - 1:
                    c[2] = c[3] = 0
                return c

            def payload(r, side, nbytes):
                g = np.random.default_rng(7 + 10 * r + side + 100 * rnd)
                return torch.from_numpy(g.integers(0, 256, nbytes, dtype=np.uint8))
            c = counts_of(rank)
            to_left, to_right = torch.zeros(cap, dtype=torch.uint8), torch.zeros(cap, dtype=torch.uint8)
            nl, nr = message_bytes(c[0], c[1]), message_bytes(c[2], c[3])
            to_left[:nl] = payload(rank, 0, nl)
            to_right[:nr] = payload(rank, 1, nr)
            from_left, from_right = torch.full((cap,), 255, dtype=torch.uint8), torch.full((cap,), 255, dtype=torch.uint8)
            fl, fr = ex.exchange(c, to_left, to_right, from_left, from_right)
            if rank > 0:
                cl = counts_of(rank - 1)
                assert fl == (cl[2], cl[3])
                nb = message_bytes(*fl)
                assert torch.equal(from_left[:nb], payload(rank - 1, 1, nb))
            else:
                assert fl == (0, 0)
            if rank < world - 1:
                cr = counts_of(rank + 1)
                assert fr == (cr[0], cr[1])
                nb = message_bytes(*fr)
                assert torch.equal(from_right[:nb], payload(rank + 1, 0, nb))
            else:
                assert fr == (0, 0)
            tot = ex.global_sum([c[0] + c[2], 1], "cpu")
            assert tot[1] == world and tot[0] == sum(counts_of(r)[0] + counts_of(r)[2] for r in range(world))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_dist_exchange_gloo(world):
    mp.spawn(_exchange_worker, args=(world, _free_port(), 3), nprocs=world, join=True)
