"""CPU check of the decomposition the symmetric pair kernel (fluidsolvergpu_b200/csrc/fsg_pair_v3.cu) rests on.

The reference enumerates the neighbour bins of bin b by 27 LINEAR offsets dx*G^2 + dy*G + dz, dx, dy, dz in {-1, 0, 1}
(FluidGPU.cu:124-126), clipped only to [0, G^3) — so the neighbourhood wraps across rows and planes at the grid faces
(SURVEY.md B.3).  The kernel evaluates a pair of particles in two different bins once, in the bin with the lower id.  That is
exact iff (1) the neighbour relation is symmetric, wrap-around included, and (2) the neighbours with a larger id are exactly the
five contiguous runs of bin ids the kernel stages: [b, b+1], [b+G-1, b+G+1] and [b+G^2+dy*G-1, b+G^2+dy*G+1] for dy = -1, 0, 1."""
import numpy as np
import pytest


def neighbours(b, G):
    nc = G ** 3
    offs = [dx * G * G + dy * G + dz for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1)]
    return {b + o for o in offs if 0 <= b + o < nc}


def forward_runs(b, G):
    """The runs as fsg_pair_v3.cu's prefetch() forms them: run r has a centre bin c0 and takes c0-1 (not for run 0), c0, c0+1."""
    nc, G2 = G ** 3, G * G
    out = []
    for r in range(5):
        c0 = b + (0 if r == 0 else G if r == 1 else G2 + (r - 3) * G)
        bins = ([] if r == 0 else [c0 - 1]) + [c0, c0 + 1]
        out.append([c for c in bins if 0 <= c < nc])
    return out


@pytest.mark.parametrize("G", [4, 5, 8, 17])
def test_linear_offset_neighbourhood_is_symmetric_and_its_upper_half_is_five_runs(G):
    nc = G ** 3
    nb = [neighbours(b, G) for b in range(nc)]
    for b in range(nc):
        assert all(b in nb[c] for c in nb[b]), b                          # (1) symmetric, wrap-around included
        runs = forward_runs(b, G)
        flat = [c for run in runs for c in run]
        assert len(flat) == len(set(flat)), (b, runs)                       # the runs do not overlap (needs G >= 4) ...
        assert flat == sorted(flat), (b, runs)                              # ... and are staged in ascending bin order
        assert set(flat) - {b} == {c for c in nb[b] if c > b}, (b, runs)    # (2) exactly the neighbours with a larger id
        for run in runs:                                                    # each run is one contiguous range of bin ids
            assert run == list(range(run[0], run[0] + len(run))) if run else True


def test_every_unordered_pair_of_neighbouring_bins_is_owned_by_exactly_one_bin():
    G = 6
    nc = G ** 3
    owned = {}
    for b in range(nc):
        for c in {c for run in forward_runs(b, G) for c in run} - {b}:
            key = (min(b, c), max(b, c))
            owned[key] = owned.get(key, 0) + 1
    pairs = {(min(b, c), max(b, c)) for b in range(nc) for c in neighbours(b, G) if c != b}
    assert set(owned) == pairs and set(owned.values()) == {1}
