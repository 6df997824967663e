"""Slab decomposition of the particle step across devices — host side.

The reference's multi-device design (solver-unidyn.cu:187-195, 396-470) cuts the bin grid in two along
the linear bin id (= along x, the slowest axis), keeps a one-layer `buffer` of foreign particles and
moves whole 340-byte Particle ranges through host memory every step.  Here the same decomposition is
generalised to N slabs, one process (and one libfsg context) per device:

    every step:  fsg_slab_pack  ->  exchange with the two x-neighbours  ->  fsg_slab_unpack  ->  fsg_step(1)

Only the exchange lives in this file; it moves two device buffers per neighbour with
torch.distributed P2P (NCCL over NVLink on GPUs, gloo in the CPU tests) after a count exchange.
There is no data-path collective other than that neighbour exchange; `global_sum` is the one small
all-reduce used for diagnostics (particle count conservation).

`SlabGroup` drives W contexts in ONE process on one device with an in-process exchange — the way the
multi-rank algorithm is tested on a single GPU against the single-slab result.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import FsgConfig
from .solver import FluidSolver


# ---------------------------------------------------------------------------------------------
# partition
# ---------------------------------------------------------------------------------------------
def slab_cuts(hist, world: int, min_layers: int = 2) -> list[tuple[int, int]]:
    """Cuts bin layers 0..G-1 into `world` contiguous slabs of (nearly) equal particle count.
    hist[ix] = particles in bin layer ix.  Empty outer layers go to the end ranks.
    Every slab is at least `min_layers` thick: a particle that migrates into a slab must not at the
    same time be needed as a ghost by the slab beyond it (the one-layer ghost band of the reference,
    solver-unidyn.cu:187, assumes the same)."""
    hist = np.asarray(hist, dtype=np.int64)
    G = hist.shape[0]
    if world < 1 or world > G:
        raise ValueError(f"cannot cut {G} bin layers into {world} slabs")
    if world == 1:
        return [(0, G)]
    if G < min_layers * world:
        raise ValueError(f"{G} bin layers cannot hold {world} slabs of at least {min_layers} layers")
    total = int(hist.sum())
    cum = np.concatenate([[0], np.cumsum(hist)])
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        x = int(np.searchsorted(cum, target, side="left"))
        # nearest layer boundary to the target, leaving room for the remaining slabs
        if x > 0 and abs(cum[x - 1] - target) <= abs(cum[min(x, G)] - target):
            x -= 1
        x = max(x, cuts[-1] + min_layers)
        x = min(x, G - min_layers * (world - r))
        cuts.append(x)
    cuts.append(G)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def plume_layer_hist(cfg: FsgConfig, spacing: float = 0.05) -> np.ndarray:
    hist = np.zeros(cfg.grid, np.int64)
    rc = _lib.load().fsg_scene_plume_hist(C.byref(cfg), spacing, hist.ctypes.data)
    if rc != 0:
        raise _lib.FsgError(rc, "fsg_scene_plume_hist")
    return hist


def layer_hist_from_positions(cfg: FsgConfig, pos: np.ndarray) -> np.ndarray:
    """Particles per bin layer for an arbitrary host scene (the x part of the bin id expression,
    FluidGPU.cu:419: float subtraction, double division, truncation)."""
    fx = (np.asarray(pos, np.float32)[:, 0] - np.float32(cfg.origin)).astype(np.float32).astype(np.float64)
    ix = np.clip(np.trunc(fx / cfg.cellsize).astype(np.int64), 0, cfg.grid - 1)
    return np.bincount(ix, minlength=cfg.grid).astype(np.int64)


def slab_config(base: FsgConfig, rank: int, world: int, cuts, capacity: int, device: int | None = None) -> FsgConfig:
    cfg = FsgConfig()
    C.memmove(C.byref(cfg), C.byref(base), C.sizeof(FsgConfig))
    cfg.rank, cfg.world = rank, world
    cfg.slab_x0, cfg.slab_x1 = cuts[rank]
    cfg.capacity = capacity
    if device is not None:
        cfg.device = device
    return cfg


def message_bytes(m: int, g: int) -> int:
    return (4 * m + 2 * g) * 16


# ---------------------------------------------------------------------------------------------
# exchange back-ends
# ---------------------------------------------------------------------------------------------
class DistExchange:
    """Neighbour exchange over torch.distributed (one rank per process).  Buffers are torch uint8
    tensors on the communication device (cuda for NCCL, cpu for gloo)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def exchange_counts(self, counts4, device):
        """counts4 = (mig_left, ghost_left, mig_right, ghost_right) of this rank -> what the left
        neighbour sends right to us and what the right neighbour sends left to us."""
        import torch
        mine = torch.tensor(list(counts4), dtype=torch.int64, device=device)
        every = torch.empty(self.world * 4, dtype=torch.int64, device=device)
        self.dist.all_gather_into_tensor(every, mine, group=self.group)
        every = every.cpu().view(self.world, 4)
        from_left = (int(every[self.rank - 1, 2]), int(every[self.rank - 1, 3])) if self.rank > 0 else (0, 0)
        from_right = (int(every[self.rank + 1, 0]), int(every[self.rank + 1, 1])) if self.rank < self.world - 1 else (0, 0)
        return from_left, from_right

    def exchange(self, counts4, to_left, to_right, from_left, from_right):
        """Moves to_left -> rank-1's from_right and to_right -> rank+1's from_left.  Returns
        ((mig, ghost) from the left neighbour, (mig, ghost) from the right neighbour)."""
        dist = self.dist
        fl, fr = self.exchange_counts(counts4, to_left.device)
        ops = []
        if self.rank > 0:
            nb = message_bytes(counts4[0], counts4[1])
            if nb:
                ops.append(dist.P2POp(dist.isend, to_left[:nb], self.rank - 1, self.group))
            nb = message_bytes(*fl)
            if nb:
                if nb > from_left.numel():
                    raise RuntimeError("slab exchange: receive buffer too small")
                ops.append(dist.P2POp(dist.irecv, from_left[:nb], self.rank - 1, self.group))
        if self.rank < self.world - 1:
            nb = message_bytes(counts4[2], counts4[3])
            if nb:
                ops.append(dist.P2POp(dist.isend, to_right[:nb], self.rank + 1, self.group))
            nb = message_bytes(*fr)
            if nb:
                if nb > from_right.numel():
                    raise RuntimeError("slab exchange: receive buffer too small")
                ops.append(dist.P2POp(dist.irecv, from_right[:nb], self.rank + 1, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return fl, fr

    def global_sum(self, values, device):
        import torch
        t = torch.tensor(list(values), dtype=torch.int64, device=device)
        self.dist.all_reduce(t, group=self.group)
        return [int(v) for v in t.cpu()]


# ---------------------------------------------------------------------------------------------
# one slab
# ---------------------------------------------------------------------------------------------
class SlabSolver(FluidSolver):
    """One slab: a libfsg context with rank/world/slab range set + the message buffers.  With
    `exchange` = DistExchange this is the per-process solver of a multi-GPU run."""

    def __init__(self, cfg: FsgConfig, exchange=None, msg_bytes: int | None = None):
        import torch
        super().__init__(cfg)
        self.torch = torch
        self.exchange = exchange
        self.tdev = torch.device("cuda", cfg.device)
        if msg_bytes is None:
            msg_bytes = max(1 << 20, int(cfg.capacity) * 64 // 4)
        self.msg_bytes = int(msg_bytes)
        with torch.cuda.device(self.tdev):
            self.to_left, self.to_right, self.from_left, self.from_right = (
                torch.empty(self.msg_bytes, dtype=torch.uint8, device=self.tdev) for _ in range(4))
        self.tstream = torch.cuda.ExternalStream(self.stream(), device=self.tdev)
        self.last_counts = (0, 0, 0, 0)
        self.traffic_bytes = 0

    # -- the three slab phases --
    def pack(self):
        counts = (C.c_int64 * 5)()
        self._check(self._lib.fsg_slab_pack(self._ctx, self.to_left.data_ptr(), self.to_right.data_ptr(), self.msg_bytes,
                                            C.byref(counts)), "fsg_slab_pack")
        self.last_counts = tuple(int(v) for v in counts[:4])
        return self.last_counts

    def unpack(self, from_left_ptr, fl, from_right_ptr, fr):
        self._check(self._lib.fsg_slab_unpack(self._ctx, from_left_ptr, fl[0], fl[1], from_right_ptr, fr[0], fr[1]), "fsg_slab_unpack")

    def step(self, nsteps: int = 1, sync: bool = True):
        torch = self.torch
        for _ in range(nsteps):
            counts = self.pack()
            with torch.cuda.stream(self.tstream):       # P2P ops are ordered after the pack kernels on the solver's stream
                fl, fr = self.exchange.exchange(counts, self.to_left, self.to_right, self.from_left, self.from_right)
            self.traffic_bytes += message_bytes(counts[0], counts[1]) + message_bytes(counts[2], counts[3])
            self.unpack(self.from_left.data_ptr(), fl, self.from_right.data_ptr(), fr)
            self._check(self._lib.fsg_step(self._ctx, 1), "fsg_step")
        if sync:
            self.sync()

    def download(self, fields=None) -> dict:
        """Only the particles this slab owns (ghost / migrated slots are dropped)."""
        out = super().download()
        keep = out["cell"] <= self.numcells
        return {k: v[keep] for k, v in out.items() if fields is None or k in fields}

    def owned_count(self) -> int:
        return int((super().download(("cell",))["cell"] <= self.numcells).sum())


# ---------------------------------------------------------------------------------------------
# W slabs in one process on one device (tests; single-GPU emulation of the multi-rank algorithm)
# ---------------------------------------------------------------------------------------------
class SlabGroup:
    def __init__(self, base_cfg: FsgConfig, world: int, cuts, capacity: int, device: int = 0, msg_bytes: int | None = None):
        self.world = world
        self.cuts = cuts
        self.slabs = [SlabSolver(slab_config(base_cfg, r, world, cuts, capacity, device), None, msg_bytes) for r in range(world)]

    def close(self):
        for s in self.slabs:
            s.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def upload(self, state: dict):
        for s in self.slabs:
            s.upload(state)          # every slab sees the scene; the key kernel keeps what it owns

    def scene_plume(self, *a, **kw):
        return sum(s.scene_plume(*a, **kw) for s in self.slabs)

    def step(self, nsteps: int = 1):
        for _ in range(nsteps):
            counts = [s.pack() for s in self.slabs]
            for s in self.slabs:
                s.sync()             # messages are read by the neighbour's stream
            for r, s in enumerate(self.slabs):
                left, right = (self.slabs[r - 1] if r > 0 else None), (self.slabs[r + 1] if r < self.world - 1 else None)
                fl = (counts[r - 1][2], counts[r - 1][3]) if left else (0, 0)
                fr = (counts[r + 1][0], counts[r + 1][1]) if right else (0, 0)
                s.unpack(left.to_right.data_ptr() if left else None, fl, right.to_left.data_ptr() if right else None, fr)
                s.traffic_bytes += message_bytes(counts[r][0], counts[r][1]) + message_bytes(counts[r][2], counts[r][3])
            for s in self.slabs:
                s.sync()
            for s in self.slabs:
                s._check(s._lib.fsg_step(s._ctx, 1), "fsg_step")
            for s in self.slabs:
                s.sync()

    def download(self) -> dict:
        parts = [s.download() for s in self.slabs]
        return {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
