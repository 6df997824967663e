"""Slab decomposition of the particle step across devices — host side.

The reference's multi-device design (solver-unidyn.cu:187-195, 396-470) cuts the bin grid in two along
the linear bin id (= along x, the slowest axis), keeps a one-layer `buffer` of foreign particles and
moves whole 340-byte Particle ranges through host memory every step.  Here the same decomposition is
generalised to N slabs, one process (and one libfsg context) per device:

    every step:  fsg_slab_pack  ->  exchange with the two x-neighbours  ->  fsg_slab_unpack  ->  fsg_step(1)

Only the exchange lives in this file: one fixed-size message per neighbour and direction, moved with
torch.distributed P2P (NCCL over NVLink on GPUs, gloo in the CPU tests).  The particle counts travel in
the message header and are read on the device, so a step never waits for the host: every call only
enqueues work on the solver's stream and the host runs steps ahead of the device.  There is no
data-path collective other than that neighbour exchange; `global_sum` is the one small all-reduce used
for diagnostics (particle count conservation).

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
def slab_cuts(hist, world: int, min_layers: int = 2, ghost_weight: float = 0.0) -> list[tuple[int, int]]:
    """Cuts bin layers 0..G-1 into `world` contiguous slabs minimising the largest load, load = particles owned +
    ghost_weight x particles of the layer below the slab (the symmetric pair kernel walks that ghost layer as home bins
    restricted to 3 of its 5 runs: ghost_weight = 0.6; 0 = plain particle count).
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
    cum = np.concatenate([[0], np.cumsum(hist)])
    gw = float(ghost_weight)

    def ghost(x):
        return int(gw * hist[x - 1]) if (gw > 0.0 and x > 0) else 0

    def feasible(limit):
        """Greedy: can the layers be covered by `world` slabs of >= min_layers layers and <= limit particles?"""
        cuts, x = [0], 0
        for r in range(world):
            remaining = world - r - 1
            hi = G - min_layers * remaining                       # leave room for the slabs to come
            lo = x + min_layers
            if lo > hi:
                return None
            # furthest end with load <= limit
            e = int(np.searchsorted(cum, cum[x] + limit - ghost(x), side="right")) - 1
            e = min(e, hi)
            if e < lo:
                return None
            if remaining == 0:
                if cum[G] - cum[x] + ghost(x) > limit:
                    return None
                e = G
            cuts.append(e)
            x = e
        return cuts

    # smallest achievable maximum load (binary search over the load limit)
    lo_l, hi_l = int(hist.max()) * min_layers // 2, int(hist.sum() * (1.0 + gw)) + 1
    best = feasible(hi_l)
    while lo_l < hi_l:
        mid = (lo_l + hi_l) // 2
        c = feasible(mid)
        if c is not None:
            best, hi_l = c, mid
        else:
            lo_l = mid + 1
    cuts = best
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


def message_bytes(cap_m: int, cap_g: int, model: int = 0) -> int:
    """Bytes of one slab message with room for cap_m migrants and cap_g ghosts (fsg_slab_message_bytes_model): the unidyn
    model (1) also carries the volume fractions."""
    mix = model == 1
    return 64 + ((5 if mix else 4) * cap_m + (3 if mix else 2) * cap_g) * 16 + 64


def message_caps(hist, cuts, slack: float = 1.2, floor: int = 4096) -> tuple[int, int]:
    """Message capacities every rank agrees on: ghosts = the most populated face layer x slack, migrants =
    a tenth of that (a particle moves much less than one bin per step)."""
    hist = np.asarray(hist, dtype=np.int64)
    faces = [int(hist[a]) for a, b in cuts] + [int(hist[b - 1]) for a, b in cuts]
    cap_g = int(max(faces) * slack) + floor
    return max(floor, cap_g // 32), cap_g


# ---------------------------------------------------------------------------------------------
# exchange back-end
# ---------------------------------------------------------------------------------------------
class DistExchange:
    """Neighbour exchange over torch.distributed (one rank per process).  Buffers are torch uint8 tensors on
    the communication device (cuda for NCCL, cpu for gloo); every message has the same, fixed size."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def exchange(self, to_left, to_right, from_left, from_right, nbytes: int):
        """to_left -> rank-1's from_right, to_right -> rank+1's from_left; asynchronous on the current stream."""
        dist = self.dist
        ops = []
        if self.rank > 0:
            ops.append(dist.P2POp(dist.isend, to_left[:nbytes], self.rank - 1, self.group))
            ops.append(dist.P2POp(dist.irecv, from_left[:nbytes], self.rank - 1, self.group))
        if self.rank < self.world - 1:
            ops.append(dist.P2POp(dist.isend, to_right[:nbytes], self.rank + 1, self.group))
            ops.append(dist.P2POp(dist.irecv, from_right[:nbytes], self.rank + 1, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()            # CUDA: makes the current stream wait, not the host

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

    def __init__(self, cfg: FsgConfig, exchange=None, cap_m: int = 4096, cap_g: int = 65536):
        import torch
        super().__init__(cfg)
        self.torch = torch
        self.exchange = exchange
        self.tdev = torch.device("cuda", cfg.device)
        self.cap_m, self.cap_g = int(cap_m), int(cap_g)
        self.msg_bytes = message_bytes(self.cap_m, self.cap_g, cfg.model)
        with torch.cuda.device(self.tdev):
            self.to_left, self.to_right, self.from_left, self.from_right = (
                torch.zeros(self.msg_bytes, dtype=torch.uint8, device=self.tdev) for _ in range(4))
        self.tstream = torch.cuda.ExternalStream(self.stream(), device=self.tdev)
        self.steps_exchanged = 0
        self.peer = False               # True after setup_peer_exchange(): messages go straight into the neighbours' memory
        self.time_exchange = False      # record CUDA events around pack / exchange / unpack (exchange_ms())
        self._ev = []

    def close(self):
        if self._ctx and self.peer:
            self._lib.fsg_slab_close_peers(self._ctx)
            self.peer = False
        super().close()

    # -- the slab phases (all asynchronous) --
    def pack(self):
        self._check(self._lib.fsg_slab_pack(self._ctx, self.to_left.data_ptr(), self.to_right.data_ptr(), self.cap_m, self.cap_g),
                    "fsg_slab_pack")

    def unpack(self, from_left_ptr, from_right_ptr):
        self._check(self._lib.fsg_slab_unpack(self._ctx, from_left_ptr, from_right_ptr, self.cap_m, self.cap_g), "fsg_slab_unpack")

    def setup_peer_exchange(self, overlap: bool = False, classic: bool = False):
        """Maps the neighbours' inboxes into this process (CUDA IPC; one process per GPU on one node).  From
        then on a step copies its messages straight into the neighbours' memory over NVLink; the
        sequence number written last tells the receiver, on the device, that the message is complete — no
        communication library and no host inside a step.  Base-model contexts run the sorted-ghost pipeline
        (fsg_slab2.cu: migrants before the sort, the sorted face layers after the reorder) unless `classic` or
        `overlap` asks for the one whose ghosts travel through the sort."""
        ex = self.exchange
        self._check(self._lib.fsg_slab_alloc_messages2(self._ctx, self.cap_m, self.cap_g, 1 if (classic or overlap) else 0),
                    "fsg_slab_alloc_messages2")
        mine = {}
        for side in (0, 1):
            for par in (0, 1):
                buf = (C.c_ubyte * 64)()
                self._check(self._lib.fsg_slab_inbox_handle(self._ctx, side, par, buf), "fsg_slab_inbox_handle")
                mine[(side, par)] = bytes(buf)
        every = [None] * ex.world
        ex.dist.all_gather_object(every, mine, group=ex.group)
        for par in (0, 1):
            if ex.rank > 0:                  # the left neighbour receives from its right (side 1) what I send to my left
                self._check(self._lib.fsg_slab_open_peer(self._ctx, 0, par, every[ex.rank - 1][(1, par)]), "fsg_slab_open_peer")
            if ex.rank < ex.world - 1:
                self._check(self._lib.fsg_slab_open_peer(self._ctx, 1, par, every[ex.rank + 1][(0, par)]), "fsg_slab_open_peer")
        self.peer = True
        if overlap:
            self.set_overlap(True)

    @property
    def mode(self) -> int:
        """1: classic pipeline (ghosts appended and sorted), 2: sorted ghosts."""
        return int(self._lib.fsg_slab_mode(self._ctx))

    def ghost_ms(self) -> float:
        """Mean device milliseconds per step of the ghost exchange inside fsg_step (sorted-ghost pipeline, profiling on)."""
        ms, k = C.c_double(0.0), C.c_int64(0)
        self._check(self._lib.fsg_slab_get_ghost_ms(self._ctx, C.byref(ms), C.byref(k)), "fsg_slab_get_ghost_ms")
        return float(ms.value)

    def set_overlap(self, on: bool = True):
        """Boundary bins first, then the next step's pack + peer copies on a second stream beside the interior bins."""
        self._check(self._lib.fsg_slab_set_overlap(self._ctx, int(on)), "fsg_slab_set_overlap")

    def keep_foreign(self, on: bool = True):
        """on: uploads return a raw download (all slots, `cell` marking the empty ones) and are not filtered by position."""
        self._check(self._lib.fsg_slab_keep_foreign(self._ctx, int(on)), "fsg_slab_keep_foreign")
        self._upload_cell = bool(on)

    def download_slots(self) -> dict:
        """Every slot this context holds, including empty ones (cell == grid^3 + 1): what keep_foreign uploads take back."""
        return FluidSolver.download(self)

    def check(self) -> dict:
        """Synchronises; raises FsgError if a message / the capacity overflowed or a particle left the ghost band."""
        info = (C.c_int64 * 9)()
        self._check(self._lib.fsg_slab_check(self._ctx, C.byref(info)), "fsg_slab_check")
        v = [int(x) for x in info]
        return dict(sent=(v[0], v[1], v[2], v[3]), received=(v[4], v[5], v[6], v[7]), slots_in_use=v[8])

    def step(self, nsteps: int = 1, sync: bool = True):
        torch = self.torch

        def mark():
            if self.time_exchange:
                e = torch.cuda.Event(enable_timing=True)
                e.record(self.tstream)
                self._ev.append(e)
        for _ in range(nsteps):
            mark()
            if self.peer:
                # pack + copy into the neighbours' inboxes (a no-op when the previous fsg_step already issued it: overlap mode)
                self._check(self._lib.fsg_slab_pack_send(self._ctx), "fsg_slab_pack_send")
                mark()
                mark()
                self._check(self._lib.fsg_slab_unpack_recv(self._ctx), "fsg_slab_unpack_recv")     # waits for the neighbours' stamps on the device
            else:
                self.pack()
                mark()
                with torch.cuda.stream(self.tstream):   # P2P ops are ordered after the pack kernels on the solver's stream
                    self.exchange.exchange(self.to_left, self.to_right, self.from_left, self.from_right, self.msg_bytes)
                mark()
                self.unpack(self.from_left.data_ptr(), self.from_right.data_ptr())
            mark()
            self._check(self._lib.fsg_step(self._ctx, 1), "fsg_step")
            self.steps_exchanged += 1
        if sync:
            self.sync()

    def exchange_ms(self) -> dict:
        """Mean device milliseconds per step of pack / exchange (incl. waiting for the neighbours) / unpack."""
        self.sync()
        ev, out = self._ev, {"pack": 0.0, "exchange": 0.0, "unpack": 0.0}
        k = len(ev) // 4
        for i in range(k):
            a, b, c, d = ev[4 * i:4 * i + 4]
            out["pack"] += a.elapsed_time(b)
            out["exchange"] += b.elapsed_time(c)
            out["unpack"] += c.elapsed_time(d)
        self._ev = []
        return {q: v / max(1, k) for q, v in out.items()}

    @property
    def wire_bytes_per_step(self) -> int:
        """Bytes this rank sends per step: fixed-size messages to its 1 or 2 neighbours; on the sorted-ghost pipeline the
        fixed-size migrant part + the ghosts actually sent in the last step (36 B each: posd, velp, key)."""
        nb = (self.cfg.rank > 0) + (self.cfg.rank < self.cfg.world - 1)
        if self.mode == 2:
            sent = self.check()["sent"]
            return (128 + 80 * self.cap_m) * nb + 36 * (sent[1] + sent[3])
        return self.msg_bytes * nb

    def download(self, fields=None) -> dict:
        """Only the particles this slab owns (ghost / migrated / unused slots are dropped)."""
        out = super().download()
        keep = out["cell"] <= self.numcells
        return {k: v[keep] for k, v in out.items() if fields is None or k in fields}

    def owned_count(self) -> int:
        return int((super().download(("cell",))["cell"] <= self.numcells).sum())


# ---------------------------------------------------------------------------------------------
# W slabs in one process on one device (tests; single-GPU emulation of the multi-rank algorithm)
# ---------------------------------------------------------------------------------------------
class SlabGroup:
    def __init__(self, base_cfg: FsgConfig, world: int, cuts, capacity: int, device: int = 0, cap_m: int = 4096, cap_g: int = 65536,
                 peer: bool = False, overlap: bool = False, classic: bool = False):
        self.world = world
        self.cuts = cuts
        self.peer = peer
        self.slabs = [SlabSolver(slab_config(base_cfg, r, world, cuts, capacity, device), None, cap_m, cap_g) for r in range(world)]
        if peer:         # the peer-memory protocol with plain pointers instead of IPC mappings (same process)
            for s in self.slabs:
                s._check(s._lib.fsg_slab_alloc_messages2(s._ctx, cap_m, cap_g, 1 if (classic or overlap) else 0), "fsg_slab_alloc_messages2")
            for r, s in enumerate(self.slabs):
                for par in (0, 1):
                    if r > 0:
                        s._check(s._lib.fsg_slab_set_peer(s._ctx, 0, par, self.slabs[r - 1]._lib.fsg_slab_inbox_ptr(self.slabs[r - 1]._ctx, 1, par)), "fsg_slab_set_peer")
                    if r < world - 1:
                        s._check(s._lib.fsg_slab_set_peer(s._ctx, 1, par, self.slabs[r + 1]._lib.fsg_slab_inbox_ptr(self.slabs[r + 1]._ctx, 0, par)), "fsg_slab_set_peer")
                s.peer = True
                if overlap:
                    s.set_overlap(True)
                if s.mode == 2:     # one device runs every slab: first halves of all steps, then the second halves
                    s._check(s._lib.fsg_slab_set_split_step(s._ctx, 1), "fsg_slab_set_split_step")

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
            if self.peer:
                # every stamp is in place before a wait kernel is launched: nothing spins on this one device
                for s in self.slabs:
                    s._check(s._lib.fsg_slab_pack_send(s._ctx), "fsg_slab_pack_send")
                for s in self.slabs:
                    s.sync()
                for s in self.slabs:
                    s._check(s._lib.fsg_slab_unpack_recv(s._ctx), "fsg_slab_unpack_recv")
                for s in self.slabs:
                    s.sync()
                for s in self.slabs:
                    s._check(s._lib.fsg_step(s._ctx, 1), "fsg_step")
                for s in self.slabs:
                    s.sync()
                for s in self.slabs:      # (a no-op unless the step was split: sorted-ghost pipeline)
                    s._check(s._lib.fsg_slab_step_finish(s._ctx), "fsg_slab_step_finish")
                for s in self.slabs:
                    s.sync()
                continue
            for s in self.slabs:
                s.pack()
            for s in self.slabs:
                s.sync()             # messages are read by the neighbour's stream
            for r, s in enumerate(self.slabs):
                left, right = (self.slabs[r - 1] if r > 0 else None), (self.slabs[r + 1] if r < self.world - 1 else None)
                s.unpack(left.to_right.data_ptr() if left else None, right.to_left.data_ptr() if right else None)
            for s in self.slabs:
                s.sync()
            for s in self.slabs:
                s._check(s._lib.fsg_step(s._ctx, 1), "fsg_step")
            for s in self.slabs:
                s.sync()

    def check(self) -> list:
        return [s.check() for s in self.slabs]

    def download(self) -> dict:
        parts = [s.download() for s in self.slabs]
        return {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}


# ---------------------------------------------------------------------------------------------
# N slab processes against one context, through the real transport (bench.py `slab_parity`, tests/test_multi_gpu.py)
# ---------------------------------------------------------------------------------------------
def parity_against_single(grid: int, rank: int, world: int, device: int, exchange: str = "peer", steps: int = 4, pair_mode: int = 0,
                          spacing: float = 0.05, jitter: float = 0.005, seed: int = 20261018, drift: float = 25.0,
                          free_steps: int = 6) -> dict | None:
    """Runs the plume scene at grid^3 bins on `world` slab PROCESSES (this is one of them; torch.distributed is initialised,
    one GPU per rank) over the `exchange` transport ('peer': CUDA-IPC inboxes + device-side stamps, 'nccl': send/recv) and,
    on rank 0, on a single context.  Every step starts from identical bits (the single context is re-uploaded from the slabs'
    gathered state, like tests/test_parity_gpu.py::test_slabs_match_single_device does in one process), so that positions,
    velocities, bin ids and boundary flags must agree bit for bit and the pair sums to rounding.  A common drift along x
    (drift * DT * steps >= one lattice spacing, so some lattice plane crosses every slab face; still far below one bin layer per
    step) makes particles migrate.  Then `free_steps` steps WITHOUT any download in between — on the sorted-ghost pipeline the update
    stays deferred and migrants travel with their pre-update state and pending pair sums — against the same steps on the single
    context: same particle set, trajectories equal to rounding (`free_running`).  Returns the record on rank 0, None elsewhere."""
    import torch.distributed as dist
    from . import scenes
    from .solver import by_index

    base = scenes.plume_config(grid)
    base.pair_mode = pair_mode
    base.device = device
    state = scenes.plume_scene(base, spacing, jitter, seed)
    state["vel"][:, 0] += np.float32(drift)
    n = state["pos"].shape[0]
    hist = layer_hist_from_positions(base, state["pos"])
    cuts = slab_cuts(hist, world)
    owned = [int(hist[a:b].sum()) for a, b in cuts]
    # (the upload hands every rank the whole scene: until the first sort trims the foreign slots, arrivals are appended behind all n)
    cap = max(int(max(owned) * 1.1), n) + 4 * int(hist.max()) + 65536
    cap_m, cap_g = message_caps(hist, cuts)
    cap_m = cap_g                     # with the drift whole lattice planes cross a face within one step
    ex = DistExchange()
    s = SlabSolver(slab_config(base, rank, world, cuts, cap, device), ex, cap_m, cap_g)
    single = None
    try:
        if exchange == "peer":
            s.setup_peer_exchange()
        s.upload(state)                       # every rank is handed the whole scene and keeps its own slab
        if rank == 0:
            one = scenes.plume_config(grid, n)
            one.pair_mode = pair_mode
            one.device = device
            single = FluidSolver(one)
        rec = {"grid": grid, "particles": n, "world": world, "exchange": exchange, "pair_mode": pair_mode, "steps": steps,
               "bit_exact": True, "cells_equal_frac": 1.0, "max_rel_l2": 0.0, "migrated": 0, "conserved": True}
        ints = ("pos", "vel", "cell", "boundary")
        flds = ("acc", "dens", "press", "delpress")

        def gather():
            mine = s.download()
            parts = [None] * world if rank == 0 else None
            dist.gather_object(mine, parts, dst=0)
            if rank != 0:
                return None
            return by_index({k: np.concatenate([p[k] for p in parts]) for k in parts[0]})

        for _ in range(steps):
            cur = gather()
            if rank == 0:
                rec["conserved"] &= bool(cur["index"].shape[0] == n and np.array_equal(cur["index"], np.arange(n)))
                single.upload({f: v for f, v in cur.items() if f != "cell"})
            s.step(1)
            info = s.check()
            rec["migrated"] += ex.global_sum([info["sent"][0] + info["sent"][2]], s.tdev)[0]
            if rank == 0:
                single.step(1)
            a = gather()
            if rank == 0:
                b = by_index(single.download())
                ok_n = a["index"].shape[0] == b["index"].shape[0] and np.array_equal(a["index"], b["index"])
                rec["conserved"] &= bool(ok_n)
                if not ok_n:
                    rec["bit_exact"] = False
                    break
                for f in ints:
                    same = np.array_equal(a[f].view(np.uint8) if a[f].dtype == np.uint8 else a[f], b[f])
                    rec["bit_exact"] &= bool(same)
                rec["cells_equal_frac"] = min(rec["cells_equal_frac"], float((a["cell"] == b["cell"]).mean()))
                for f in flds:
                    den = float(np.sqrt((b[f].astype(np.float64) ** 2).sum()))
                    err = float(np.sqrt(((a[f].astype(np.float64) - b[f].astype(np.float64)) ** 2).sum()))
                    rec["max_rel_l2"] = max(rec["max_rel_l2"], err / den if den > 0 else err)
        if free_steps > 0 and rec["bit_exact"]:
            cur = gather()
            if rank == 0:
                single.upload({f: v for f, v in cur.items() if f != "cell"})
            sent = 0
            for _ in range(free_steps):
                s.step(1)
                info = s.check()                      # (reads counters only: nothing is materialised)
                sent += info["sent"][0] + info["sent"][2]
            migrated = ex.global_sum([sent], s.tdev)[0]
            if rank == 0:
                single.step(free_steps)
            a = gather()
            if rank == 0:
                b = by_index(single.download())
                fr = {"steps": free_steps, "migrated": migrated, "pipeline": "sorted ghosts" if s.mode == 2 else "classic",
                      "conserved": bool(a["index"].shape[0] == n and np.array_equal(a["index"], b["index"])), "max_rel_l2": 0.0}
                if fr["conserved"]:
                    fr["cells_equal_frac"] = float((a["cell"] == b["cell"]).mean())
                    for f in ("pos", "vel") + flds:
                        den = float(np.sqrt((b[f].astype(np.float64) ** 2).sum()))
                        err = float(np.sqrt(((a[f].astype(np.float64) - b[f].astype(np.float64)) ** 2).sum()))
                        fr["max_rel_l2"] = max(fr["max_rel_l2"], err / den if den > 0 else err)
                rec["free_running"] = fr
        return rec if rank == 0 else None
    finally:
        s.close()
        if single is not None:
            single.close()
