"""Scenes of the reference drivers and the synthetic throughput scene, as numpy state dicts.

A state dict holds float32 arrays pos/vel/acc [n,3], dens/press/newdens [n], delpress/newdelpress [n,3],
int32 index [n], uint8 boundary [n] — the live fields of `class Particle` (FluidGPU.cuh:112-162).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

RHO_0 = 9550.0
GRAVITY = -9.8


def default_state(pos: np.ndarray, vel: np.ndarray | None = None, boundary: np.ndarray | None = None) -> dict:
    """Class defaults of Particle (FluidGPU.cuh:64-71, 132-148) around given positions."""
    n = pos.shape[0]
    b = np.zeros(n, np.uint8) if boundary is None else boundary.astype(np.uint8)
    acc = np.zeros((n, 3), np.float32)
    acc[b == 0, 2] = np.float32(GRAVITY)
    return dict(
        pos=np.ascontiguousarray(pos, np.float32),
        vel=np.zeros((n, 3), np.float32) if vel is None else np.ascontiguousarray(vel, np.float32),
        acc=acc,
        dens=np.full(n, RHO_0, np.float32),
        press=np.zeros(n, np.float32),
        delpress=np.zeros((n, 3), np.float32),
        newdens=np.full(n, RHO_0, np.float32),      # FluidGPU.cuh:144 — not zero on the first step
        newdelpress=np.zeros((n, 3), np.float32),
        index=np.arange(n, dtype=np.int32),
        boundary=b,
    )


def base_default_scene() -> dict:
    """solver.cu:115-121 — 8000 fluid particles on a 15 x 36 x 15 lattice, spacing 0.04, no boundary."""
    j = np.arange(8000)
    x = -.16 + 0.04 * ((j // 15) % 15)
    y = -0.76 + 0.04 * (j // 15 // 15)
    z = -0.20 + (j % 15) * 0.04
    return default_state(np.stack([x, y, z], 1).astype(np.float32))


def random_base_scene(n: int, seed: int, box=((-0.3, 0.3), (-0.3, 0.3), (-0.3, 0.3)), spacing: float = 0.04,
                      jitter: float = 0.01, vel_scale: float = 0.2, boundary_frac: float = 0.0) -> dict:
    """Seeded jittered lattice inside the reference's [-1,1]^3 domain (keeps the neighbourhoods near the
    densities the reference's 64-thread cap was tuned to)."""
    rng = np.random.default_rng(seed)
    axes = [np.arange(lo, hi, spacing) for lo, hi in box]
    g = np.stack(np.meshgrid(*axes, indexing="ij"), -1).reshape(-1, 3)
    if g.shape[0] > n:
        g = g[rng.permutation(g.shape[0])[:n]]
    pos = g + rng.uniform(-jitter, jitter, g.shape)
    vel = rng.uniform(-vel_scale, vel_scale, g.shape)
    b = (rng.uniform(size=g.shape[0]) < boundary_frac).astype(np.uint8) if boundary_frac > 0 else None
    return default_state(pos.astype(np.float32), vel.astype(np.float32), b)


def plume_config(grid: int, capacity: int = 0) -> "_lib.FsgConfig":
    """Throughput configuration of SURVEY.md §8d: base physics, CELLSIZE = 2h = 0.12, domain
    [-0.06 G, 0.06 G]^3, no neighbour cap."""
    lib = _lib.load()
    cfg = _lib.FsgConfig()
    lib.fsg_config_default(C.byref(cfg), _lib.FSG_MODEL_BASE)
    cfg.grid = grid
    cfg.cellsize = 0.12
    cfg.origin = -0.06 * grid
    cfg.neighbour_cap = 0
    cfg.bin_cap = 0
    cfg.capacity = capacity
    return cfg


def plume_count(cfg, spacing: float = 0.05) -> int:
    lib = _lib.load()
    n = C.c_int64(0)
    rc = lib.fsg_scene_plume_host(C.byref(cfg), spacing, 0.0, 0, None, None, 0, C.byref(n))
    if rc != 0:
        raise _lib.FsgError(rc, "fsg_scene_plume_host")
    return n.value


def plume_scene(cfg, spacing: float = 0.05, jitter: float = 0.005, seed: int = 20261018) -> dict:
    """Host copy of the plume scene (identical bits to fsg_scene_plume on the device)."""
    lib = _lib.load()
    n = plume_count(cfg, spacing)
    pos = np.empty((n, 3), np.float32)
    vel = np.empty((n, 3), np.float32)
    m = C.c_int64(0)
    rc = lib.fsg_scene_plume_host(C.byref(cfg), spacing, jitter, seed, pos.ctypes.data, vel.ctypes.data, n, C.byref(m))
    if rc != 0:
        raise _lib.FsgError(rc, "fsg_scene_plume_host")
    return default_state(pos, vel)


def unidyn_default_scene() -> dict:
    """solver-unidyn.cu:124-185 — 10 000 fluid particles (30 x 30 x 12 lattice, spacing 0.05) over a floor
    (2 020) and four walls (4 x 505) of boundary particles; solid/fluid = 0/1 for fluid, 1/0 for boundary."""
    nspts, nbpts = 10000, 4040
    j = np.arange(nspts)
    fx = -.76 + 0.05 * ((j // 30) % 30)
    fy = -0.76 + 0.05 * (j % 30)
    fz = -0.40 + (j // 30 // 30) * 0.05
    i = np.arange(nbpts // 2)
    floor = np.stack([-0.96 + 0.04 * (i % 45), -0.96 + 0.04 * (i // 45), np.full(i.shape, -0.7)], 1)
    i = np.arange(nbpts // 8)
    a, b = -0.96 + 0.04 * (i % 45), -0.74 + 0.04 * (i // 45)
    walls = [np.stack([a, np.full(i.shape, -0.96), b], 1), np.stack([a, np.full(i.shape, 0.84), b], 1),
             np.stack([np.full(i.shape, -0.96), a, b], 1), np.stack([np.full(i.shape, 0.76), a, b], 1)]
    pos = np.concatenate([np.stack([fx, fy, fz], 1), floor] + walls).astype(np.float32)
    bnd = np.r_[np.zeros(nspts, np.uint8), np.ones(nbpts, np.uint8)]
    st = default_state(pos, boundary=bnd)
    st["solid"] = bnd.astype(np.float32)
    st["fluid"] = (1 - bnd).astype(np.float32)
    return st


def random_unidyn_scene(n: int, seed: int, boundary_frac: float = 0.1, vel_scale: float = 0.2) -> dict:
    """Seeded jittered lattice in the unidyn domain (CELLSIZE 0.12): fluid particles + boundary particles."""
    st = random_base_scene(n, seed, box=((-0.5, 0.5),) * 3, spacing=0.05, jitter=0.012, vel_scale=vel_scale, boundary_frac=boundary_frac)
    st["solid"] = st["boundary"].astype(np.float32)
    st["fluid"] = (1 - st["boundary"]).astype(np.float32)
    st["vel"][st["boundary"] != 0] = 0
    return st


def mixed_unidyn_scene(n: int = 4000, seed: int = 7, stress_scale: float = 1e3) -> dict:
    """Sand on water for the unidyn model: a seeded jittered lattice in the unidyn domain whose upper part is
    granular material (solid = 1), whose lower part is water (solid = 0) and whose middle band is a mixture (0 < solid < 1,
    fluid = 1 - solid) — the particles that activate the mixed-phase block (FluidGPU-unidyn.cu:317-357), vel_grad / stress_accel /
    mixture_accel / delsolid (:368-401) and the granular stress update (:410-446).  Boundary particles as in the default scene
    (solid = 1, fluid = 0).  stress_tensor / stress_rate start from seeded values on the granular particles so that the stress terms
    are exercised from the first step."""
    st = random_unidyn_scene(n, seed, boundary_frac=0.1, vel_scale=0.3)
    rng = np.random.default_rng(seed + 1)
    z = st["pos"][:, 2]
    solid = np.where(z > 0.1, 1.0, np.where(z > -0.1, rng.uniform(0.05, 0.95, z.shape), 0.0)).astype(np.float32)
    fluid = (1 - solid).astype(np.float32)
    bnd = st["boundary"] != 0
    solid[bnd], fluid[bnd] = 1.0, 0.0
    st["solid"], st["fluid"] = solid, fluid
    m = z.shape[0]
    sym = rng.standard_normal((m, 3, 3))
    sym = (sym + sym.transpose(0, 2, 1)) * 0.5 * stress_scale
    st["stress_tensor"] = (sym.reshape(m, 9) * (solid[:, None] > 0)).astype(np.float32)
    st["stress_rate"] = (rng.standard_normal((m, 9)) * stress_scale * 100 * (solid[:, None] > 0)).astype(np.float32)
    return st
