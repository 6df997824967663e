"""fluidsolvergpu_b200 — B200-native (sm_100a) implementation of FluidSolverGPU's per-timestep
particle update, behind the C ABI in include/fsg.h.  This package is the thin host-side mirror of
the reference drivers' loop; every compute stage is a CUDA kernel in libfsg.so (no CPU fallback)."""
from . import _lib, scenes, sections, slab  # noqa: F401
from ._lib import FsgConfig, FsgError, FsgSoa, FsgStats  # noqa: F401
from .solver import FluidSolver, by_index, write_point_mesh  # noqa: F401
from .slab import DistExchange, SlabGroup, SlabSolver, slab_config, slab_cuts  # noqa: F401

__all__ = ["DistExchange", "SlabGroup", "SlabSolver", "slab_config", "slab_cuts", "FluidSolver", "FsgConfig", "FsgError", "FsgSoa", "FsgStats", "by_index", "scenes", "sections"]
