"""Host-side mirror of the reference drivers' time loop (solver.cu:171-216) over the C ABI of libfsg.

    s = FluidSolver(FluidSolver.base_config())     # FluidGPU.cuh:1-31 constants
    s.upload(scenes.base_default_scene())          # solver.cu:115-131
    s.step(100)                                    # solver.cu:181-198, 100 times
    out = s.download()                             # like cudaMemcpy(SPptr, d_SPptr, ...), bin-sorted order

numpy in, numpy out; all compute happens in the CUDA library.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import FsgConfig, FsgError, FsgSoa, FsgStats

_F3 = ("pos", "vel", "acc", "delpress", "newdelpress")
_F1 = ("dens", "press", "newdens")
_FU = ("solid", "fluid", "mass")      # unidyn model only (mass != 1 needs fsg_config.unidyn_adapt)
_F9 = ("stress_tensor", "stress_rate")      # unidyn, granular state [n, 9]


class FluidSolver:
    def __init__(self, cfg: FsgConfig):
        self._lib = _lib.load()
        self._ctx = C.c_void_p()
        self.cfg = cfg
        rc = self._lib.fsg_create(C.byref(cfg), C.byref(self._ctx))
        if rc != 0:
            msg = self._lib.fsg_last_error(None)
            self._ctx = None
            raise FsgError(rc, "fsg_create", msg.decode() if msg else "")

    # ---- configuration helpers ----
    @staticmethod
    def base_config(capacity: int = 8000, device: int = 0, **kw) -> FsgConfig:
        cfg = FsgConfig()
        _lib.load().fsg_config_default(C.byref(cfg), _lib.FSG_MODEL_BASE)
        cfg.capacity = capacity
        cfg.device = device
        for k, v in kw.items():
            setattr(cfg, k, v)
        return cfg

    @staticmethod
    def unidyn_config(capacity: int = 14040, device: int = 0, **kw) -> FsgConfig:
        """FluidGPU-unidyn.cuh:1-36 constants (GRIDSIZE 17, CELLSIZE 0.12, DT 0.0018, ...)."""
        cfg = FsgConfig()
        _lib.load().fsg_config_default(C.byref(cfg), _lib.FSG_MODEL_UNIDYN)
        cfg.capacity = capacity
        cfg.device = device
        for k, v in kw.items():
            setattr(cfg, k, v)
        return cfg

    @property
    def is_unidyn(self) -> bool:
        return self.cfg.model == _lib.FSG_MODEL_UNIDYN

    @property
    def numcells(self) -> int:
        return self.cfg.grid ** 3

    def _check(self, rc: int, where: str):
        if rc != 0:
            msg = self._lib.fsg_last_error(self._ctx)
            raise FsgError(rc, where, msg.decode() if msg else "")

    def close(self):
        if self._ctx:
            self._lib.fsg_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- data movement ----
    def upload(self, state: dict):
        n = state["pos"].shape[0]
        soa = FsgSoa()
        soa.n = n
        keep = []
        for k in _F3 + _F1 + (_FU + _F9 if self.is_unidyn else ()):
            if state.get(k) is not None:
                a = np.ascontiguousarray(state[k], np.float32)
                keep.append(a)
                setattr(soa, k, a.ctypes.data)
        if state.get("index") is not None:
            a = np.ascontiguousarray(state["index"], np.int32); keep.append(a); soa.index = a.ctypes.data
        if state.get("boundary") is not None:
            a = np.ascontiguousarray(state["boundary"], np.uint8); keep.append(a); soa.boundary = a.ctypes.data
        if getattr(self, "_upload_cell", False) and state.get("cell") is not None:      # slab contexts, fsg_slab_keep_foreign
            a = np.ascontiguousarray(state["cell"], np.int32); keep.append(a); soa.cell = a.ctypes.data
        self._check(self._lib.fsg_upload_soa(self._ctx, C.byref(soa)), "fsg_upload_soa")
        self._check(self._lib.fsg_sync(self._ctx), "fsg_sync")

    def upload_raw(self, soa: FsgSoa):
        self._check(self._lib.fsg_upload_soa(self._ctx, C.byref(soa)), "fsg_upload_soa")

    def download_raw(self, soa: FsgSoa):
        self._check(self._lib.fsg_download_soa(self._ctx, C.byref(soa)), "fsg_download_soa")

    def download(self, fields=None) -> dict:
        st = self.stats()
        n = st["n"]
        out = {}
        soa = FsgSoa()
        fields = fields or (_F3 + _F1 + ("index", "cell", "boundary") + (_FU + _F9 if self.is_unidyn else ()))
        for k in fields:
            if k in _F3:
                out[k] = np.empty((n, 3), np.float32)
            elif k in _F9:
                out[k] = np.empty((n, 9), np.float32)
            elif k in _F1 or k in _FU:
                out[k] = np.empty(n, np.float32)
            elif k in ("index", "cell"):
                out[k] = np.empty(n, np.int32)
            else:
                out[k] = np.empty(n, np.uint8)
            setattr(soa, k, out[k].ctypes.data)
        self._check(self._lib.fsg_download_soa(self._ctx, C.byref(soa)), "fsg_download_soa")
        return out

    def upload_aos(self, records: np.ndarray):
        records = np.ascontiguousarray(records, np.uint8).reshape(-1, _lib.FSG_AOS_STRIDE)
        self._check(self._lib.fsg_upload_aos(self._ctx, records.ctypes.data, records.shape[0]), "fsg_upload_aos")

    def download_aos(self) -> np.ndarray:
        n = self.stats()["n"]
        out = np.empty((n, _lib.FSG_AOS_STRIDE), np.uint8)
        self._check(self._lib.fsg_download_aos(self._ctx, out.ctypes.data, n), "fsg_download_aos")
        return out

    # ---- the step ----
    def step(self, nsteps: int = 1, sync: bool = True):
        self._check(self._lib.fsg_step(self._ctx, nsteps), "fsg_step")
        if sync:
            self._check(self._lib.fsg_sync(self._ctx), "fsg_sync")

    def sync(self):
        self._check(self._lib.fsg_sync(self._ctx), "fsg_sync")

    def stream(self) -> int:
        return self._lib.fsg_get_stream(self._ctx) or 0

    def export_viz(self):
        n = self.stats()["n"]
        spts, a3, b3 = np.empty(3 * n, np.float32), np.empty(n, np.float32), np.empty(n, np.float32)
        self._check(self._lib.fsg_export_viz(self._ctx, spts.ctypes.data, a3.ctypes.data, b3.ctypes.data), "fsg_export_viz")
        return spts, a3, b3

    def write_frame(self, filename: str, binary: bool = False):
        """The VTK frame the drivers write after a step (solver-unidyn.cu:472-493), byte-compatible with visit_writer."""
        self._check(self._lib.fsg_write_frame(self._ctx, str(filename).encode(), int(binary)), "fsg_write_frame")

    def write_frame_async(self, filename: str, binary: bool = False):
        """write_frame off the critical path: export kernel + copy stream + writer thread (fsg_write_frame_async)."""
        self._check(self._lib.fsg_write_frame_async(self._ctx, str(filename).encode(), int(binary)), "fsg_write_frame_async")

    def frame_wait(self) -> int:
        """Blocks until every asynchronous frame is on disk; returns how many have been written since the context was created."""
        n = C.c_int64(0)
        self._check(self._lib.fsg_frame_wait(self._ctx, C.byref(n)), "fsg_frame_wait")
        return n.value

    def tables(self):
        n = self.stats()["n"]
        cells = np.empty(n, np.int32)
        start = np.empty(self.numcells, np.int32)
        end = np.empty(self.numcells, np.int32)
        self._check(self._lib.fsg_get_tables(self._ctx, cells.ctypes.data, start.ctypes.data, end.ctypes.data), "fsg_get_tables")
        return cells, start, end

    def split(self) -> np.ndarray:
        """unidyn: split[] of the last step (FluidGPU-unidyn.cu:181-190)."""
        out = np.empty(self.numcells, np.int32)
        self._check(self._lib.fsg_get_split(self._ctx, out.ctypes.data), "fsg_get_split")
        return out

    def adapt_counts(self) -> dict:
        """unidyn_adapt contexts: (pairs merged, particles split, children created) of the last step and since the upload."""
        last, total = (C.c_int64 * 3)(), (C.c_int64 * 3)()
        self._check(self._lib.fsg_unidyn_adapt_counts(self._ctx, C.byref(last), C.byref(total)), "fsg_unidyn_adapt_counts")
        return {"last": tuple(int(v) for v in last), "total": tuple(int(v) for v in total)}

    def stats(self) -> dict:
        st = FsgStats()
        self._check(self._lib.fsg_get_stats(self._ctx, C.byref(st)), "fsg_get_stats")
        return {k: getattr(st, k) for k, _ in FsgStats._fields_}

    def pair_stats_one_step(self) -> dict:
        """Takes ONE extra step with the pair counters on and returns that step's statistics."""
        self._check(self._lib.fsg_set_collect_stats(self._ctx, 1), "fsg_set_collect_stats")
        self.step(1)
        st = self.stats()
        self._check(self._lib.fsg_set_collect_stats(self._ctx, 0), "fsg_set_collect_stats")
        return st

    def set_profiling(self, on: bool = True):
        self._check(self._lib.fsg_set_profiling(self._ctx, int(on)), "fsg_set_profiling")

    def phase_ms(self) -> dict:
        """Milliseconds per phase accumulated since the last call (CUDA events on the solver's stream)."""
        ms = (C.c_double * 4)()
        steps = C.c_int64(0)
        self._check(self._lib.fsg_get_phase_ms(self._ctx, C.byref(ms), C.byref(steps)), "fsg_get_phase_ms")
        return dict(sort=ms[0], reorder=ms[1], pair_update=ms[2], other=ms[3], steps=steps.value)

    def scene_plume(self, spacing: float = 0.05, jitter: float = 0.005, seed: int = 20261018) -> int:
        n = C.c_int64(0)
        self._check(self._lib.fsg_scene_plume(self._ctx, spacing, jitter, seed, C.byref(n)), "fsg_scene_plume")
        return n.value

    # ---- stage API: caller-owned DEVICE buffers in the reference's own layout (device pointers as ints) ----
    def stage_sort(self, d_cells: int, d_particles: int, n: int):
        self._check(self._lib.fsg_stage_sort(self._ctx, d_cells, d_particles, n), "fsg_stage_sort")

    def stage_findneighbours(self, d_cells: int, d_start: int, d_end: int, n: int):
        self._check(self._lib.fsg_stage_findneighbours(self._ctx, d_cells, d_start, d_end, n), "fsg_stage_findneighbours")

    def stage_mykernel(self, d_particles: int, d_cells: int, d_start: int, d_end: int, n: int):
        self._check(self._lib.fsg_stage_mykernel(self._ctx, d_particles, d_cells, d_start, d_end, n), "fsg_stage_mykernel")

    def stage_mykernel2(self, d_particles: int, d_cells: int, d_start: int, d_end: int, n: int, spts: int = 0, a3: int = 0, b3: int = 0):
        self._check(self._lib.fsg_stage_mykernel2(self._ctx, d_particles, d_cells, d_start, d_end, n, spts or None, a3 or None,
                                                  b3 or None), "fsg_stage_mykernel2")

    def device_ptr(self, which: int) -> int:
        p = C.c_void_p()
        self._check(self._lib.fsg_device_ptr(self._ctx, which, C.byref(p)), "fsg_device_ptr")
        return p.value or 0


def by_index(state: dict) -> dict:
    """Re-orders a downloaded (bin-sorted) state by Particle::index so that runs can be compared."""
    order = np.argsort(state["index"], kind="stable")
    return {k: v[order] for k, v in state.items()}


def write_point_mesh(filename, pts: np.ndarray, variables: dict, binary: bool = False):
    """fsg_write_point_mesh from numpy: pts [n,3] float32, variables = {name: [n] or [n,3] float32} (insertion order)."""
    lib = _lib.load()
    pts = np.ascontiguousarray(pts, np.float32)
    n = pts.shape[0]
    arrs = [np.ascontiguousarray(v, np.float32) for v in variables.values()]
    dims = (C.c_int * len(arrs))(*[1 if a.ndim == 1 else a.shape[1] for a in arrs])
    names = (C.c_char_p * len(arrs))(*[k.encode() for k in variables])
    ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    rc = lib.fsg_write_point_mesh(str(filename).encode(), int(binary), n, pts.ctypes.data, len(arrs), dims, names, ptrs)
    if rc != 0:
        raise FsgError(rc, "fsg_write_point_mesh", str(filename))
