"""ctypes front-end of oracle/liboracle.so — the CPU restatement of the reference (TEST INFRASTRUCTURE).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this."""
from __future__ import annotations

import ctypes as C
import pathlib
import subprocess

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
LIB = ROOT / "oracle" / "liboracle.so"


class Params(C.Structure):
    _fields_ = [("grid", C.c_int), ("origin", C.c_float), ("cellsize", C.c_double), ("h", C.c_double), ("dt", C.c_double),
                ("alpha_fluid", C.c_double), ("alpha_boundary", C.c_double), ("sound", C.c_double), ("gravity", C.c_double),
                ("block_threads", C.c_int), ("bin_cap", C.c_int), ("threads", C.c_int)]


class State(C.Structure):
    _fields_ = [("n", C.c_int)] + [(k, C.c_void_p) for k in ("pos", "vel", "acc", "dens", "press", "delpress", "newdens",
                                                              "newdelpress", "index", "cell", "boundary")]


class UState(C.Structure):
    _fields_ = [("n", C.c_int)] + [(k, C.c_void_p) for k in ("pos", "vel", "acc", "dens", "press", "delpress", "newdens",
                                                              "newdelpress", "index", "cell", "boundary", "solid", "fluid", "subindex",
                                                              "stress_tensor", "stress_rate", "mass")]


class Adapt(C.Structure):
    _fields_ = [("merge_distance", C.c_double), ("split_mass_min", C.c_double), ("capacity", C.c_int), ("next_index", C.c_int),
                ("merged", C.c_int), ("split", C.c_int), ("added", C.c_int)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", str(ROOT / "oracle"), "liboracle.so"])


def load():
    global _lib
    if _lib is None:
        srcs = list((ROOT / "oracle").glob("fsg_oracle*.c")) + [ROOT / "oracle" / "fsg_oracle.h"]
        if not LIB.exists() or any(s.stat().st_mtime > LIB.stat().st_mtime for s in srcs):
            build()
        lib = C.CDLL(str(LIB))
        for f in ("fsgo_kernel", "fsgo_kernel_test", "fsgo_kernel_derivative", "fsgo_pressure"):
            getattr(lib, f).restype = C.c_float
            getattr(lib, f).argtypes = [C.c_float]
        lib.fsgo_set_dens.restype = C.c_float
        lib.fsgo_set_dens.argtypes = [C.c_float, C.c_int]
        lib.fsgo_cell_id.restype = C.c_int
        lib.fsgo_cell_id.argtypes = [C.POINTER(Params), C.c_float, C.c_float, C.c_float]
        lib.fsgo_params_base.argtypes = [C.POINTER(Params)]
        lib.fsgo_base_step.restype = C.c_int
        lib.fsgo_base_step.argtypes = [C.POINTER(Params), C.POINTER(State)] + [C.c_void_p] * 7
        lib.fsgo_params_unidyn.argtypes = [C.POINTER(Params)]
        lib.fsgo_unidyn_step.restype = C.c_int
        lib.fsgo_unidyn_step.argtypes = [C.POINTER(Params), C.POINTER(UState), C.c_int] + [C.c_void_p] * 8
        lib.fsgo_unidyn_step_adapt.restype = C.c_int
        lib.fsgo_unidyn_step_adapt.argtypes = [C.POINTER(Params), C.POINTER(UState), C.POINTER(Adapt), C.c_int] + [C.c_void_p] * 8
        _lib = lib
    return _lib


def base_params(**kw) -> Params:
    p = Params()
    load().fsgo_params_base(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def params_from_cfg(cfg, threads=0) -> Params:
    """Oracle parameters matching an fsg_config (base model)."""
    return base_params(grid=cfg.grid, origin=cfg.origin, cellsize=cfg.cellsize, h=cfg.h, dt=cfg.dt,
                       alpha_fluid=cfg.alpha_fluid, alpha_boundary=cfg.alpha_boundary, sound=cfg.sound,
                       gravity=cfg.gravity, block_threads=cfg.neighbour_cap, bin_cap=cfg.bin_cap, threads=threads)


def unidyn_params(**kw) -> Params:
    p = Params()
    load().fsgo_params_unidyn(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class OracleSimUnidyn:
    """Runs fsgo_unidyn_step on a copy of a state dict that also holds solid / fluid.  adapt = (merge_distance, split_mass_min,
    capacity): fsgo_unidyn_step_adapt — particle merging / splitting, the state also carries `mass` and may grow up to `capacity`."""

    def __init__(self, params: Params, state: dict, adapt=None):
        self.lib = load()
        self.p = params
        self.s = {k: np.array(v, copy=True) for k, v in state.items()}
        n = self.s["pos"].shape[0]
        self.n = n
        self.s["cell"] = np.clip(cell_ids(self.p, self.s["pos"]), -1, params.grid ** 3).astype(np.int32)
        self.s.setdefault("subindex", np.zeros(n, np.int32))
        self.s.setdefault("stress_tensor", np.zeros((n, 9), np.float32))
        self.s.setdefault("stress_rate", np.zeros((n, 9), np.float32))
        self.adapt = None
        if adapt is not None:
            self.s.setdefault("mass", np.ones(n, np.float32))
            self.adapt = Adapt(merge_distance=adapt[0], split_mass_min=adapt[1], capacity=int(adapt[2]), next_index=n)
            cap = int(adapt[2])
            for k, v in list(self.s.items()):          # room for the children behind the n particles
                pad = np.zeros((cap - n,) + v.shape[1:], v.dtype)
                self.s[k] = np.ascontiguousarray(np.concatenate([v, pad]))
            n = cap
        nc = params.grid ** 3
        self.cells_sorted = np.zeros(n, np.int32)
        self.start, self.end, self.split = np.zeros(nc, np.int32), np.zeros(nc, np.int32), np.zeros(nc, np.int32)
        self.spts, self.a3, self.b3 = np.zeros(3 * n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
        self.stats = np.zeros(4, np.int64)
        self.t = 0

    def step(self, nsteps=1):
        st = UState()
        st.n = self.n
        for k in ("pos", "vel", "acc", "dens", "press", "delpress", "newdens", "newdelpress", "index", "cell", "boundary", "solid",
                  "fluid", "subindex", "stress_tensor", "stress_rate") + (("mass",) if self.adapt is not None else ()):
            setattr(st, k, self.s[k].ctypes.data)
        out = (self.cells_sorted.ctypes.data, self.start.ctypes.data, self.end.ctypes.data, self.split.ctypes.data, self.spts.ctypes.data,
               self.a3.ctypes.data, self.b3.ctypes.data, self.stats.ctypes.data)
        self.events = []
        for _ in range(nsteps):
            if self.adapt is not None:
                rc = self.lib.fsgo_unidyn_step_adapt(C.byref(self.p), C.byref(st), C.byref(self.adapt), self.t, *out)
                self.events.append((self.adapt.merged, self.adapt.split, self.adapt.added))
            else:
                rc = self.lib.fsgo_unidyn_step(C.byref(self.p), C.byref(st), self.t, *out)
            assert rc == 0, rc
            self.n = st.n
            self.t += 1
        return self

    def state(self) -> dict:
        return {k: v[:self.n].copy() for k, v in self.s.items()}


class OracleSim:
    """Runs fsgo_base_step on a copy of a state dict (see fluidsolvergpu_b200.scenes)."""

    def __init__(self, params: Params, state: dict):
        self.lib = load()
        self.p = params
        self.s = {k: np.array(v, copy=True) for k, v in state.items()}
        n = self.s["pos"].shape[0]
        self.n = n
        pos = self.s["pos"]
        self.s["cell"] = np.clip(cell_ids(self.p, pos), -1, params.grid ** 3).astype(np.int32)
        nc = params.grid ** 3
        self.cells_sorted = np.zeros(n, np.int32)
        self.start = np.zeros(nc, np.int32)
        self.end = np.zeros(nc, np.int32)
        self.spts = np.zeros(3 * n, np.float32)
        self.a3 = np.zeros(n, np.float32)
        self.b3 = np.zeros(n, np.float32)
        self.stats = np.zeros(4, np.int64)

    def step(self, nsteps=1):
        st = State()
        st.n = self.n
        for k in ("pos", "vel", "acc", "dens", "press", "delpress", "newdens", "newdelpress", "index", "cell", "boundary"):
            setattr(st, k, self.s[k].ctypes.data)
        for _ in range(nsteps):
            rc = self.lib.fsgo_base_step(C.byref(self.p), C.byref(st), self.cells_sorted.ctypes.data, self.start.ctypes.data,
                                         self.end.ctypes.data, self.spts.ctypes.data, self.a3.ctypes.data, self.b3.ctypes.data,
                                         self.stats.ctypes.data)
            assert rc == 0
        return self

    def state(self) -> dict:
        return {k: v.copy() for k, v in self.s.items()}


def cell_ids(p: Params, pos: np.ndarray) -> np.ndarray:
    """fsgo_cell_id (FluidGPU.cu:419) over an array of positions; int64 so that out-of-grid ids show."""
    lib = load()
    pos = np.asarray(pos, np.float32)
    q = (pos - np.float32(p.origin)).astype(np.float32).astype(np.float64) / p.cellsize
    q = np.trunc(q).astype(np.int64)
    ids = q[:, 0] * p.grid * p.grid + q[:, 1] * p.grid + q[:, 2]
    if len(pos):   # spot-check the vectorised form against the C restatement
        for i in np.linspace(0, len(pos) - 1, min(len(pos), 64)).astype(int):
            c = lib.fsgo_cell_id(C.byref(p), float(pos[i, 0]), float(pos[i, 1]), float(pos[i, 2]))
            assert abs(ids[i]) >= 2 ** 31 or c == ids[i] or abs(q[i]).max() >= 2 ** 31
    return ids


def rel_l2(a, b) -> float:
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.sqrt((b * b).sum())
    return float(np.sqrt(((a - b) ** 2).sum()) / den) if den > 0 else float(np.sqrt(((a - b) ** 2).sum()))
