"""The reference's 340-byte `Particle` record (base model) as a numpy structured view.
Offsets: FluidGPU.cuh:112-162 as laid out by nvcc/gcc (pinned by tests/golden/kat_base.json)."""
import numpy as np

BASE_OFFSETS = dict(pos=0, vel=12, acc=24, index=36, cellnumber=40, mass=44, dens=48, press=52, delpressz=56, delpressy=60,
                    delpressx=64, diffusion=68, sigma=80, newdens=84, newpress=88, newdelpressz=92, newdelpressy=96,
                    newdelpressx=100, newsigma=104, vel_grad=108, strain_rate=144, stress_rate=180, strain_rate_squared=216,
                    stress_tensor=252, stress_tensor_squared=288, stress_accel=324, boundary=336, solid=337, flag=338)
STRIDE = 340

BASE_DTYPE = np.dtype({
    "names": ["pos", "vel", "acc", "index", "cellnumber", "mass", "dens", "press", "delpressz", "delpressy", "delpressx",
              "newdens", "newpress", "newdelpressz", "newdelpressy", "newdelpressx", "boundary", "solid", "flag"],
    "formats": [(np.float32, 3), (np.float32, 3), (np.float32, 3), np.int32, np.int32, np.float32, np.float32, np.float32,
                np.float32, np.float32, np.float32, np.float32, np.float32, np.float32, np.float32, np.float32, np.uint8,
                np.uint8, np.uint8],
    "offsets": [BASE_OFFSETS[k] for k in ["pos", "vel", "acc", "index", "cellnumber", "mass", "dens", "press", "delpressz",
                                          "delpressy", "delpressx", "newdens", "newpress", "newdelpressz", "newdelpressy",
                                          "newdelpressx", "boundary", "solid", "flag"]],
    "itemsize": STRIDE,
})


def pack_base(state: dict) -> np.ndarray:
    n = state["pos"].shape[0]
    r = np.zeros(n, BASE_DTYPE)
    r["pos"], r["vel"], r["acc"] = state["pos"], state["vel"], state["acc"]
    r["index"] = state["index"]
    r["mass"] = 1.0
    r["dens"], r["press"] = state["dens"], state["press"]
    r["delpressx"], r["delpressy"], r["delpressz"] = state["delpress"].T
    r["newdens"] = state["newdens"]
    r["newpress"] = 101325.0
    r["newdelpressx"], r["newdelpressy"], r["newdelpressz"] = state["newdelpress"].T
    r["boundary"] = state["boundary"]
    return r.view(np.uint8).reshape(n, STRIDE)


def unpack_base(rec: np.ndarray) -> dict:
    r = np.ascontiguousarray(rec).reshape(-1).view(BASE_DTYPE)
    return dict(pos=r["pos"].copy(), vel=r["vel"].copy(), acc=r["acc"].copy(), dens=r["dens"].copy(), press=r["press"].copy(),
                delpress=np.stack([r["delpressx"], r["delpressy"], r["delpressz"]], 1),
                newdens=r["newdens"].copy(),
                newdelpress=np.stack([r["newdelpressx"], r["newdelpressy"], r["newdelpressz"]], 1),
                index=r["index"].copy(), cell=r["cellnumber"].copy(), boundary=r["boundary"].copy())


# ---- unidyn record (FluidGPU-unidyn.cuh:119-181; offsets SURVEY.md §8 a1) ----
UNI_OFFSETS = dict(pos=0, vel=12, vel_prev=24, acc=36, acc_prev=48, index=60, cellnumber=64, subindex=68, mass=72, dens=76, press=80,
                   delpressz=84, delpressy=88, delpressx=92, diffusion=96, newdens=108, newdelpressz=112, newdelpressy=116,
                   newdelpressx=120, boundary=316, solid=320, fluid=324, delsolid=328, delfluid=332, flag=336, split=337)
_UN = ["pos", "vel", "acc", "index", "cellnumber", "subindex", "mass", "dens", "press", "delpressz", "delpressy", "delpressx",
       "newdens", "newdelpressz", "newdelpressy", "newdelpressx", "boundary", "solid", "fluid", "flag"]
_UF = [(np.float32, 3), (np.float32, 3), (np.float32, 3), np.int32, np.int32, np.int32, np.float32, np.float32, np.float32, np.float32,
       np.float32, np.float32, np.float32, np.float32, np.float32, np.float32, np.uint8, np.float32, np.float32, np.uint8]
UNI_DTYPE = np.dtype({"names": _UN, "formats": _UF, "offsets": [UNI_OFFSETS[k] for k in _UN], "itemsize": STRIDE})


def pack_unidyn(state: dict) -> np.ndarray:
    n = state["pos"].shape[0]
    r = np.zeros(n, UNI_DTYPE)
    r["pos"], r["vel"], r["acc"] = state["pos"], state["vel"], state["acc"]
    r["index"] = state["index"]
    r["mass"] = 1.0
    r["dens"], r["press"] = state["dens"], state["press"]
    r["delpressx"], r["delpressy"], r["delpressz"] = state["delpress"].T
    r["newdens"] = state["newdens"]
    r["newdelpressx"], r["newdelpressy"], r["newdelpressz"] = state["newdelpress"].T
    r["boundary"] = state["boundary"]
    r["solid"], r["fluid"] = state["solid"], state["fluid"]
    return r.view(np.uint8).reshape(n, STRIDE)


def unpack_unidyn(rec: np.ndarray) -> dict:
    r = np.ascontiguousarray(rec).reshape(-1).view(UNI_DTYPE)
    return dict(pos=r["pos"].copy(), vel=r["vel"].copy(), acc=r["acc"].copy(), dens=r["dens"].copy(), press=r["press"].copy(),
                delpress=np.stack([r["delpressx"], r["delpressy"], r["delpressz"]], 1), newdens=r["newdens"].copy(),
                newdelpress=np.stack([r["newdelpressx"], r["newdelpressy"], r["newdelpressz"]], 1), index=r["index"].copy(),
                cell=r["cellnumber"].copy(), boundary=r["boundary"].copy(), solid=r["solid"].copy(), fluid=r["fluid"].copy(),
                subindex=r["subindex"].copy(), mass=r["mass"].copy())
