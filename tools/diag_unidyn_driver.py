"""Diagnostic: the real solver-unidyn.cu driver on a 2-GPU box, reference kernels vs libfsg, each run twice: how far apart are the frames?"""
import sys, pathlib, tempfile
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT))
from test_parity_gpu import _read_vtk_ascii, run_real_driver
from scipy.spatial import cKDTree
tmp = pathlib.Path(tempfile.mkdtemp())
runs = {}
for name, exe in (("ref1", "solver_unidyn_ref"), ("ref2", "solver_unidyn_ref"), ("fsg1", "solver_unidyn_compat"), ("fsg2", "solver_unidyn_compat")):
    (tmp / name / "anim-uni").mkdir(parents=True)
    out = run_real_driver(ROOT / "oracle" / "_ref" / exe, tmp / name)
    runs[name] = {t: _read_vtk_ascii(tmp / name / "anim-uni" / f"anim_s_GPU0_{t}.vtk") for t in (20, 40, 60, 80)}
    print(name, "lines", len(out.splitlines()), [runs[name][t][0].shape for t in (20, 80)], flush=True)
for a, b in (("ref1", "ref2"), ("fsg1", "fsg2"), ("fsg1", "ref1")):
    for t in (20, 40, 60, 80):
        pa, pb = runs[a][t][0], runs[b][t][0]
        if pa.shape != pb.shape:
            print(a, b, t, "shapes", pa.shape, pb.shape); continue
        dist, nn = cKDTree(pb).query(pa)
        print(a, b, t, "unique", len(np.unique(nn)), "of", len(nn), "rel", float(np.sqrt((dist ** 2).sum() / (pb ** 2).sum())), "max", float(dist.max()),
              "n>1e-3", int((dist > 1e-3).sum()), flush=True)
