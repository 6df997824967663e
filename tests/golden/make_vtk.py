"""Writes the golden frames tests/golden/vtk_*.vtk with the reference's UNMODIFIED VisIt writer
(oracle/_ref/ref_vtk = /root/reference/visit_writer.cpp + oracle/ref_vtk.cpp; `make -C oracle ref`).
Run in the build container (the reference is not on the GPU box):  python tests/golden/make_vtk.py"""
import pathlib
import subprocess
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from fluidsolvergpu_b200 import sections  # noqa: E402

REF = ROOT / "oracle" / "_ref" / "ref_vtk"


def cases():
    rng = np.random.default_rng(42)
    n = 23        # not a multiple of 9: exercises the 9-per-line wrapping and the section breaks
    pts = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    pts[0] = (0.0, -0.0, 1.0e-30)
    pts[1] = (123456.789, -9.87654321e-12, 3.0e20)
    yield "frame_unidyn", pts, {"mass": np.ones(n, np.float32), "surface_level": rng.uniform(0, 4e4, n).astype(np.float32)}
    yield "frame_base", pts, {"dens": rng.uniform(9000, 11000, n).astype(np.float32), "cellnumber": rng.integers(0, 64000, n).astype(np.float32)}
    yield "vectors", pts[:9], {"a": rng.normal(size=9).astype(np.float32), "v": rng.normal(size=(9, 3)).astype(np.float32),
                               "b": rng.normal(size=9).astype(np.float32), "w": rng.normal(size=(9, 3)).astype(np.float32),
                               "c": rng.normal(size=9).astype(np.float32)}
    yield "single_point", pts[:1], {"x": np.array([2.5], np.float32)}
    yield "no_vars", pts[:10], {}


def main():
    out = pathlib.Path(__file__).parent
    for name, pts, variables in cases():
        for binary in (0, 1):
            tmp = out / f"_{name}.bin"
            arrays = {"pts": pts.reshape(-1)}
            arrays.update({k: v.reshape(-1) for k, v in variables.items()})
            sections.write_sections(tmp, arrays)
            subprocess.check_call([str(REF), str(tmp), str(out / f"vtk_{name}_{'bin' if binary else 'ascii'}.vtk"), str(binary)])
            tmp.unlink()
    print("golden frames written")


if __name__ == "__main__":
    main()
