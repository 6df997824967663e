"""Runs the reference's UNMODIFIED base kernels (oracle/_ref/ref_harness_base, built from
/root/reference by `make -C oracle ref`) on a B200 and turns its dumps into the compressed golden
fixtures under tests/golden/.  Run on the GPU box:

    gpurun -- python tools/make_golden.py gpurun_out/golden

then copy gpurun_out/golden/*.npz and golden_noise.json into tests/golden/.  The harness binary is
prebuilt in the build container (the GPU box has no /root/reference).
"""
import json
import pathlib
import subprocess
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from fluidsolvergpu_b200 import scenes, sections  # noqa: E402

import os
HARNESS = ROOT / "oracle" / "_ref" / os.environ.get("FSG_REF_HARNESS", "ref_harness_base_nodivsync")
KEEP = ("pos", "vel", "acc", "dens", "press", "delpress", "newdens", "newdelpress", "index", "cell", "boundary",
        "cells_sorted", "start", "end", "spts", "a3", "b3")

CASES = {
    # name: (scene factory or None for the built-in scene of solver.cu:115-121, steps to dump)
    "config1": (None, (1, 2, 10, 100)),
    "random_boundary": (lambda: scenes.random_base_scene(5000, 2, boundary_frac=0.15), (1, 2, 3)),
    "dense_overflow": (lambda: scenes.random_base_scene(6000, 7, box=((-0.2, 0.2),) * 3, spacing=0.025, jitter=0.005), (1, 3)),
}


def run(out: pathlib.Path, name: str, tag: str, scene, steps):
    prefix = out / f"{name}_{tag}"
    cmd = [str(HARNESS), "--steps", str(max(steps)), "--dump", ",".join(map(str, steps)), "--out", str(prefix)]
    if scene is not None:
        inp = out / f"{name}_in.bin"
        sections.write_sections(inp, {k: scene[k] for k in ("pos", "vel", "acc", "dens", "press", "newdens", "newdelpress", "index", "boundary")})
        cmd += ["--in", str(inp)]
    timing = subprocess.check_output(cmd, timeout=180).decode().strip().splitlines()[-1]
    dumps = {k: sections.read_sections(f"{prefix}_step{k}.bin") for k in steps}
    return dumps, json.loads(timing)


UNIDYN_HARNESS = ROOT / "oracle" / "_ref" / "ref_harness_unidyn"
UKEEP = KEEP[:-3] + ("solid", "fluid", "diffusion", "subindex", "split", "spts", "a3", "b3")
UNIDYN_CASES = {
    "config2": (lambda: scenes.unidyn_default_scene(), (1, 2, 10, 100)),
    "unidyn_random": (lambda: scenes.random_unidyn_scene(6000, 5), (1, 2, 3)),
}


def run_unidyn(out: pathlib.Path, name: str, tag: str, scene, steps):
    prefix = out / f"{name}_{tag}"
    inp = out / f"{name}_in.bin"
    sections.write_sections(inp, {k: scene[k] for k in ("pos", "vel", "acc", "dens", "press", "newdens", "index", "boundary", "solid", "fluid")})
    cmd = [str(UNIDYN_HARNESS), "--steps", str(max(steps)), "--dump", ",".join(map(str, steps)), "--out", str(prefix), "--in", str(inp)]
    timing = subprocess.check_output(cmd, timeout=300).decode().strip().splitlines()[-1]
    dumps = {k: sections.read_sections(f"{prefix}_step{k}.bin") for k in steps}
    return dumps, json.loads(timing)


def rel_l2(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    d = np.sqrt((b * b).sum())
    return float(np.sqrt(((a - b) ** 2).sum()) / d) if d > 0 else 0.0


def main():
    out = pathlib.Path(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
    out.mkdir(parents=True, exist_ok=True)
    noise, timings = {}, {}
    for name, (factory, steps) in CASES.items():
        scene = factory() if factory else None
        d1, t1 = run(out, name, "run1", scene, steps)
        d2, _ = run(out, name, "run2", scene, steps)
        timings[name] = t1
        for k in steps:
            a, b = d1[k], d2[k]
            oa, ob = np.argsort(a["index"], kind="stable"), np.argsort(b["index"], kind="stable")
            noise[f"{name}_step{k}"] = {f: rel_l2(a[f].reshape(len(oa), -1)[oa], b[f].reshape(len(ob), -1)[ob])
                                        for f in ("pos", "vel", "acc", "dens", "press", "delpress")}
            noise[f"{name}_step{k}"]["int_equal"] = bool(all(np.array_equal(a[f], b[f]) for f in ("cells_sorted", "start", "end", "index", "cell")))
            np.savez_compressed(out / f"ref_{name}_step{k}.npz", **{f: a[f] for f in KEEP})
    if UNIDYN_HARNESS.exists():
        for name, (factory, steps) in UNIDYN_CASES.items():
            scene = factory()
            d1, t1 = run_unidyn(out, name, "run1", scene, steps)
            d2, _ = run_unidyn(out, name, "run2", scene, steps)
            timings[name] = t1
            for k in steps:
                a, b = d1[k], d2[k]
                oa, ob = np.argsort(a["index"], kind="stable"), np.argsort(b["index"], kind="stable")
                noise[f"{name}_step{k}"] = {f: rel_l2(a[f].reshape(len(oa), -1)[oa], b[f].reshape(len(ob), -1)[ob])
                                            for f in ("pos", "vel", "acc", "dens", "press", "delpress", "fluid")}
                noise[f"{name}_step{k}"]["int_equal"] = bool(all(np.array_equal(a[f], b[f]) for f in ("cells_sorted", "start", "end", "split", "index", "cell")))
                np.savez_compressed(out / f"ref_{name}_step{k}.npz", **{f: a[f] for f in UKEEP})
    (out / "golden_noise.json").write_text(json.dumps({"run_to_run_rel_l2": noise, "timing": timings}, indent=1))
    for p in out.glob("*.bin"):
        p.unlink()
    print(json.dumps(noise, indent=1))
    print(json.dumps(timings))


if __name__ == "__main__":
    main()
