"""The reference's own GPU kernels next to libfsg on the two scenes the reference can run (configs[0] and
configs[1] of BASELINE.json): device milliseconds per step over 100 steps on the same B200.

The reference side is oracle/_ref/ref_harness_* (the reference's kernel objects built from /root/reference,
driven by our harness; test infrastructure).  The reference cannot run the 256^3..1024^3 scenes: its
`int idx = blockIdx.x*blockDim.x + threadIdx.x` launch shape overflows beyond 128^3 bins and the base kernel
drops neighbours beyond 64 per bin neighbourhood.

    gpurun -- python tools/ref_gpu_compare.py > gpurun_out/ref_gpu_compare.json
"""
import json
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import fluidsolvergpu_b200 as fsg  # noqa: E402
from fluidsolvergpu_b200 import scenes, sections  # noqa: E402

REF = ROOT / "oracle" / "_ref"
STEPS = 100


def ours(cfg, state, steps=STEPS):
    with fsg.FluidSolver(cfg) as s:
        s.upload(state)
        s.step(5)
        stream = torch.cuda.ExternalStream(s.stream())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        s.step(steps, sync=False)
        e1.record(stream)
        s.sync()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps


def ref(binary, scene=None, keys=(), steps=STEPS, timeout=300):
    cmd = [str(REF / binary), "--steps", str(steps), "--out", "/tmp/ref_cmp"]
    if scene is not None:
        inp = "/tmp/ref_cmp_in.bin"
        sections.write_sections(inp, {k: scene[k] for k in keys})
        cmd += ["--in", inp]
    if not (REF / binary).exists():
        return None
    try:
        out = subprocess.check_output(cmd, timeout=timeout, stderr=subprocess.DEVNULL).decode().strip().splitlines()[-1]
        return json.loads(out)
    except Exception as e:      # a hang or a crash of the reference at this scale is a result too
        return {"error": repr(e)[:200]}


def main():
    res = {}
    s1 = scenes.base_default_scene()
    r1 = ref("ref_harness_base_nodivsync")
    o1 = ours(fsg.FluidSolver.base_config(), s1)
    res["config1_base_default_scene"] = {"particles": 8000, "steps": STEPS, "libfsg_ms_per_step": o1,
                                         "reference_gpu_ms_per_step": r1 and r1["ms_per_step"], "reference_detail": r1,
                                         "speedup": r1 and r1["ms_per_step"] / o1,
                                         "note": "reference = its kernels with the divergent __syncthreads of FluidGPU.cu:280 removed (the unmodified kernel hangs on B200)"}
    s2 = scenes.unidyn_default_scene()
    r2 = ref("ref_harness_unidyn", s2, ("pos", "vel", "acc", "dens", "press", "newdens", "index", "boundary", "solid", "fluid"))
    o2 = ours(fsg.FluidSolver.unidyn_config(), s2)
    res["config2_unidyn_default_scene"] = {"particles": 14040, "steps": STEPS, "libfsg_ms_per_step": o2,
                                           "reference_gpu_ms_per_step": r2 and r2["ms_per_step"], "speedup": r2 and r2["ms_per_step"] / o2,
                                           "note": "reference = unmodified unidyn kernels; both sides are launch-latency bound at this size"}
    # The plume scene at 128^3 bins (1.07 M particles): the largest grid the reference's launch shapes can address.  Reference side:
    # its unidyn kernels (the base kernel drops neighbours beyond 64 per neighbourhood) rebuilt with GRIDSIZE = 128 and the domain
    # of the plume, unit-box walls moved away (oracle/Makefile, _ref/ref_harness_unidyn_g128).  Different update physics than the base
    # step libfsg runs on this scene (leapfrog instead of Euler, same pair sums): a THROUGHPUT comparison, not a parity case.
    cfg = scenes.plume_config(128)
    s3 = scenes.plume_scene(cfg)
    n3 = s3["pos"].shape[0]
    cfg.capacity = n3
    s3["solid"] = __import__("numpy").zeros(n3, "float32")
    s3["fluid"] = __import__("numpy").ones(n3, "float32")
    r3 = ref("ref_harness_unidyn_g128", s3, ("pos", "vel", "acc", "dens", "press", "newdens", "index", "boundary", "solid", "fluid"), steps=10,
             timeout=600)
    o3 = ours(cfg, s3, 20)
    ok = bool(r3) and "ms_per_step" in r3
    res["plume_128"] = {"particles": n3, "bins": 128 ** 3, "libfsg_ms_per_step": o3, "libfsg_particle_steps_per_s": n3 / o3 * 1e3,
                        "reference_gpu_ms_per_step": r3["ms_per_step"] if ok else None, "reference_detail": r3,
                        "reference_particle_steps_per_s": n3 / r3["ms_per_step"] * 1e3 if ok else None,
                        "speedup": r3["ms_per_step"] / o3 if ok else None,
                        "note": "reference = its unidyn kernels rebuilt for a 128^3 grid (build-time constants only); beyond 128^3 its launch "
                                "index overflows int, so this is the largest comparison point there is"}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
