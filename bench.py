#!/usr/bin/env python
"""bench.py — throughput of the per-timestep particle update (the loop body of solver.cu:181-198) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--grid G] [--impl fsg|reference]

A "step" is one pass of the hot path (key sort -> reorder + bin tables -> pair sums + EOS/integrate/re-bin)
over the whole synthetic plume scene of SURVEY.md §8d: a column of fluid particles in a G^3 bin grid
(BASELINE.json configs[3], 512^3; configs[2] and [4] with --grid 256 / 1024).

  value     cell-updates/s = G^3 bins x steps / s, state resident in HBM (device-generated scene),
            timed with CUDA events on the solver's stream, max over ranks.
  e2e       the same metric through the C ABI with HOST buffers: every step uploads the full particle
            state from pinned host memory (fsg_upload_soa), steps once, and downloads it again
            (fsg_download_soa); copies are inside the timed region.
  roofline  the dominant kernel (pair sums + update), timed live with CUDA events on the solver's
            stream inside the timed region (fsg_set_profiling), as algorithmic bytes / time vs the
            measured HBM peak, plus the FP32 issue-rate view that actually bounds it.
  cpu_baseline  the CPU oracle (oracle/, test infrastructure) on a bounded sample of the same scene
            family on this box's host cores — a reported baseline, not a target.

--impl reference times the CPU restatement of the reference path (the reference ships CUDA kernels
only; its "CPU path" is the host restatement oracle/fsg_oracle.c) with every host thread.
"""
from __future__ import annotations

import argparse
import json
import os
import pathlib
import subprocess
import sys
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

ALG_BYTES_STEP = 272      # algorithmic HBM bytes per particle-step (reorder 136 + tables 4 + pair/update 132)
ALG_BYTES_PAIR = 132      # pair sums + EOS/integrate/re-bin as ONE phase (survey's figure): read 64 + write 64 + new key 4
# With the deferred update (single device and the sorted-ghost slab pipeline) the pair phase is the pair sums alone — it reads the
# read state (posd, velp: 32 B) and writes the four sums (16 B) — and Particle::update runs inside the next step's reorder:
ALG_BYTES_PAIR_DEFERRED = 48
ALG_BYTES_STREAM_DEFERRED = 188   # key sort (8 + 8 + 16) + reorder with update (8 keys/perm, 64 + 16 in, 64 + 4 out)
# measured DRAM bytes per particle of the pair kernel (ncu --set full at 256^3): [symmetric k_pair_v3, gather k_pair_v2]
NCU_PAIR_DRAM_BYTES_PER_PARTICLE = [(65.2, "profiles/r2_ncu_pair_v3_512.txt"), (47.2, "profiles/r1_ncu_pair_v2_final.txt")]
FLOP_IN_RANGE, FLOP_REJECTED = 50, 12   # SURVEY.md §8d algorithmic flop per in-range / rejected candidate
SPACING, JITTER, SEED = 0.05, 0.005, 20261018
CPU_SAMPLE_GRID = 128     # bounded sample for the CPU legs: the same plume at 128^3 bins (1.07 M particles)
# the arithmetic type of the path, stated in full: production pair sums are pure fp32 (+ rsqrt.approx / div.approx) where the
# reference promotes some sub-expressions to double; the update follows the reference's promotions exactly; pair_fp64 = 1 is the faithful path
DTYPE = "f32 (reference: f32 with f64 sub-expressions)"
DTYPE_NOTE = ("pair sums: pure fp32 with rsqrt.approx / div.approx where the reference promotes sub-expressions to double (<= 1e-5 per step, "
              "tested; fsg_config.pair_fp64 = 1 is the promotion-faithful path); EOS / integration / bin id follow the reference's promotions exactly")


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.proc, self.path = device, None, f"/tmp/fsg_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.proc.wait()
        self.f.close()
        sm, mx, pw, reasons = [], [], [], set()
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU legs (oracle = test infrastructure; this is one of the two places allowed to execute it)
# ------------------------------------------------------------------------------------------------
def cpu_run(grid: int, steps: int, warmup: int, threads: int = 0):
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_py
    import fluidsolvergpu_b200 as fsg
    cfg = fsg.scenes.plume_config(grid)
    state = fsg.scenes.plume_scene(cfg, SPACING, JITTER, SEED)
    n = state["pos"].shape[0]
    cores = threads or (os.cpu_count() or 1)
    sim = oracle_py.OracleSim(oracle_py.params_from_cfg(cfg, threads=cores), state)
    sim.step(warmup)
    t0 = time.perf_counter()
    sim.step(steps)
    dt = time.perf_counter() - t0
    return dict(n=n, grid=grid, cores=cores, seconds=dt, ms_per_step=dt / steps * 1e3,
                cell_updates_per_s=grid ** 3 * steps / dt, particle_steps_per_s=n * steps / dt)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_run(CPU_SAMPLE_GRID, args.steps, args.warmup)
    sample = f"plume {CPU_SAMPLE_GRID}^3 bins, {r['n']} particles per step, {args.steps} steps (bounded sample of the {args.grid}^3 scene family; rate per bin is size-independent)"
    line = {
        "impl": "reference", "metric": "cell-updates/s", "value": r["cell_updates_per_s"], "unit": "cell-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": dict(workload_config(CPU_SAMPLE_GRID, r["n"], 1), decomposition=f"host CPU, {r['cores']} threads (OpenMP over bins)",
                       sample_of=f"{args.grid}^3 plume (the fsg arm's workload); this arm times the same scene family at {CPU_SAMPLE_GRID}^3 bins "
                                 f"(same {r['n'] / CPU_SAMPLE_GRID ** 3:.3f} particles per bin), bounded so the run ends within minutes"),
        "particle_steps_per_s": r["particle_steps_per_s"],
        "cpu_baseline": {"value": r["cell_updates_per_s"], "unit": "cell-updates/s", "cores": r["cores"], "kind": "port",
                         "sample": sample},
        "e2e": {"value": r["cell_updates_per_s"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference ships CUDA kernels only; this arm is the host restatement of those kernels (oracle/fsg_oracle.c, OpenMP over bins)",
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# Parity evidence AT the benchmarked configuration (outside the timed region; oracle = the checker)
# ------------------------------------------------------------------------------------------------
def parity_sample(solver, cfg, nbins=256, seed=1234):
    """One more step of the benchmarked solver, checked against the CPU oracle on a random sample of home bins:
    the particles of `nbins` occupied bins + everything in their 27-bin neighbourhoods (linear offsets, FluidGPU.cu:124-126)
    form a sub-scene on the SAME grid; the oracle steps it once; the home-bin particles (whose neighbourhoods are complete)
    are compared by Particle::index with what the device produced — new bin ids bit-exact, fields <= 1e-5 relative L2."""
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_py
    G = cfg.grid
    before = solver.download()
    solver.step(1)
    after = solver.download()
    cell0 = before["cell"]
    rng = np.random.default_rng(seed)
    live = np.flatnonzero(cell0 < G ** 3)
    # the bins of randomly drawn particles: occupied by construction (no 68 M-element sort to list the occupied bins)
    home = np.unique(cell0[live[rng.integers(0, live.size, size=4 * nbins)]])
    home = np.sort(rng.permutation(home)[:nbins])
    off = np.array([a * G * G + b * G + c for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1)], np.int64)
    nb = np.unique(np.clip(home[:, None].astype(np.int64) + off[None, :], 0, G ** 3 - 1))
    mask = np.isin(cell0, nb)
    sub = {k: np.ascontiguousarray(v[mask]) for k, v in before.items() if k != "cell"}
    sim = oracle_py.OracleSim(oracle_py.params_from_cfg(cfg), sub).step(1)
    ref = sim.state()
    is_home = np.isin(cell0[mask], home)                 # rows of `sub` (upload order) that are home-bin particles
    want_idx = sub["index"][is_home]
    ro = np.argsort(ref["index"], kind="stable")
    ref = {k: v[ro] for k, v in ref.items()}
    pick_r = np.searchsorted(ref["index"], want_idx)
    aidx = after["index"]
    if aidx.size and int(aidx.min()) >= 0 and int(aidx.max()) < aidx.size:       # Particle::index is a permutation of 0..n-1: O(n) inverse
        inv = np.empty(aidx.size, np.int64)
        inv[aidx] = np.arange(aidx.size)
        pick_g = inv[want_idx]
    else:
        go = np.argsort(aidx, kind="stable")
        pick_g = go[np.searchsorted(aidx[go], want_idx)]
    assert np.array_equal(ref["index"][pick_r], want_idx) and np.array_equal(after["index"][pick_g], want_idx)
    ints_equal = bool(np.array_equal(ref["cell"][pick_r], after["cell"][pick_g]))
    errs = {f: oracle_py.rel_l2(after[f][pick_g], ref[f][pick_r]) for f in ("pos", "vel", "acc", "dens", "press", "delpress")}
    pos_bits = bool(np.array_equal(after["pos"][pick_g].view(np.uint32), ref["pos"][pick_r].view(np.uint32)))
    rec = {"grid": G, "bins": int(home.size), "particles_compared": int(want_idx.size), "sub_scene_particles": int(mask.sum()),
           "ints_equal": ints_equal, "positions_bit_equal": pos_bits, "max_rel_l2": max(errs.values()), "rel_l2": errs,
           "tolerance": 1e-5, "ok": bool(ints_equal and max(errs.values()) <= 1e-5),
           "what": "one device step vs one CPU-oracle step from the same bits, home-bin particles of a random bin sample, compared by index"}
    return rec


def slab_parity(fsg, dist, rank, world, local, exchange, steps=4, grid=128):
    """N > 1: the 128^3 plume (with a drift along x so that particles cross the faces) on the N slab PROCESSES through the real
    transport (CUDA-IPC peer copies + device-side stamps, or NCCL send/recv) against a single context on rank 0, compared by
    Particle::index, every step from identical bits: positions / velocities / bin ids bit for bit, pair sums <= 1e-5.  Run with the
    default symmetric pair kernel (what the timed steps run) and with the deterministic gather kernel."""
    out = {}
    for mode, name in ((0, "symmetric_kernel"), (1, "gather_kernel")):
        res = fsg.slab.parity_against_single(grid, rank, world, local, exchange=exchange, steps=steps, pair_mode=mode,
                                             spacing=SPACING, jitter=JITTER, seed=SEED)
        if rank == 0:
            out[name] = res
    if rank != 0:
        return None
    out["ok"] = bool(all(out[k]["bit_exact"] and out[k]["conserved"] and out[k]["max_rel_l2"] <= 1e-5 and out[k]["migrated"] > 0 and
                         out[k].get("free_running", {}).get("conserved", False) and out[k]["free_running"]["max_rel_l2"] <= 2e-4
                         for k in ("symmetric_kernel", "gather_kernel")))
    out["what"] = (f"plume {grid}^3 + x-drift, {steps} steps: {world} slab processes over the '{exchange}' transport vs one context, by index, "
                   "each step from identical bits; bit_exact = pos / vel / cell / boundary, max_rel_l2 = acc / dens / press / delpress; "
                   "free_running = 6 more steps without any download in between (deferred update, migrants with pending sums) against the "
                   "same steps on one context: particle set conserved, every field within 2e-4 (the two differ by the order of the pair sums)")
    return out


# ------------------------------------------------------------------------------------------------
# The reference's OWN CUDA kernels next to libfsg on the scenes the reference can run (N = 1, outside the timed region).
# oracle/_ref/* are the reference's kernel objects built from /root/reference by oracle/Makefile (test infrastructure).
# ------------------------------------------------------------------------------------------------
def reference_gpu_leg(fsg, torch, steps=100):
    from fluidsolvergpu_b200 import scenes, sections
    REF = ROOT / "oracle" / "_ref"

    def ours(cfg, state, k):
        with fsg.FluidSolver(cfg) as s:
            s.upload(state)
            s.step(5)
            stream = torch.cuda.ExternalStream(s.stream())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            s.step(k, sync=False)
            e1.record(stream)
            s.sync()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / k

    def ref(binary, scene, keys, k, timeout):
        if not (REF / binary).exists():
            return {"error": f"oracle/_ref/{binary} not built (needs /root/reference at build time)"}
        tag = f"/tmp/fsg_refgpu_{os.getpid()}"
        cmd = [str(REF / binary), "--steps", str(k), "--out", tag]
        if scene is not None:
            sections.write_sections(tag + "_in.bin", {q: scene[q] for q in keys})
            cmd += ["--in", tag + "_in.bin"]
        try:
            out = subprocess.check_output(cmd, timeout=timeout, stderr=subprocess.DEVNULL).decode().strip().splitlines()[-1]
            return json.loads(out)
        except Exception as e:          # a hang or a crash of the reference at this scale is a result too
            return {"error": repr(e)[:200]}
        finally:
            for p in pathlib.Path("/tmp").glob(pathlib.Path(tag).name + "*"):
                try:
                    p.unlink()
                except OSError:
                    pass

    ukeys = ("pos", "vel", "acc", "dens", "press", "newdens", "index", "boundary", "solid", "fluid")
    res = {}
    r = ref("ref_harness_base_nodivsync", None, (), steps, 300)
    o = ours(fsg.FluidSolver.base_config(), scenes.base_default_scene(), steps)
    res["config1"] = {"scene": "solver.cu default scene, 8000 particles, 40^3 bins", "steps": steps, "ref_ms": r.get("ms_per_step"), "fsg_ms": o,
                      "ref_detail": r, "note": "reference = its own kernels, __syncthreads of FluidGPU.cu:280 removed at build time (the unmodified kernel hangs on B200)"}
    s2 = scenes.unidyn_default_scene()
    r = ref("ref_harness_unidyn", s2, ukeys, steps, 300)
    o = ours(fsg.FluidSolver.unidyn_config(), s2, steps)
    res["config2"] = {"scene": "solver-unidyn.cu default scene, 14040 particles, 17^3 bins", "steps": steps, "ref_ms": r.get("ms_per_step"), "fsg_ms": o,
                      "ref_detail": r, "note": "reference = its unmodified unidyn kernels"}
    cfg = scenes.plume_config(128)
    s3 = scenes.plume_scene(cfg, SPACING, JITTER, SEED)
    n3 = s3["pos"].shape[0]
    cfg.capacity = n3
    s3["solid"], s3["fluid"] = np.zeros(n3, np.float32), np.ones(n3, np.float32)
    r = ref("ref_harness_unidyn_g128", s3, ukeys, 10, 600)
    o = ours(cfg, s3, 20)
    # the SAME model on both sides: libfsg's unidyn path on that scene (unit-box floor / walls off, as in the rebuilt reference)
    ucfg = fsg.FluidSolver.unidyn_config(capacity=n3, grid=128, origin=cfg.origin, unidyn_open_box=1)
    ou = ours(ucfg, s3, 10)
    # ... and its mixed-phase / granular path (two-pass kernels): the same particles stratified into sand over a mixture band over water
    z = s3["pos"][:, 2]
    zq = np.quantile(z, [0.45, 0.55])
    rng = np.random.default_rng(SEED)
    s3m = dict(s3)
    s3m["solid"] = np.where(z > zq[1], 1.0, np.where(z > zq[0], rng.uniform(0.05, 0.95, n3), 0.0)).astype(np.float32)
    s3m["fluid"] = (1.0 - s3m["solid"]).astype(np.float32)
    try:
        om = ours(ucfg, s3m, 5)
    except Exception as e:                      # noqa: BLE001
        om = None
    res["plume128"] = {"scene": f"synthetic plume 128^3 bins, {n3} particles (the largest grid the reference's launch shapes address)", "steps": 10,
                       "ref_ms": r.get("ms_per_step"), "fsg_ms": o, "fsg_unidyn_ms": ou, "fsg_unidyn_mixed_phase_ms": om, "ref_detail": r,
                       "speedup_same_model": (r["ms_per_step"] / ou) if r.get("ms_per_step") else None,
                       "note": "reference = its unidyn kernels rebuilt with build-time constants for a 128^3 grid (oracle/Makefile); throughput comparison, "
                               "different update physics (leapfrog vs Euler), same pair sums"}
    for v in res.values():
        v["speedup"] = (v["ref_ms"] / v["fsg_ms"]) if v["ref_ms"] else None
    return res


def frame_output_leg(fsg, torch, grid=128, steps=60, every=20):
    """Frame output beside the step loop: the plume at grid^3, `steps` steps with a legacy-VTK ASCII frame
    (positions + 2 scalars, "%20.12e " per number: the reference's format, solver-unidyn.cu:472-493) every `every` steps — none,
    synchronous (fsg_write_frame, the reference's way) and asynchronous (fsg_write_frame_async: export kernel + copy stream +
    writer thread).  Wall clock per step including the final wait for the files."""
    import shutil
    import tempfile
    cfg = fsg.scenes.plume_config(grid)
    cfg.capacity = fsg.scenes.plume_count(cfg, SPACING)
    out = {"grid": grid, "particles": int(cfg.capacity), "steps": steps, "frame_every": every}
    tmp = tempfile.mkdtemp(prefix="fsg_frames_")
    try:
        for mode in ("none", "sync", "async"):
            with fsg.FluidSolver(cfg) as s:
                s.scene_plume(SPACING, JITTER, SEED)
                s.step(3)
                t0 = time.perf_counter()
                loop_s = 0.0
                for k in range(1, steps + 1):
                    s.step(1, sync=False)
                    if mode != "none" and k % every == 0:
                        name = f"{tmp}/{mode}_{k}.vtk"
                        (s.write_frame if mode == "sync" else s.write_frame_async)(name)
                s.sync()
                loop_s = time.perf_counter() - t0
                if mode == "async":
                    s.frame_wait()
                total_s = time.perf_counter() - t0
            out[mode] = {"ms_per_step_loop": loop_s / steps * 1e3, "ms_per_step_incl_final_wait": total_s / steps * 1e3}
        f0 = next(pathlib.Path(tmp).glob("sync_*.vtk"))
        out["frame_bytes"] = f0.stat().st_size
        out["files_identical"] = all((pathlib.Path(tmp) / f"sync_{k}.vtk").stat().st_size == (pathlib.Path(tmp) / f"async_{k}.vtk").stat().st_size
                                     for k in range(every, steps + 1, every))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


def workload_config(grid, n, gpus):
    return {"workload": f"synthetic plume, {grid}^3 bins (CELLSIZE 0.12 = 2h, lattice spacing {SPACING}, jitter {JITTER}, seed {SEED}), "
                        f"base SPH step: key sort + reorder/bin tables + pair sums + EOS/integrate/re-bin",
            "grid": grid, "particles": n, "decomposition": "single device" if gpus == 1 else f"{gpus} x-slabs, ghost exchange per step",
            "l2_policy": "working set (>= 64 B x particles x 2 buffers) exceeds the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def make_solver(fsg, G, rank, world, local, args):
    """One context holding the whole G^3 plume (world == 1) or this rank's x-slab of it."""
    cfg = fsg.scenes.plume_config(G)
    cfg.pair_mode = args.pair_mode
    cfg.device = local
    cfg.rank, cfg.world = rank, world
    n_total = fsg.scenes.plume_count(cfg, SPACING)
    if world == 1:
        cfg.capacity = n_total
        return fsg.FluidSolver(cfg), cfg, n_total
    # x-slabs of equal particle count; capacity = owned + two ghost layers + slack
    hist = fsg.slab.plume_layer_hist(cfg, SPACING)
    cuts = fsg.slab_cuts(hist, world)
    owned = [int(hist[a:b].sum()) for a, b in cuts]
    cap_m, cap_g = fsg.slab.message_caps(hist, cuts)
    if args.exchange == "peer" and not (args.overlap or args.classic_slabs):
        cap = int(max(owned) * 1.03) + 2 * cap_g + 65536          # sorted ghosts: own particles + slack | the two ghost zones
    else:
        cap = int(max(owned) * 1.05) + 3 * int(hist.max()) + 65536  # classic: ghosts are appended behind the own particles
    cfg = fsg.slab_config(cfg, rank, world, cuts, cap, local)
    solver = fsg.SlabSolver(cfg, fsg.DistExchange(), cap_m, cap_g)
    if args.exchange == "peer":
        solver.setup_peer_exchange(overlap=args.overlap, classic=args.classic_slabs)
    return solver, cfg, n_total


def short_leg(fsg, torch, dist, G, rank, world, local, args, steps=5, warmup=3):
    """A short resident-state run at another grid (the 1024^3 leg of the N = 1 ... 8 lines): ms per step, max over ranks."""
    try:
        solver, cfg, n_total = make_solver(fsg, G, rank, world, local, args)
    except Exception as e:            # e.g. not enough memory on a shared device: say so instead of failing the bench line
        return {"grid": G, "error": repr(e)[:200]}
    try:
        solver.scene_plume(SPACING, JITTER, SEED)
        stream = torch.cuda.ExternalStream(solver.stream())
        solver.step(warmup)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        solver.step(steps, sync=False)
        e1.record(stream)
        solver.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        owned = n_total
        if world > 1:
            solver.check()
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
            own = torch.tensor([float(solver.owned_count())], device="cuda", dtype=torch.float64)
            dist.all_reduce(own)
            owned = int(own[0])
        ms /= steps
        out = {"grid": G, "particles": n_total, "particles_conserved": owned == n_total, "steps": steps, "warmup": warmup, "ms_per_step": ms,
               "value": G ** 3 / (ms * 1e-3), "unit": "cell-updates/s", "particle_steps_per_s": n_total / (ms * 1e-3)}
        # the oracle-sampled parity check at THIS grid too (N = 1; two whole-state downloads: ~70 GB of host memory at 1024^3)
        if world == 1 and not args.no_parity:
            try:
                import psutil
                room = psutil.virtual_memory().available >= 160 * n_total
            except Exception:
                room = False
            out["parity_check"] = parity_sample(solver, cfg, args.parity_bins) if room else {"skipped": "not enough host memory for two whole-state downloads"}
        return out
    finally:
        solver.close()


def fsg_arm(args):
    import torch
    import torch.distributed as dist
    import fluidsolvergpu_b200 as fsg

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libfsg has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    hbm_peak, peak_src, sm_max = peaks()
    G = args.grid
    solver, cfg, n_total = make_solver(fsg, G, rank, world, local, args)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    solver.scene_plume(SPACING, JITTER, SEED)
    n_local = solver.owned_count() if world > 1 else n_total
    stream = torch.cuda.ExternalStream(solver.stream())
    solver.step(args.warmup)
    barrier()
    l0 = solver.stats()["kernel_launches"]
    solver.set_profiling(True)
    if world > 1:
        solver.time_exchange = True
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    solver.step(args.steps, sync=False)
    e1.record(stream)
    solver.sync()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    phase = solver.phase_ms()
    solver.set_profiling(False)
    launches = solver.stats()["kernel_launches"] - l0
    halo = None
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t[0])
        info = solver.check()          # raises if a message / the capacity overflowed or a particle left the ghost band
        ph = [phase[k] / max(1, phase["steps"]) for k in ("sort", "reorder", "pair_update", "other")]
        payload = 64.0 * (info["sent"][0] + info["sent"][2]) + 32.0 * (info["sent"][1] + info["sent"][3])
        xm = solver.exchange_ms()
        gms = solver.ghost_ms() if solver.mode == 2 else 0.0
        mine = torch.tensor(ph + [float(n_local), payload, float(solver.wire_bytes_per_step), xm["pack"], xm["exchange"], xm["unpack"], gms],
                            device="cuda", dtype=torch.float64)
        allr = torch.empty(world * mine.numel(), device="cuda", dtype=torch.float64)
        dist.all_gather_into_tensor(allr, mine)
        allr = allr.cpu().view(world, -1)
        per_rank = {"ms_sort": allr[:, 0].tolist(), "ms_reorder": allr[:, 1].tolist(), "ms_pair_update": allr[:, 2].tolist(),
                    "ms_pack": allr[:, 7].tolist(), "ms_exchange_incl_wait": allr[:, 8].tolist(), "ms_unpack": allr[:, 9].tolist(),
                    "ms_ghost_exchange_in_reorder": allr[:, 10].tolist(), "particles": [int(v) for v in allr[:, 4].tolist()]}
        t2 = torch.tensor([float(launches), float(solver.owned_count())], device="cuda", dtype=torch.float64)
        dist.all_reduce(t2, op=dist.ReduceOp.SUM)
        launches = int(t2[0])
        if int(t2[1]) != n_total:
            raise SystemExit(f"particle count not conserved across slabs: {int(t2[1])} != {n_total}")
        wire, payload_all = float(allr[:, 6].sum()), float(allr[:, 5].sum())
        halo = {"wire_bytes_per_step_all_ranks": wire, "payload_bytes_last_step_all_ranks": payload_all,
                "wire_GBps_all_ranks": wire / (ms_total / args.steps * 1e-3) / 1e9, "nvlink_peak_GBps_per_direction": 770.0,
                "exchange": args.exchange, "overlap": args.exchange == "peer" and args.overlap,
                "pipeline": "sorted ghosts (fsg_slab2.cu)" if solver.mode == 2 else "classic (ghosts appended and sorted)",
                "note": "no host synchronisation inside a step, every count is read on the device.  sorted ghosts: migrants (pre-update "
                        "state + pending pair sums) are copied into the neighbour's inbox before the key sort (CUDA IPC mapping, copy engines, "
                        "sequence stamp last, awaited on the device); after the reorder the face layers of the SORTED state are written "
                        "straight into the neighbours' ghost zones by a kernel (remote stores over NVLink) — ghosts never pass through the "
                        "sort; per_rank.ms_ghost_exchange_in_reorder (send + wait + install) is part of ms_reorder.  classic: one fixed-size "
                        "message per neighbour and direction with migrants and ghosts, appended and sorted; overlap = boundary bins first, "
                        "next step's pack + copies on a second stream beside the interior bins; nccl = NCCL send/recv"}
    ms_step = ms_total / args.steps
    value = G ** 3 / (ms_step * 1e-3)

    # pair statistics of one extra step (outside the timed region) for the FP32 view of the roofline
    st = None
    if world == 1:
        st = solver.pair_stats_one_step()

    # ---- parity at THIS configuration: one more step against the CPU oracle on a random sample of home bins ----
    parity = None
    if world == 1 and not args.no_parity and G <= 512:
        parity = parity_sample(solver, cfg, args.parity_bins)

    # ---- roofline of the dominant kernel ----
    pair_ms = phase["pair_update"] / max(1, phase["steps"])
    # the update is deferred into the next reorder on a single device and on the sorted-ghost slab pipeline: the pair phase is the pair sums alone
    deferred = os.environ.get("FSG_DEFER_UPDATE", "1") != "0" and (world == 1 or solver.mode == 2)
    alg_pair = ALG_BYTES_PAIR_DEFERRED if deferred else ALG_BYTES_PAIR
    achieved = alg_pair * n_local / (pair_ms * 1e-3) / 1e9
    symmetric = args.pair_mode == 0 and not (world > 1 and args.exchange == "peer" and args.overlap)
    kname = "k_pair_v3 (symmetric pair sums)" if symmetric else "k_pair_v2 (gather pair sums)"
    ncu_b = NCU_PAIR_DRAM_BYTES_PER_PARTICLE[0 if symmetric else 1]
    roofline = {"bound": "hbm", "kernel": kname + ("; Particle::update is deferred into the next step's k_reorder" if deferred else
                                                   " + k_update (EOS/integrate/re-bin)"), "achieved": achieved,
                "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": ncu_b[0] * n_local + (0.0 if deferred else 152.0 * n_local), "peak_source": peak_src,
                "traffic_note": f"pair kernel: dram__bytes_read+write = {ncu_b[0]} B/particle under ncu --set full ({ncu_b[1]}), "
                                "scaled by the particle count" + ("" if deferred else "; k_update: its 152 B/particle of streaming reads + writes"),
                "kernel_ms": pair_ms, "share_of_step": pair_ms / ms_step,
                "algorithmic_bytes_per_launch": alg_pair * n_local,
                "algorithmic_bytes_per_particle": alg_pair,
                "note": "this kernel is CUDA-core (FP32 issue) bound, not HBM bound: see roofline_fp32; the HBM-bound phases are in `phases`"}
    extra = {}
    if st is not None:
        flop = FLOP_IN_RANGE * st["pairs_in_range"] + FLOP_REJECTED * (st["pairs_tested"] - st["pairs_in_range"])
        mhz = (clocks or {}).get("sm_mhz") or sm_max
        peak_tf = 148 * 128 * 2 * mhz * 1e6 / 1e12
        extra["roofline_fp32"] = {"bound": "fp32 cuda cores", "achieved": flop / (pair_ms * 1e-3) / 1e12, "peak": peak_tf,
                                  "unit": "TFLOP/s", "frac": flop / (pair_ms * 1e-3) / 1e12 / peak_tf,
                                  "peak_source": f"148 SM x 128 lanes x 2 flop x {mhz:.0f} MHz (SM clock sampled during the run)",
                                  "pairs_tested": st["pairs_tested"], "pairs_in_range": st["pairs_in_range"],
                                  "pair_tests_per_s": st["pairs_tested"] / (pair_ms * 1e-3)}
    phases = {k: phase[k] / max(1, phase["steps"]) for k in ("sort", "reorder", "pair_update", "other")}
    stream_ms = phases["sort"] + phases["reorder"]
    extra["phases_ms"] = phases
    alg_stream = ALG_BYTES_STREAM_DEFERRED if deferred else ALG_BYTES_STEP - ALG_BYTES_PAIR
    extra["streaming_phases"] = {"what": "key sort + reorder/bin tables" + (" + the previous step's Particle::update" if deferred else "") + " (HBM-bound)",
                                 "algorithmic_bytes_per_particle": alg_stream,
                                 "algorithmic_GBps": alg_stream * n_local / (stream_ms * 1e-3) / 1e9,
                                 "frac_of_hbm_peak": alg_stream * n_local / (stream_ms * 1e-3) / 1e9 / hbm_peak}
    extra["step_hbm"] = {"algorithmic_bytes_per_particle": ALG_BYTES_STEP, "algorithmic_GBps": ALG_BYTES_STEP * n_local / (ms_step * 1e-3) / 1e9,
                         "frac_of_hbm_peak": ALG_BYTES_STEP * n_local / (ms_step * 1e-3) / 1e9 / hbm_peak}

    # ---- end to end through the C ABI with host buffers ----
    args.n_total = n_total
    if world == 1 and args.e2e_contexts > 1:
        solver.close()              # its memory goes to the pipeline's contexts
        solver = None
        e2e = e2e_pipelined(fsg, torch, cfg, n_total, args, G)
    else:
        e2e = e2e_leg(solver, torch, stream, args, G, world, barrier, dist)
        solver.close()
        solver = None

    # ---- the slab path against one context, through the real multi-process transport (N > 1) ----
    slabp = None
    if world > 1 and not args.no_parity:
        slabp = slab_parity(fsg, dist, rank, world, local, args.exchange)

    # ---- the 1024^3 configuration (BASELINE configs[4]) in the same line: a short resident-state run ----
    g1024 = None
    if args.grid1024_steps > 0 and G != 1024:
        g1024 = short_leg(fsg, torch, dist, 1024, rank, world, local, args, steps=args.grid1024_steps)

    # ---- the reference's own CUDA kernels beside libfsg (N = 1) ----
    refgpu = None
    if rank == 0 and world == 1 and not args.no_reference_gpu:
        refgpu = reference_gpu_leg(fsg, torch)

    # ---- frame output beside the step loop (N = 1) ----
    frames = None
    if rank == 0 and world == 1 and not args.no_frames:
        frames = frame_output_leg(fsg, torch)

    # ---- CPU baseline (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = cpu_run(CPU_SAMPLE_GRID, args.cpu_steps, 1)
        cpu = {"value": r["cell_updates_per_s"], "unit": "cell-updates/s", "cores": r["cores"], "kind": "port",
               "sample": f"plume {CPU_SAMPLE_GRID}^3 bins ({r['n']} particles) x {args.cpu_steps} steps = {r['seconds']:.1f} s of the CPU oracle (OpenMP over bins)",
               "particle_steps_per_s": r["particle_steps_per_s"]}

    if rank == 0:
        line = {"metric": "cell-updates/s", "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": DTYPE, "dtype_note": DTYPE_NOTE, "data": "synthetic", "config": workload_config(G, n_total, world),
                "particle_steps_per_s": n_total / (ms_step * 1e-3), "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
                "roofline": roofline, "cpu_baseline": cpu}
        if halo:
            line["halo"] = halo
            line["per_rank"] = per_rank
        line.update(extra)
        line["parity_check"] = parity
        line["parity_note"] = ("free-running trajectories cannot be held to 1e-5: the reference's own two GPU runs differ by 1e-3..1e-2 after 10 steps "
                               "(float-atomic order amplified by the discontinuous friction / dead zone of FluidGPU.cuh:290-295, tests/golden/golden_noise.json); "
                               "the 1e-5 bar is applied per step from identical bits, here and in tests/")
        if slabp is not None:
            line["slab_parity"] = slabp
        if g1024 is not None:
            line["grid1024"] = g1024
        if refgpu is not None:
            line["reference_gpu"] = refgpu
        if frames is not None:
            line["frame_output"] = frames
        print(json.dumps(line))
    if solver is not None:
        solver.close()
    if world > 1:
        dist.destroy_process_group()


def e2e_leg(solver, torch, stream, args, G, world, barrier, dist):
    """Upload (pinned host -> device) + one step + download (device -> pinned host), every step."""
    import ctypes as C
    from fluidsolvergpu_b200 import FsgSoa
    slab = world > 1
    if slab:
        # every rank round-trips ALL the slots it holds (owned particles, particles about to migrate, empty slots marked in `cell`)
        solver.check()
        solver._check(solver._lib.fsg_slab_keep_foreign(solver._ctx, 1), "fsg_slab_keep_foreign")
    n = solver.stats()["n"]
    shapes = {"pos": (n, 3), "vel": (n, 3), "acc": (n, 3), "dens": (n,), "press": (n,), "delpress": (n, 3), "newdens": (n,),
              "newdelpress": (n, 3)}
    bufs = {k: torch.empty(s, dtype=torch.float32, pin_memory=True) for k, s in shapes.items()}
    bufs["index"] = torch.empty(n, dtype=torch.int32, pin_memory=True)
    bufs["boundary"] = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    if slab:
        bufs["cell"] = torch.empty(n, dtype=torch.int32, pin_memory=True)
    soa = FsgSoa()
    soa.n = n
    for k, b in bufs.items():
        setattr(soa, k, b.data_ptr())
    solver.download_raw(soa)          # the current state now lives in host memory
    # Per step only what the step reads goes up and only what it writes comes down (fsg_upload_soa / fsg_download_soa
    # skip NULL fields): `delpress` is pure output (set_delpress overwrites it, FluidGPU.cuh:276), and the accumulators
    # `newdens` / `newdelpress` are zero after every step (mykernel2 clears them, FluidGPU.cu:422-425) — the host copies
    # taken above already hold those zeros.
    up_skip, down_skip = ("delpress",), ("newdens", "newdelpress")
    soa_up, soa_down = FsgSoa(), FsgSoa()
    soa_up.n = soa_down.n = n
    for k, b in bufs.items():
        if k not in up_skip:
            setattr(soa_up, k, b.data_ptr())
        if k not in down_skip:
            setattr(soa_down, k, b.data_ptr())
    h2d = sum(b.numel() * b.element_size() for k, b in bufs.items() if k not in up_skip)
    d2h = sum(b.numel() * b.element_size() for k, b in bufs.items() if k not in down_skip)
    if not slab:
        d2h += 4 * n                  # + the new bin ids (`cell` is recomputed on upload, so it only travels down)
        bufs["cell"] = torch.empty(n, dtype=torch.int32, pin_memory=True)
        soa_down.cell = bufs["cell"].data_ptr()
    steps = args.e2e_steps

    def one():
        solver.upload_raw(soa_up)
        solver.step(1, sync=False)
        solver.download_raw(soa_down)  # synchronises
    one()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    barrier()
    dt = (time.perf_counter() - t0) / steps
    if slab:
        solver.check()
        t = torch.tensor([dt, float(h2d), float(d2h)], device="cuda", dtype=torch.float64)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dt, h2d, d2h = float(tmax[0]), int(t[1]), int(t[2])
        solver._check(solver._lib.fsg_slab_keep_foreign(solver._ctx, 0), "fsg_slab_keep_foreign")
        own = torch.tensor([float(solver.owned_count())], device="cuda", dtype=torch.float64)
        dist.all_reduce(own)
        if int(own[0]) != args.n_total:
            raise SystemExit(f"end-to-end leg lost particles: {int(own[0])} != {args.n_total}")
    return {"value": G ** 3 / dt, "unit": "cell-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "ms_per_step": dt * 1e3, "steps": steps,
            "what": ("fsg_upload_soa(pinned host: every field the step reads) + fsg_step(1) + fsg_download_soa(pinned host: every field "
                     "the step writes) per step, wall clock around synchronised calls"
                     if not slab else
                     "per rank: fsg_upload_soa(pinned host, every slot the rank holds) + slab exchange + fsg_step(1) + fsg_download_soa(pinned host) "
                     "per step; wall clock, max over ranks; bytes summed over ranks")}


def e2e_pipelined(fsg, torch, cfg, n, args, G):
    """End to end with HOST buffers, pipelined over independent batches: every step uploads the step's input from pinned host
    memory (fsg_upload_soa), runs fsg_step(1) and downloads the result to pinned host memory (fsg_download_soa).  `--e2e-contexts`
    contexts, each driven by its own host thread with the blocking C-ABI calls (the API's rule: one host thread per context), take
    the steps in turn, so the upload of one batch, the compute of another and the download of a third overlap (PCIe is full
    duplex).  Wall clock from the first upload to the last completed download; all K steps' copies are inside."""
    import threading
    from fluidsolvergpu_b200 import FsgSoa
    nctx = args.e2e_contexts
    steps = max(args.e2e_steps, nctx)
    solvers = [fsg.FluidSolver(cfg) for _ in range(nctx)]
    f3, f1 = ("pos", "vel", "acc", "delpress", "newdelpress"), ("dens", "press", "newdens")

    def host_set(fields):
        b = {k: torch.empty((n, 3) if k in f3 else (n,), dtype=torch.float32, pin_memory=True) for k in fields if k in f3 + f1}
        for k, dt in (("index", torch.int32), ("cell", torch.int32), ("boundary", torch.uint8)):
            if k in fields:
                b[k] = torch.empty(n, dtype=dt, pin_memory=True)
        return b

    def soa_of(bufs):
        soa = FsgSoa()
        soa.n = n
        for k, b in bufs.items():
            setattr(soa, k, b.data_ptr())
        return soa
    # what a step reads goes up, what it writes comes down: `delpress` is pure output (set_delpress overwrites it, FluidGPU.cuh:276);
    # the accumulators newdens / newdelpress are inputs (mykernel adds onto them) and are zero after every step (FluidGPU.cu:422-425)
    up_fields = ("pos", "vel", "acc", "dens", "press", "newdens", "newdelpress", "index", "boundary")
    down_fields = ("pos", "vel", "acc", "dens", "press", "delpress", "index", "boundary", "cell")
    inp = host_set(up_fields)
    solvers[0].scene_plume(SPACING, JITTER, SEED)
    solvers[0].step(1)
    solvers[0].download_raw(soa_of(inp))            # the batch: a developed plume state, now in pinned host memory
    outs = [host_set(down_fields) for _ in range(nctx)]
    soa_up, soa_down = soa_of(inp), [soa_of(o) for o in outs]
    h2d = sum(b.numel() * b.element_size() for b in inp.values())
    d2h = sum(b.numel() * b.element_size() for b in outs[0].values())

    def one(w):
        solvers[w].upload_raw(soa_up)
        solvers[w].step(1, sync=False)
        solvers[w].download_raw(soa_down[w])        # synchronises
    for w in range(nctx):
        one(w)                                      # warm-up: every context once (staging areas, first-touch)
    errors = []

    def worker(w):
        try:
            for _ in range(w, steps, nctx):
                one(w)
        except Exception as e:                      # noqa: BLE001
            errors.append(repr(e))
    threads = [threading.Thread(target=worker, args=(w,)) for w in range(nctx)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    if errors:
        raise SystemExit("end-to-end leg failed: " + errors[0])
    # every context processed the same batch: the new positions and bin ids do not depend on the order of the pair sums
    same = all(bool(torch.equal(outs[0][k], o[k])) for o in outs[1:] for k in ("pos", "cell", "index"))
    conserved = bool(torch.equal(torch.sort(outs[0]["index"]).values, torch.sort(inp["index"]).values))
    for s_ in solvers:
        s_.close()
    return {"value": G ** 3 / dt, "unit": "cell-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": dt * 1e3,
            "steps": steps, "contexts": nctx, "outputs_identical_across_contexts": same, "particles_conserved": conserved,
            "pcie_GBps_each_way": [h2d / dt / 1e9, d2h / dt / 1e9],
            "what": (f"{steps} independent batches (the same developed 512^3-class plume state in pinned host memory) over {nctx} contexts, one host thread "
                     "each: fsg_upload_soa(every field the step reads) + fsg_step(1) + fsg_download_soa(every field the step writes) per step; "
                     "uploads, compute and downloads of different batches overlap; wall clock / steps")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--grid", type=int, default=512)
    ap.add_argument("--impl", default="fsg", choices=["fsg", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=12)
    ap.add_argument("--cpu-steps", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle-sampled parity check (N=1) / the slab-vs-single check (N>1)")
    ap.add_argument("--parity-bins", type=int, default=256)
    ap.add_argument("--no-frames", action="store_true", help="skip the frame-output leg (N=1)")
    ap.add_argument("--no-reference-gpu", action="store_true", help="skip timing the reference's own CUDA kernels (oracle/_ref) beside libfsg")
    ap.add_argument("--grid1024-steps", type=int, default=5, help="steps of the short 1024^3 leg (0 = off)")
    ap.add_argument("--e2e-contexts", type=int, default=4,
                    help="N=1: contexts the end-to-end leg pipelines independent batches over (1 = strictly serial upload/step/download)")
    ap.add_argument("--pair-mode", type=int, default=0, choices=[0, 1],
                    help="0: symmetric pair kernel (default); 1: deterministic gather kernel (fsg_config.pair_mode)")
    ap.add_argument("--overlap", action="store_true",
                    help="N>1, peer exchange: boundary bins first, next step's pack + copies on a second stream beside the interior bins "
                         "(measured slower than the plain order at 512^3: splitting the pair kernel costs more than the exchange it hides)")
    ap.add_argument("--classic-slabs", action="store_true", help="N > 1: the classic slab pipeline (ghosts appended and sorted) instead of sorted ghosts")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N>1: peer = messages copied into the neighbours' memory over NVLink (CUDA IPC) + 4-byte NCCL signal; "
                         "nccl = whole messages through NCCL send/recv")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "fsg":
        args.warmup = 3
    if args.impl == "reference":
        reference_arm(args)
    else:
        fsg_arm(args)


if __name__ == "__main__":
    main()
