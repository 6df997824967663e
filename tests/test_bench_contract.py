"""bench.py's own code, run on the CPU: the reference arm end to end on a tiny grid (its JSON line carries the contract keys and
names the grid it actually timed), the workload description, and the oracle-sampled parity check driven by a stand-in solver."""
import argparse
import io
import json
import contextlib
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
import oracle_py  # noqa: E402


def test_reference_arm_line_names_the_grid_it_timed(monkeypatch):
    monkeypatch.setattr(bench, "CPU_SAMPLE_GRID", 16)
    monkeypatch.delenv("RANK", raising=False)
    args = argparse.Namespace(steps=2, warmup=1, grid=512, gpus=1)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        bench.reference_arm(args)
    j = json.loads(out.getvalue())
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in j, k
    assert j["impl"] == "reference" and j["metric"] == "cell-updates/s" and j["higher_is_better"] is True and j["vs_baseline"] is None
    assert j["config"]["grid"] == 16 and "16^3" in j["config"]["workload"] and "512^3" in j["config"]["sample_of"]
    assert j["config"]["particles"] > 0 and "model" not in j["config"]
    assert abs(j["value"] - 16 ** 3 / (j["ms_per_step"] * 1e-3)) <= 1e-6 * j["value"]
    e = j["e2e"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0 and e["value"] == j["value"] and e["unit"] == j["unit"]
    c = j["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] == j["value"] and "16^3" in c["sample"]


def test_reference_arm_is_silent_on_other_ranks(monkeypatch):
    monkeypatch.setenv("RANK", "1")
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        bench.reference_arm(argparse.Namespace(steps=1, warmup=0, grid=512, gpus=2))
    assert out.getvalue() == ""


def test_workload_config_has_no_model_keys():
    c = bench.workload_config(512, 123, 8)
    assert "workload" in c and c["grid"] == 512 and c["particles"] == 123 and "8 x-slabs" in c["decomposition"]
    assert "model" not in c and "seq_len" not in c and "l2_policy" in c


class _OracleBackedSolver:
    """Stands in for FluidSolver in parity_sample: download() / step() served by the CPU oracle, optionally with a fault injected."""

    def __init__(self, cfg, state, corrupt=None):
        self.cfg, self.corrupt = cfg, corrupt
        self.sim = oracle_py.OracleSim(oracle_py.params_from_cfg(cfg), state)
        self.sim.step(1)

    def download(self):
        st = self.sim.state()
        if self.corrupt and self.sim_steps > 0:
            st = self.corrupt(st)
        return st

    sim_steps = 0

    def step(self, k):
        self.sim.step(k)
        self.sim_steps += k


def _plume(grid=12):
    import fluidsolvergpu_b200 as fsg
    cfg = fsg.scenes.plume_config(grid)
    return cfg, fsg.scenes.plume_scene(cfg)


def test_parity_sample_passes_on_identical_results_and_covers_home_bins():
    cfg, state = _plume()
    rec = bench.parity_sample(_OracleBackedSolver(cfg, state), cfg, nbins=16, seed=3)
    assert rec["ok"] and rec["ints_equal"] and rec["positions_bit_equal"] and rec["max_rel_l2"] == 0.0
    assert rec["bins"] == 16 and rec["particles_compared"] >= 16 and rec["sub_scene_particles"] > rec["particles_compared"]


def test_parity_sample_detects_a_wrong_field_and_a_wrong_bin():
    cfg, state = _plume()

    def bad_dens(st):
        st = dict(st)
        st["dens"] = st["dens"] * np.float32(1.001)
        return st

    def bad_cell(st):
        st = dict(st)
        st["cell"] = st["cell"] + 1
        return st
    rec = bench.parity_sample(_OracleBackedSolver(cfg, state, bad_dens), cfg, nbins=16, seed=3)
    assert not rec["ok"] and rec["ints_equal"] and rec["rel_l2"]["dens"] > 1e-4
    rec = bench.parity_sample(_OracleBackedSolver(cfg, state, bad_cell), cfg, nbins=16, seed=3)
    assert not rec["ok"] and not rec["ints_equal"]
