"""The bench.py output contract, checked on the JSON lines kept under profiles/ (written by bench.py on a B200): every key the
driver reads is there, and the derived quantities are consistent with each other."""
import json
import pathlib

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
LINES = ["r1_bench512_n1_sortmerge.json", "r1_bench512_n1_sym_e2e.json", "r1_bench512_n2_sym.json", "r1_bench512_n8_sym.json",
         "r1_bench1024_n8_sym.json"]


def load(name):
    t = (ROOT / "profiles" / name).read_text()
    return json.loads(t[t.index("{"):])


@pytest.mark.parametrize("name", LINES)
def test_bench_line_has_the_contract_keys(name):
    j = load(name)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in j, k
    assert j["metric"] == "cell-updates/s" and j["unit"] == "cell-updates/s" and j["higher_is_better"] is True
    assert j["scaling"] == "strong" and j["vs_baseline"] is None and j["data"] == "synthetic" and j["dtype"] == "f32"
    assert "workload" in j["config"] and "model" not in j["config"]
    G = j["config"]["grid"]
    assert abs(j["value"] - G ** 3 / (j["ms_per_step"] * 1e-3)) <= 1e-6 * j["value"]      # whole-job cell updates per second
    assert j["gpu_launches"] > 0 and j["warmup"] >= 3
    r = j["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) <= 1e-9
    e = j["e2e"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in e, k
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < j["value"]
    c = j["clocks"]
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(c)
    assert not [x for x in c["reasons"] if "slowdown" in x]
    if j["n_gpus"] == 1 and j["cpu_baseline"] is not None:
        assert set(("value", "unit", "cores", "kind", "sample")) <= set(j["cpu_baseline"])


def test_reference_arm_line():
    j = load("r1_benchref_sym.json")
    assert j["impl"] == "reference" and j["metric"] == "cell-updates/s" and j["cpu_baseline"]["kind"] in ("port", "reference")
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0 and j["e2e"]["value"] == j["value"]
