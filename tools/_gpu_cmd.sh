timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --no-cpu --e2e-steps 1 > gpurun_out/bench512_pk2.json 2> gpurun_out/bench512_pk2.err; echo "rc=$?"
python - <<'PY'
import json
for t in ("pk2",):
    try:
        j=json.load(open(f"gpurun_out/bench512_{t}.json")); print(t, j["ms_per_step"], j["phases_ms"])
    except Exception as e: print(t, "fail", e)
PY
