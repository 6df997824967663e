timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python bench.py --no-cpu --e2e-steps 1 --steps 5 > gpurun_out/bench512_last.json 2> gpurun_out/bench512_last.err; echo "rc=$?"
python - <<'PY'
import json
j=json.load(open("gpurun_out/bench512_last.json")); print(j["ms_per_step"], j["phases_ms"], j["e2e"]["ms_per_step"])
PY
