timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python bench.py --no-cpu > gpurun_out/bench512_e2e2.json 2> gpurun_out/bench512_e2e2.err; echo "rc=$?"; tail -2 gpurun_out/bench512_e2e2.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/bench512_e2e2.json")); print(j["ms_per_step"], j["e2e"])
PY
