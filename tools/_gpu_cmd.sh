timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python bench.py --no-cpu --e2e-steps 1 > gpurun_out/bench512_v3i.json 2> gpurun_out/bench512_v3i.err; echo "rc=$?"
python - <<'PY'
import json
for t in ("v3i",):
    try:
        j=json.load(open(f"gpurun_out/bench512_{t}.json")); print(t, j["ms_per_step"], j["phases_ms"], j["roofline"]["kernel"])
    except Exception as e: print(t, "fail", e)
PY
CMD="python bench.py --grid 256 --steps 2 --warmup 3 --no-cpu --e2e-steps 1"
$CMD > gpurun_out/plain_r1q.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_pair_v3 -s 3 -c 1 -f -o gpurun_out/prof_pair_v3_256 $CMD > gpurun_out/ncu_full_r1q.log 2>&1
echo "ncu rc=$?"
