cd fluidsolvergpu_b200; cp libfsg.so libfsg_w4.so; cd ..
for v in b4 b2; do
cp fluidsolvergpu_b200/libfsg_$v.so fluidsolvergpu_b200/libfsg.so
timeout 300 python bench.py --no-cpu --e2e-steps 1 > gpurun_out/bench512_$v.json 2> gpurun_out/bench512_$v.err; echo "rc=$?"
python - <<PY
import json
j=json.load(open("gpurun_out/bench512_$v.json")); print("$v", j["ms_per_step"], j["phases_ms"])
PY
done
