N=${N:-2}; G=${G:-512}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --grid $G --steps 10 --warmup 3 --e2e-steps 1 > gpurun_out/bench${G}_n${N}_v3d.json 2> gpurun_out/bench${G}_n${N}_v3d.err; echo "rc=$?"
python - <<PY
import json
t=open("gpurun_out/bench${G}_n${N}_v3d.json").read(); j=json.loads(t[t.index('{'):])
print(j["ms_per_step"], {k:[round(x,3) for x in v] for k,v in j["per_rank"].items() if k.startswith("ms_")})
PY
