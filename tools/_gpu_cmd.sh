N=${N:-2}; G=${G:-512}
if [ "$N" = "1" ]; then
timeout 900 python bench.py --grid $G --steps 10 --warmup 3 --e2e-steps 1 --no-cpu > gpurun_out/bench${G}_n${N}_v3.json 2> gpurun_out/bench${G}_n${N}_v3.err; echo "rc=$?"
else
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --grid $G --steps 10 --warmup 3 --e2e-steps 1 > gpurun_out/bench${G}_n${N}_v3.json 2> gpurun_out/bench${G}_n${N}_v3.err; echo "rc=$?"
fi
tail -2 gpurun_out/bench${G}_n${N}_v3.err; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench${G}_n${N}_v3.json | head -2
