timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_r1h.log 2>&1; tail -4 gpurun_out/pytest_r1h.log | cut -c1-250
for N in 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --grid 256 --steps 10 --warmup 3 > gpurun_out/bench256_n${N}_r1h.json 2> gpurun_out/bench256_n${N}_r1h.err; echo "n=$N 256 rc=$?"; tail -5 gpurun_out/bench256_n${N}_r1h.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench512_n${N}_r1h.json 2> gpurun_out/bench512_n${N}_r1h.err; echo "n=$N 512 rc=$?"; tail -5 gpurun_out/bench512_n${N}_r1h.err | cut -c1-300
done
timeout 600 python bench.py --no-cpu > gpurun_out/bench512_n1_r1h.json 2> gpurun_out/bench512_n1_r1h.err; echo rc=$?
