timeout 700 python -m pytest tests -m gpu -q -k "slab" 2>&1 | grep -E "^E  |passed|failed|Error" | cut -c1-300 | head
N=2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > "gpurun_out/bench512_n2_rg.json" 2> "gpurun_out/bench512_n2_rg.err"; echo "rc=$?"; tail -4 "gpurun_out/bench512_n2_rg.err" | cut -c1-300
