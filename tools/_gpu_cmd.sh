N=2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --grid 256 --steps 10 --warmup 3 > "gpurun_out/bench256_n2_e2e.json" 2> "gpurun_out/bench256_n2_e2e.err"; echo "rc=$?"; tail -4 "gpurun_out/bench256_n2_e2e.err" | cut -c1-300
