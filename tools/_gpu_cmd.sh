for N in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench512_n${N}_peer.json 2> gpurun_out/bench512_n${N}_peer.err; echo "n=$N rc=$?"; tail -2 gpurun_out/bench512_n${N}_peer.err | cut -c1-200
done
