#!/bin/bash
# Multi-GPU pass: the multi-process slab parity test + one bench line at N GPUs.  Usage (under gpurun --gpus N): bash tools/gpu_r2_multi.sh N [tag]
N=${1:-2}
TAG=${2:-r2}
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) > $O/pytest_multi_$TAG.log 2>&1; echo "pytest multi rc=$?"; tail -15 $O/pytest_multi_$TAG.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N ) > $O/bench512_n${N}_$TAG.json 2> $O/bench512_n${N}_$TAG.err; echo "bench rc=$?"; tail -c 2500 $O/bench512_n${N}_$TAG.json; tail -5 $O/bench512_n${N}_$TAG.err
