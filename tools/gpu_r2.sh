#!/bin/bash
# One GPU-box pass of round 2: parity tests + the full bench line.  Usage (under gpurun): bash tools/gpu_r2.sh [tag]
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > $O/gpu_$TAG.txt 2>&1
free -g | head -2 >> $O/gpu_$TAG.txt; nproc >> $O/gpu_$TAG.txt
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_$TAG.log
( time timeout 900 python bench.py ) > $O/bench512_$TAG.json 2> $O/bench512_$TAG.err; echo "bench512 rc=$?"; tail -c 3000 $O/bench512_$TAG.json; tail -5 $O/bench512_$TAG.err
