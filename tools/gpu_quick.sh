#!/bin/bash
# quick GPU pass: selected parity tests + resident timing (quick_bench) at given grids.  Usage: bash tools/gpu_quick.sh TAG "GRIDS" [pytest -k expr]
TAG=${1:-q}; GRIDS=${2:-"256 512"}; K=${3:-""}
O=gpurun_out; mkdir -p $O
if [ -n "$K" ]; then (timeout 900 python -m pytest tests -m gpu -x -q -k "$K") > $O/pytest_$TAG.log 2>&1; else (timeout 1200 python -m pytest tests -m gpu -x -q) > $O/pytest_$TAG.log 2>&1; fi
echo "pytest rc=$?"; tail -6 $O/pytest_$TAG.log
timeout 600 python tools/quick_bench.py $GRIDS 2>&1 | tee $O/quick_$TAG.log
