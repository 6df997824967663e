#!/bin/bash
# One GPU-box pass: parity tests, reference goldens, bench lines, ncu launch list + full capture.
# Usage (from the repo root, under gpurun): bash tools/gpu_check.sh [tag]
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_$TAG.log
if [ -z "$SKIP_GOLDEN" ]; then timeout 600 python tools/make_golden.py $O/golden > $O/golden_$TAG.log 2>&1; echo "golden rc=$?"; tail -2 $O/golden_$TAG.log; fi
timeout 600 python bench.py --grid 256 > $O/bench256_$TAG.json 2> $O/bench256_$TAG.err; echo "bench256 rc=$?"; cat $O/bench256_$TAG.json; tail -3 $O/bench256_$TAG.err
timeout 900 python bench.py > $O/bench512_$TAG.json 2> $O/bench512_$TAG.err; echo "bench512 rc=$?"; cat $O/bench512_$TAG.json; tail -3 $O/bench512_$TAG.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/benchref_$TAG.json 2> $O/benchref_$TAG.err; echo "benchref rc=$?"; cat $O/benchref_$TAG.json
if [ -z "$SKIP_NCU" ]; then
CMD="python bench.py --grid 256 --steps 2 --warmup 3 --no-cpu --e2e-steps 1"
$CMD > $O/plain_$TAG.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv $CMD > $O/ncu_list_$TAG.log 2>&1
echo "ncu list rc=$?"
CMD2="python tools/quick_bench.py 128"
$CMD2 > $O/plain2_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_pair_v2 -s 4 -c 2 -f -o $O/prof_pair_$TAG $CMD2 > $O/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
fi
