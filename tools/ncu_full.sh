#!/bin/bash
# one ncu --set full capture of a kernel (regex) from tools/quick_bench.py; raw + source pages as CSV.  Usage: bash tools/ncu_full.sh GRID REGEX TAG [skip]
G=${1:-256}; RX=${2:-k_pair_v3}; TAG=${3:-r2}; SKIP=${4:-4}; O=gpurun_out; mkdir -p $O
CMD="python tools/quick_bench.py $G"
$CMD > $O/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$RX -s $SKIP -c 1 -f -o $O/prof_$TAG $CMD > $O/ncu_full_$TAG.log 2>&1; echo "ncu rc=$?"
ncu -i $O/prof_$TAG.ncu-rep --page raw --csv > $O/prof_${TAG}_raw.csv 2>/dev/null
ncu -i $O/prof_$TAG.ncu-rep --page source --csv > $O/prof_${TAG}_source.csv 2>/dev/null
ls -la $O/prof_$TAG*; rm -f $O/prof_$TAG.ncu-rep
