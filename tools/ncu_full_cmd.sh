#!/bin/bash
# one ncu --set full capture of a kernel (regex) from an arbitrary command; raw + source pages as CSV.  Usage: bash tools/ncu_full_cmd.sh TAG REGEX SKIP cmd...
TAG=$1; RX=$2; SKIP=$3; shift 3; O=gpurun_out; mkdir -p $O
"$@" > $O/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$RX -s $SKIP -c 1 -f -o $O/prof_$TAG "$@" > $O/ncu_full_$TAG.log 2>&1; echo "ncu rc=$?"
ncu -i $O/prof_$TAG.ncu-rep --page raw --csv > $O/prof_${TAG}_raw.csv 2>/dev/null
ncu -i $O/prof_$TAG.ncu-rep --page source --csv > $O/prof_${TAG}_source.csv 2>/dev/null
ls -la $O/prof_$TAG*; rm -f $O/prof_$TAG.ncu-rep
