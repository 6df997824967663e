O=gpurun_out
timeout 900 python bench.py > $O/bench512_r1final.json 2> $O/bench512_r1final.err; echo "bench512 rc=$?"; tail -2 $O/bench512_r1final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/benchref_r1final.json 2> $O/benchref_r1final.err; echo "benchref rc=$?"; cat $O/benchref_r1final.json | cut -c1-400
CMD="python bench.py --grid 256 --steps 2 --warmup 3 --no-cpu --e2e-steps 1"
$CMD > $O/plain_r1final.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r1final.csv $CMD > $O/ncu_list_r1final.log 2>&1
echo "ncu list rc=$?"
