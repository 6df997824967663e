timeout 300 python -m pytest tests -m gpu -q -k "uncapped or plume or large_bins" 2>&1 | tail -1
timeout 300 python bench.py --no-cpu --e2e-steps 1 > gpurun_out/bench512_bps7.json 2> gpurun_out/bench512_bps7.err; echo "rc=$?"
