timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 300 python bench.py --no-cpu --e2e-steps 1 > gpurun_out/bench512_fmx.json 2> gpurun_out/bench512_fmx.err; echo "rc=$?"
