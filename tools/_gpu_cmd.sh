timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_r1g.log 2>&1; tail -5 gpurun_out/pytest_r1g.log | cut -c1-250
timeout 600 python bench.py --grid 256 --no-cpu > gpurun_out/bench256_r1g.json 2> gpurun_out/bench256_r1g.err; echo rc=$?; tail -2 gpurun_out/bench256_r1g.err
timeout 600 python bench.py --no-cpu > gpurun_out/bench512_r1g.json 2> gpurun_out/bench512_r1g.err; echo rc=$?; tail -2 gpurun_out/bench512_r1g.err
