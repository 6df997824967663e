timeout 600 python -m pytest tests -m gpu -q -x -k "full_size" > gpurun_out/fullsize.log 2>&1; grep -E "^E  " gpurun_out/fullsize.log | cut -c1-300 | head -8; tail -2 gpurun_out/fullsize.log
