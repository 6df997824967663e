timeout 300 python -m pytest tests -m gpu -q > gpurun_out/final_test.log 2>&1; grep -E "^E  " gpurun_out/final_test.log | cut -c1-250 | head -6; tail -2 gpurun_out/final_test.log
