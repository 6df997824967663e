timeout 900 python tools/ref_gpu_compare.py > gpurun_out/ref_gpu_compare2.json 2> gpurun_out/ref_gpu_compare2.err; echo "rc=$?"; tail -3 gpurun_out/ref_gpu_compare2.err; python -c "
import json; j=json.load(open('gpurun_out/ref_gpu_compare2.json'))
for k,v in j.items(): print(k, v.get('libfsg_ms_per_step'), v.get('reference_gpu_ms_per_step'), v.get('speedup'), str(v.get('reference_detail'))[:200])"
