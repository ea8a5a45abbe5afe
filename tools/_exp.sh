python tools/gpu_time_bn.py 3 outer > gpurun_out/r2y_bn_default.txt 2>&1
VFD_BN_BWD8_POOL=1 python tools/gpu_time_bn.py 3 outer > gpurun_out/r2y_bn_pool8.txt 2>&1
echo default; cat gpurun_out/r2y_bn_default.txt; echo pool8; cat gpurun_out/r2y_bn_pool8.txt
VFD_BN_BWD8_POOL=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-profile --no-flow > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/r2y_bench.json').read().strip().splitlines()[-1]); print('POOL8 ms_per_step', d['ms_per_step'])"
