python -m pytest tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -40 > gpurun_out/r2r_parity.log
tail -30 gpurun_out/r2r_parity.log
for cfg in "VFD_DETERMINISTIC=0" "VFD_DETERMINISTIC=1"; do
env $cfg python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-profile --no-flow > gpurun_out/r2r_bench_$cfg.json 2> gpurun_out/r2r_bench.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/r2r_bench_$cfg.json').read().strip().splitlines()[-1]); print('$cfg ms_per_step', d['ms_per_step'])"
done
