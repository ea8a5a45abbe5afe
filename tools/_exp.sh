for i in 1 2; do python -m pytest tests/ -x -q -m gpu 2>&1 | grep -E "^E  |passed|failed" | head -12; done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-profile --no-flow > gpurun_out/r2ac_bench.json 2> gpurun_out/r2ac_bench.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/r2ac_bench.json').read().strip().splitlines()[-1]); print('ms_per_step', d['ms_per_step'], 'launches', d['gpu_launches'])"
