for rep in 1 2; do for f in 0 1; do
VFD_NARROW_WGRAD=$f python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-profile --no-flow > gpurun_out/r2ak_bench_$f.json 2> gpurun_out/r2ak_bench.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/r2ak_bench_$f.json').read().strip().splitlines()[-1]); print('NARROW_WGRAD=$f ms_per_step', d['ms_per_step'])"
done; done
