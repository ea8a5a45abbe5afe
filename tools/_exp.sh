python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "narrow or conv_fwd_dgrad" 2>&1 | grep -E "^E  |passed|failed|Error" | head -12
for f in 0 1; do
VFD_NARROW_CONV=$f python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-profile --no-flow > gpurun_out/r2ai_bench_$f.json 2> gpurun_out/r2ai_bench.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/r2ai_bench_$f.json').read().strip().splitlines()[-1]); print('NARROW=$f ms_per_step', d['ms_per_step'])"
done
