python -m pytest tests/ -x -q -m gpu 2>&1 | grep -E "^E  |passed|failed" | head -8
python bench.py > gpurun_out/r2al_bench_default.json 2> gpurun_out/r2al_bench_default.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2al_bench_default.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['cpu_baseline']['value'], d['roofline']['all_conv'], d['roofline']['frac'], d['clocks'])"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
