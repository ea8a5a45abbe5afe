python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r2z_full_gpu.log; tail -4 gpurun_out/r2z_full_gpu.log
python bench.py > gpurun_out/r2z_bench_default.json 2> gpurun_out/r2z_bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2z_bench_default.json","gpurun_out/r2z_bench_reference.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ("value","ms_per_step","impl","gpu_launches")}, d.get("e2e"), d.get("cpu_baseline"))
        r=d.get("roofline")
        if r: print({k:r[k] for k in ("kind","frac","traffic")}, r.get("all_conv"), {k:(v["frac"],v["ms_per_step"],v["traffic_over_algorithmic"]) for k,v in r["per_kind"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
