N=$1
for wl in cfg2 cfg3 cfg4 cfg5; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --workload $wl > gpurun_out/r2mg_n${N}_$wl.json 2> gpurun_out/r2mg_n${N}_$wl.err
  echo "$wl rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2mg_n${N}_$wl.json').read().strip().splitlines()[-1])
    print('$wl', 'n', d['n_gpus'], 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'])
except Exception as e:
    print('$wl ERR', e)
PY
done
