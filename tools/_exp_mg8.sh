python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2mg_n8_cfg2.json 2> gpurun_out/r2mg_n8_cfg2.err
echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2mg_n8_cfg2.json').read().strip().splitlines()[-1])
print('n', d['n_gpus'], 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks'])
PY
