for cfg in "NCCL_DEBUG=WARN" "NCCL_MAX_CTAS=8" "NCCL_MAX_CTAS=4" "NCCL_MAX_CTAS=2"; do
env $cfg python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-profile --no-flow > gpurun_out/r2mg_tmp.json 2> gpurun_out/r2mg_tmp.err
python - "$cfg" <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2mg_tmp.json').read().strip().splitlines()[-1])
    print(sys.argv[1], 'n', d['n_gpus'], 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
