python -m pytest tests/test_data_gpu.py -x -q -m gpu -k training_state 2>&1 | tail -15 > gpurun_out/r2o_state_test.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-profile --no-flow"
for cfg in "default" "VFD_BN_WAVES=1" "VFD_BN_WAVES=2" "VFD_ASYNC_WGRAD=0" "VFD_BN_WAVES=1 VFD_ASYNC_WGRAD=0"; do
  if [ "$cfg" = "default" ]; then env $B > gpurun_out/r2o_tmp.json 2>gpurun_out/r2o_tmp.err; else env $cfg $B > gpurun_out/r2o_tmp.json 2>gpurun_out/r2o_tmp.err; fi
  python - "$cfg" <<'PY' >> gpurun_out/r2o_overlap_experiment.txt
import json,sys
try:
    d=json.loads(open('gpurun_out/r2o_tmp.json').read().strip().splitlines()[-1])
    print(sys.argv[1], 'ms_per_step', round(d['ms_per_step'],3), 'e2e_ms', round(d['e2e']['ms_per_step'],3))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
done
cat gpurun_out/r2o_overlap_experiment.txt; tail -3 gpurun_out/r2o_state_test.log
