python -m pytest tests/test_extensions_gpu.py -x -q -m gpu -k "convlstm or lstm" 2>&1 | grep -E "^E  |passed|failed|Error" | head -20
python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "convlstm or conv_fwd_dgrad" 2>&1 | grep -E "^E  |passed|failed|Error" | head
for f in 0 1; do
VFD_LSTM_FUSED=$f python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-profile --no-flow > gpurun_out/r2ag_cfg3_$f.json 2> gpurun_out/r2ag_cfg3.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/r2ag_cfg3_$f.json').read().strip().splitlines()[-1]); print('cfg3 LSTM_FUSED=$f ms_per_step', d['ms_per_step'])"
done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-profile --no-flow > gpurun_out/r2ag_cfg2.json 2> gpurun_out/r2ag_cfg2.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/r2ag_cfg2.json').read().strip().splitlines()[-1]); print('cfg2 ms_per_step', d['ms_per_step'])"
