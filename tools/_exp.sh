python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "conv or small_nets or spatiotemporal or netg or netd" 2>&1 | tail -5
for b in 0 1; do
echo "== VFD_RES_DIRECT=$b"
VFD_RES_DIRECT=$b PROBE_FLAGS=0 python tools/gpu_stage_probe.py G.dconv1.s S.dconv1.s S.dconv2.s G.dconv1.t uconv1.t T.dconv3.t conv_last T.dconv1.s 2>&1 | grep -v wgrad
done > gpurun_out/r2w_direct.txt
cat gpurun_out/r2w_direct.txt
for b in 0 1; do
VFD_RES_DIRECT=$b python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-profile --no-flow > gpurun_out/r2w_bench_$b.json 2> gpurun_out/r2w_bench.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/r2w_bench_$b.json').read().strip().splitlines()[-1]); print('DIRECT=$b ms_per_step', d['ms_per_step'])"
done
