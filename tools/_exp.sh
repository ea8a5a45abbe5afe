python -m pytest tests/test_data_gpu.py -x -q -m gpu -k training_state 2>&1 | tail -5 > gpurun_out/r2p_state_test.log
python bench.py --steps 1 --warmup 3 --no-profile --no-cpu-baseline --no-flow > gpurun_out/r2p_b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2p_launches_raw.csv python bench.py --steps 1 --warmup 3 --no-profile --no-cpu-baseline --no-flow > gpurun_out/r2p_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r2p_launches_raw.csv gpurun_out/r2p > gpurun_out/r2p_summary.log 2>&1
python tools/gpu_glue_trace.py > gpurun_out/r2p_glue.log 2>&1
tail -3 gpurun_out/r2p_state_test.log; head -40 gpurun_out/r2p_step_kernel_summary.csv
