python -m pytest tests/ -x -q -m gpu 2>&1 | grep -E "^E  |passed|failed" | head -8
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-flow --dump-kernels gpurun_out/r2ah_kernels.json > gpurun_out/r2ah_bench.json 2> gpurun_out/r2ah_bench.err
python tools/per_launch_roofline.py gpurun_out/r2ah_kernels.json gpurun_out/r2ah_per_launch_roofline.csv
