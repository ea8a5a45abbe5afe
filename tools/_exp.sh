true && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2x_dram_raw.csv python tools/gpu_one_step.py 2 > gpurun_out/r2x_ncu.log 2>&1
python tools/dram_summary.py gpurun_out/r2x_dram_raw.csv gpurun_out/r2x
head -30 gpurun_out/r2x_dram_per_kernel.csv
