ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2f_dram_raw.csv python tools/gpu_one_step.py 2 > gpurun_out/r2f_ncu.log 2>&1
python tools/dram_summary.py gpurun_out/r2f_dram_raw.csv gpurun_out/r2f > gpurun_out/r2f_kind.log
head -24 gpurun_out/r2f_dram_per_kernel.csv
ncu --set full --clock-control none --import-source on -k regex:conv_narrow_fwd -c 1 -s 2 -o gpurun_out/r2f_narrow python tools/gpu_time_conv_last.py > gpurun_out/r2f_ncu2.log 2>&1
ncu -i gpurun_out/r2f_narrow.ncu-rep --page raw --csv > gpurun_out/r2f_narrow_raw.csv 2>/dev/null
ls -la gpurun_out/r2f_narrow*
