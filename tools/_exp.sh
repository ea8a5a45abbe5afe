VFD_NARROW_DBG=16 python tools/gpu_time_conv_last.py 2>&1 | tail -3
