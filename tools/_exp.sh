for d in 0 15; do echo -n "dbg=$d: "; VFD_NARROW_DBG=$d python tools/gpu_time_conv_last.py 2>&1 | tail -1; done
python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "narrow" 2>&1 | grep -E "^E  |passed|failed|Error" | head
