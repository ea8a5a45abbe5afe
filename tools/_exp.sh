PROBE_FLAGS=0,1,2,4,8,32 python tools/gpu_stage_probe.py G.dconv1.s S.dconv1.s S.dconv2.s G.dconv1.t conv_last uconv1.t > gpurun_out/r2u_stage_probe.txt 2>&1
cat gpurun_out/r2u_stage_probe.txt
