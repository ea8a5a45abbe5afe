"""Kernel time of conv_last forward (32 -> 1 channels, 3x3x3, fp32 logits) at the bench geometry: the narrow kernel
(csrc/conv_narrow.cu) against the tcgen05 tile kernel, L2 flushed between launches."""
import os
import sys
if os.environ.get("VFD_NARROW_DBG"):
    os.environ["VFD_DEBUG_LIB"] = "1"      # the stage switches exist only in libvfd_b200_debug.so
import torch
sys.path.insert(0, ".")
from vfd_gan_b200 import ops

dev = "cuda"
N, D, H, W = 32, 16, 112, 112
x = torch.randn(N, D, H, W, 32, device=dev).bfloat16()
w = torch.randn(1, 32, 3, 3, 3, device=dev) * 0.1
pk = ops._packed(w)
b = torch.zeros(pk.fwd.shape[0], device=dev)
out = torch.empty(N, D, H, W, 8, dtype=torch.float32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=5):
    ts = []
    for i in range(reps + 1):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        if i:
            ts.append(a.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]


t_narrow = timeit(lambda: ops.conv3d_fwd_narrow(x, pk.fwd, b, out))
ref = out.clone()
t_tile = timeit(lambda: ops.conv3d_fwd(x, pk.fwd, b, out, None, 3, 3, 3, pk.kc_f, 8, False))
alg = x.numel() * 2 + N * D * H * W * 4
print(f"narrow {t_narrow:.3f} ms ({alg / t_narrow / 1e6:.0f} GB/s algorithmic)  tcgen05 tile {t_tile:.3f} ms  "
      f"max |diff| {float((ref - out).abs().max()):.2e}")
