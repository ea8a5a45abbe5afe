"""Time the x2 trilinear upsample forward / backward on the decoder's shapes.   python tools/gpu_time_upsample.py [reps]"""
import sys
import torch
sys.path.insert(0, ".")
from vfd_gan_b200 import ops

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for (N, D, H, W, C) in [(32, 8, 56, 56, 64), (32, 4, 28, 28, 128), (32, 2, 14, 14, 256)]:
    x = torch.randn(N, D, H, W, C, device=dev).bfloat16().requires_grad_(True)
    tf, tb = [], []
    for i in range(reps + 1):
        flush.zero_()
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record()
        y = ops.UpsampleFn.apply(x)
        b.record()
        g = torch.ones_like(y)
        flush.zero_()
        b2 = torch.cuda.Event(enable_timing=True)
        b2.record()
        y.backward(g)
        c.record()
        torch.cuda.synchronize()
        if i:
            tf.append(a.elapsed_time(b))
            tb.append(b2.elapsed_time(c))
        x.grad = None
    out_bytes = y.numel() * 2 + x.numel() * 2
    mf, mb = sorted(tf)[len(tf) // 2], sorted(tb)[len(tb) // 2]
    print(f"{(N, D, H, W, C)}: fwd {mf:.3f} ms {out_bytes / mf / 1e6:.0f} GB/s | bwd {mb:.3f} ms {out_bytes / mb / 1e6:.0f} GB/s", flush=True)
