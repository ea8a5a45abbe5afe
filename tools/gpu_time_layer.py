"""Time one conv layer geometry (forward / dgrad / wgrad) on the device with CUDA events.
    python tools/gpu_time_layer.py cin cout kd kh kw N D H W [reps]"""
import sys
import torch
sys.path.insert(0, ".")
from vfd_gan_b200 import ops

cin, cout, kd, kh, kw, N, D, H, W = map(int, sys.argv[1:10])
reps = int(sys.argv[10]) if len(sys.argv) > 10 else 5
dev = "cuda"
x = torch.randn(N, D, H, W, ops.round_up(cin, 8), device=dev).bfloat16()
x[..., cin:] = 0
w = (torch.randn(cout, cin, kd, kh, kw, device=dev) * 0.05).requires_grad_(True)
xg = x.clone().requires_grad_(True)
y = ops.ConvFn.apply(xg, w, None, False, False)
gy = torch.randn_like(y)
gy[..., cout:] = 0
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


class P:
    def __init__(self):
        self.t = {}

    def run(self, kind, work, thunk, nbytes=0.0):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        thunk()
        b.record()
        self.t.setdefault(kind, []).append((a, b, work))


prof = P()
for i in range(reps + 2):
    flush.zero_()
    ops.PROFILER = prof if i >= 2 else None
    y = ops.ConvFn.apply(xg, w, None, False, False)
    y.backward(gy)
torch.cuda.synchronize()
for k, v in prof.t.items():
    ms = sorted(a.elapsed_time(b) for a, b, _ in v)[len(v) // 2]
    print(f"{k:11s} {ms:7.3f} ms  {v[0][2] / (ms * 1e-3) / 1e12:8.1f} TF/s")
