"""Time the fused BatchNorm+activation(+pool) forward / backward kernels on the step's largest shapes.
    python tools/gpu_time_bn.py [reps] [case-name-substring ...]"""
import sys
import torch
sys.path.insert(0, ".")
from vfd_gan_b200 import ops
from vfd_gan_b200.spatiotempconv import bn_apply

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = "cuda"
CASES = [  # name, C, (N, D, H, W), pool, want_full, want_pool, slope
    ("uconv1.inner", 86, (32, 16, 112, 112), (1, 1, 1), True, False, 0.0),
    ("G.dconv1.outer", 32, (32, 16, 112, 112), (2, 2, 2), True, True, 0.2),
    ("S.dconv1.outer", 32, (32, 16, 112, 112), (1, 2, 2), False, True, 0.01),
    ("T.dconv1.outer", 32, (32, 16, 112, 112), (2, 1, 1), False, True, 0.01),
    ("dconv1.inner", 21, (32, 16, 112, 112), (1, 1, 1), True, False, 0.0),
    ("uconv2.inner", 172, (32, 8, 56, 56), (1, 1, 1), True, False, 0.0),
]
if len(sys.argv) > 2:
    CASES = [c for c in CASES if any(s in c[0] for s in sys.argv[2:])]
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


for name, C, (N, D, H, W), pool, wf, wp, slope in CASES:
    Cp = ops.round_up(C, 8)
    bn = torch.nn.BatchNorm3d(C).to(dev).train()
    y = torch.randn(N, D, H, W, Cp, device=dev).bfloat16()
    y[..., C:] = 0
    y.requires_grad_(True)
    prof = P()
    for i in range(reps + 1):
        flush.zero_()
        ops.PROFILER = prof if i else None
        full, pooled = bn_apply(bn, y, slope, pool=pool, want_full=wf, want_pool=wp)
        outs = [t for t in (full, pooled) if t is not None]
        flush.zero_()
        torch.autograd.backward(outs, [torch.ones_like(t) for t in outs])
        y.grad = None
    ops.PROFILER = None
    torch.cuda.synchronize()
    line = f"{name:15s} C={Cp:3d}"
    for k, v in prof.t.items():
        ms = sorted(a.elapsed_time(b) for a, b, _ in v)[len(v) // 2]
        line += f" | {k} {ms:6.3f} ms {v[0][2] / (ms * 1e-3) / 1e9:6.0f} GB/s"
    print(line, flush=True)
    del y, full, pooled, outs
