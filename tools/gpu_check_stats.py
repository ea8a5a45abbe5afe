"""Fused conv-epilogue BatchNorm statistics vs the stand-alone vfd_bn_stats kernel."""
import sys
import torch
sys.path.insert(0, ".")
from vfd_gan_b200 import ops

ok = True
for cin, cout, k, N, D, H, W in [(32, 48, (1, 3, 3), 2, 2, 16, 32), (96, 86, (1, 3, 3), 1, 2, 32, 32),
                                 (64, 230, (1, 3, 3), 2, 4, 8, 8), (256, 921, (1, 3, 3), 2, 1, 2, 2),
                                 (837, 1024, (1, 1, 1), 2, 16, 2, 2), (115, 64, (3, 1, 1), 2, 8, 16, 16),
                                 (3, 21, (1, 3, 3), 2, 16, 32, 32), (21, 32, (3, 1, 1), 2, 16, 32, 32),
                                 (3, 2, (1, 1, 1), 2, 16, 32, 32)]:
    x = torch.randn(N, D, H, W, ops.round_up(cin, 8), device="cuda").bfloat16()
    x[..., cin:] = 0
    w = torch.randn(cout, cin, *k, device="cuda") * 0.1
    C = ops.round_up(cout, 8)
    y = ops.ConvFn.apply(x, w, None, False, False, True)
    fused = ops.bn_scratch(x.device, C).clone()
    ops.bn_scratch(x.device, C).zero_()
    ref = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    from vfd_gan_b200 import _lib
    _lib.call("vfd_bn_stats", y.data_ptr(), C, C, N * D * H * W, ref.data_ptr(), torch.cuda.current_stream().cuda_stream)
    yf = y.float().reshape(-1, C)
    exact = torch.cat([yf.double().sum(0), (yf.double() ** 2).sum(0)])
    e1 = float((fused - exact).abs().max() / exact.abs().max())
    e2 = float((ref - exact).abs().max() / exact.abs().max())
    good = e1 < 1e-4 and e2 < 1e-4
    ok &= good
    print(f"cin={cin} cout={cout} k={k} {N}x{D}x{H}x{W}: fused err {e1:.2e} standalone err {e2:.2e} {'ok' if good else 'FAIL'}")
    if not good:
        bad = ((fused - exact).abs() / (exact.abs() + 1e-6)).topk(5)
        print("   worst idx", bad.indices.tolist(), "fused", fused[bad.indices].tolist(), "exact", exact[bad.indices].tolist())
print("ALL OK" if ok else "SOME FAILED")
