"""Adjoint identities <conv(x),g> = <x,dgrad(g)> = <w,wgrad(x,g)> on the dominant layer over several seeds.
Prints the normalised residuals (diagnostics for the tolerance in tests/test_parity_gpu.py)."""
import sys
import torch
sys.path.insert(0, ".")
from vfd_gan_b200 import ops

N, D, S, cin, cout = 8, 16, 112, 96, 86
for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    torch.manual_seed(seed)
    x = torch.randn(N, D, S, S, cin, device="cuda").bfloat16()
    w = torch.randn(cout, cin, 1, 3, 3, device="cuda") * 0.05
    gy = torch.randn(N, D, S, S, 88, device="cuda").bfloat16()
    gy[..., cout:] = 0
    xg = x.clone().requires_grad_(True)
    wp = w.clone().requires_grad_(True)
    y = ops.ConvFn.apply(xg, wp, None, True, False)
    y.backward(gy)
    ip_y = float((y.detach().double() * gy.double()).sum())
    tx = xg.grad.double() * x.double()
    tw = wp.grad.double() * w.bfloat16().double()
    # same with an fp32 dgrad output (no bf16 rounding of the result)
    pk = ops._packed(wp)
    gx32 = torch.empty(N, D, S, S, cin, dtype=torch.float32, device="cuda")
    ops.conv3d_fwd(gy, pk.dgrad, None, gx32, None, 1, 3, 3, pk.kc_d, cin, False)
    t32 = gx32.double() * x.double()
    print(f"seed {seed}: ip_y {ip_y:12.3f}  dgrad(bf16) resid/norm {abs(float(tx.sum()) - ip_y) / float(tx.norm()):.2e}  "
          f"dgrad(fp32) {abs(float(t32.sum()) - ip_y) / float(t32.norm()):.2e}  "
          f"wgrad resid/(norm*sqrt(n)) {abs(float(tw.sum()) - ip_y) / (float(tw.norm()) * tw.numel() ** 0.5):.2e}")
