"""On-device cross-check of the tcgen05 conv kernels (forward, dgrad, wgrad) against
(a) the CUDA-core direct kernels on identical packed operands and (b) torch conv3d in fp32 on the
same bf16-rounded operands. Prints one line per case; exit code 1 if any case fails.

    python tools/gpu_check_conv.py [--quick]
"""
import sys
import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from vfd_gan_b200 import ops  # noqa: E402

CASES = [
    # cin, cout, kernel, N, D, H, W
    (8, 16, (1, 3, 3), 1, 2, 16, 16),
    (3, 21, (1, 3, 3), 2, 4, 16, 16),
    (32, 48, (1, 3, 3), 2, 2, 16, 32),
    (64, 64, (3, 1, 1), 1, 4, 16, 16),
    (96, 86, (1, 3, 3), 1, 2, 32, 32),
    (128, 300, (1, 1, 1), 1, 2, 16, 16),
    (32, 1, (3, 3, 3), 1, 4, 16, 16),
    (24, 40, (1, 3, 3), 16, 1, 7, 7),
    (2, 32, (3, 1, 1), 1, 8, 12, 20),
    (256, 72, (3, 3, 3), 1, 2, 8, 8),
    (512, 658, (1, 3, 3), 2, 1, 8, 8),
]


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-12))


def run_case(cin, cout, k, N, D, H, W, verbose=True):
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(cin * 1000 + cout)
    x = torch.randn(N, cin, D, H, W, generator=g).to(dev)
    w = (torch.randn(cout, cin, *k, generator=g) * 0.1).to(dev)
    b = torch.randn(cout, generator=g).to(dev)
    gy = torch.randn(N, cout, D, H, W, generator=g).to(dev)
    pad = tuple(kk // 2 for kk in k)

    # torch fp32 reference on bf16-rounded operands
    xr = x.bfloat16().float().requires_grad_(True)
    wr = w.bfloat16().float().requires_grad_(True)
    gyr = gy.bfloat16().float()
    yr = F.conv3d(xr, wr, b, padding=pad)
    yr.backward(gyr)

    res = {}
    for impl in ("tc", "direct"):
        ops.CONV_IMPL_DIRECT = impl == "direct"
        xc = ops.PackFn.apply(x, 0).requires_grad_(True)
        wp = w.clone().requires_grad_(True)
        bp = b.clone().requires_grad_(True)
        yc = ops.ConvFn.apply(xc, wp, bp, True, False)
        y = ops.UnpackFn.apply(yc, cout)
        gyc = ops.PackFn.apply(gy, 0)
        yc.backward(gyc.float() if False else gyc)
        gx = torch.empty(N, cin, D, H, W, device=dev)
        ops.unpack_ncdhw(xc.grad, gx)
        res[impl] = (y.detach(), gx, wp.grad, bp.grad)
    ops.CONV_IMPL_DIRECT = False
    torch.cuda.synchronize()
    out = []
    ok = True
    for name, idx, ref, tol in (("fwd", 0, yr.detach(), 2e-5), ("dgrad", 1, xr.grad, 5e-3), ("wgrad", 2, wr.grad, 5e-3),
                                ("bgrad", 3, gyr.sum((0, 2, 3, 4)), 5e-3)):
        e_tc = rel(res["tc"][idx], ref)
        e_di = rel(res["direct"][idx], ref)
        e_x = rel(res["tc"][idx], res["direct"][idx])
        good = e_tc < tol and e_x < max(tol, 1e-5)
        ok &= good
        out.append(f"{name}: tc-ref {e_tc:.2e} direct-ref {e_di:.2e} tc-direct {e_x:.2e} {'ok' if good else 'FAIL'}")
    if verbose:
        print(f"cin={cin} cout={cout} k={k} N={N} D={D} H={H} W={W} :: " + " | ".join(out), flush=True)
    return ok


if __name__ == "__main__":
    cases = CASES[:4] if "--quick" in sys.argv else CASES
    allok = True
    for c in cases:
        try:
            allok &= run_case(*c)
        except Exception as e:  # keep going: one broken configuration should not hide the others
            allok = False
            print(f"case {c}: EXCEPTION {type(e).__name__}: {e}", flush=True)
            if "CUDA error" in str(e) or "launch failure" in str(e) or "illegal" in str(e):
                break
    print("ALL OK" if allok else "SOME FAILED")
    sys.exit(0 if allok else 1)
