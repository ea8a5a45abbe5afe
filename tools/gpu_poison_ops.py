"""Op-level uninitialised-read detector (compute-sanitizer is not available on the pool): every autograd function
of vfd_gan_b200.ops runs forward + backward on small, ragged shapes with the caching allocator's free blocks
pre-filled with 0 (twice: the run-to-run noise of the atomics-based kernels), NaN and +Inf bit patterns. Outputs
that turn non-finite or move by more than the noise mean some kernel consumed memory nobody wrote.
    python tools/gpu_poison_ops.py"""
import sys
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import vfd_gan_b200 as V
from vfd_gan_b200 import ops
from vfd_gan_b200.convlstm import ConvLSTMCell

DEV = "cuda"


def poison(value):
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    t = torch.empty(int(6e9) // 4, dtype=torch.int32, device=DEV)
    t.fill_(value)
    del t


def cl(shape, C, seed, grad=False):
    g = torch.Generator().manual_seed(seed)
    Cp = ops.round_up(C, 8)
    t = torch.randn(*shape, Cp, generator=g)
    t[..., C:] = 0
    t = t.to(DEV).bfloat16()
    return t.requires_grad_(True) if grad else t


def conv_case(cin, cout, k, N, D, H, W, fuse=False, out_fp32=False, bias=False):
    def run():
        x = cl((N, D, H, W), cin, 1, True)
        g = torch.Generator().manual_seed(2)
        w = (torch.randn(cout, cin, *k, generator=g) * 0.1).to(DEV).requires_grad_(True)
        b = torch.randn(cout, generator=g).to(DEV).requires_grad_(True) if bias else None
        y = ops.ConvFn.apply(x, w, b, out_fp32, False, fuse)
        res = {"y": y.detach().float().clone()}
        if fuse:
            sc = ops.bn_scratch(x.device, y.shape[-1])
            res["stats"] = sc.clone().float()
            sc.zero_()
        gy = cl((N, D, H, W), cout, 3)
        y.backward(gy.float() if out_fp32 else gy)
        res["gx"], res["gw"] = x.grad.float().clone(), w.grad.clone()
        if bias:
            res["gb"] = b.grad.clone()
        return res
    return run


def bn_case(C, slope, pool, shape, drop=0.0, want_full=True, want_pool=True):
    def run():
        y = cl(shape, C, 4, True)
        bn = torch.nn.BatchNorm3d(C).to(DEV).train()
        from vfd_gan_b200.spatiotempconv import bn_apply
        full, pooled = bn_apply(bn, y, slope, pool=pool, drop_p=drop, seed=7, want_full=want_full, want_pool=want_pool)
        res, loss = {}, 0
        if full is not None:
            res["full"] = full.detach().float().clone()
            loss = loss + (full.float() * cl(tuple(full.shape[:-1]), C, 5).float()).sum()
        if pooled is not None:
            res["pooled"] = pooled.detach().float().clone()
            loss = loss + (pooled.float() * cl(tuple(pooled.shape[:-1]), C, 6).float()).sum()
        loss.backward()
        res["gy"], res["gg"], res["gb"] = y.grad.float().clone(), bn.weight.grad.clone(), bn.bias.grad.clone()
        res["rm"] = bn.running_mean.clone()
        return res
    return run


def up_case(shape, C, cat):
    def run():
        low = cl(shape, C, 8, True)
        N, D, H, W = shape
        if cat:
            buf = torch.empty(N, 2 * D, 2 * H, 2 * W, 2 * ops.round_up(C, 8), dtype=torch.bfloat16, device=DEV)
            skip = cl((N, 2 * D, 2 * H, 2 * W), C, 9)
            buf[..., ops.round_up(C, 8):] = skip
            out = ops.UpCatFn.apply(low, skip, [buf])
        else:
            out = ops.UpsampleFn.apply(low)
        (out.float() * cl(tuple(out.shape[:-1]), out.shape[-1], 10).float()).sum().backward()
        return {"out": out.detach().float().clone(), "g": low.grad.float().clone()}
    return run


def idpool_case(shape, C, pool, drop):
    def run():
        x = cl(shape, C, 11, True)
        out = ops.IdentityPoolFn.apply(x, pool, drop, 5)
        (out.float() * cl(tuple(out.shape[:-1]), C, 12).float()).sum().backward()
        return {"out": out.detach().float().clone(), "g": x.grad.float().clone()}
    return run


def lstm_case():
    torch.manual_seed(3)
    cell = ConvLSTMCell((5, 6), 16, 24, (3, 3), True).to(DEV)
    x, h, c = (torch.randn(2, ch, 5, 6, device=DEV) for ch in (16, 24, 24))
    hn, cn = cell(x, (h, c))
    (hn.sum() + 2 * cn.sum()).backward()
    return {"h": hn.detach().clone(), "c": cn.detach().clone(), "gw": cell.conv.weight.grad.clone()}


def misc_case():
    g = torch.Generator().manual_seed(13)
    a = torch.randn(3, 5, 2, 7, 9, generator=g).to(DEV)
    ac = ops.PackFn.apply(a, 0)
    back = ops.UnpackFn.apply(ac, 5)
    one = torch.rand(3, 1, 2, 7, 9, generator=g).to(DEV)
    rep = ops.PackFn.apply(one, 3)
    p = one.clone().requires_grad_(True)
    t = (torch.rand(3, 1, 2, 7, 9, generator=g) > 0.5).float().to(DEV)
    l = V.weighted_bce(p, t)
    l.backward()
    tm = V.evaluate.threshold_open(one)
    return {"pack": ac.float().clone(), "unpack": back.clone(), "rep": rep.float().clone(), "wbce": l.detach().reshape(1),
            "gp": p.grad.clone(), "t": tm[0], "m": tm[1]}


CASES = {
    "conv 8->16 1x3x3": conv_case(8, 16, (1, 3, 3), 1, 2, 16, 16, fuse=True),
    "conv 3->21 1x3x3 ragged": conv_case(3, 21, (1, 3, 3), 2, 3, 13, 11, fuse=True),
    "conv 21->32 3x1x1 ragged": conv_case(21, 32, (3, 1, 1), 2, 3, 13, 11, fuse=True),
    "conv 96->86 1x3x3": conv_case(96, 86, (1, 3, 3), 1, 2, 20, 12, fuse=True),
    "conv 86->32 3x1x1": conv_case(86, 32, (3, 1, 1), 1, 5, 12, 20, fuse=True),
    "conv 32->1 3x3x3 fp32": conv_case(32, 1, (3, 3, 3), 1, 4, 9, 10, out_fp32=True),
    "conv 2->32 3x1x1": conv_case(2, 32, (3, 1, 1), 1, 8, 12, 20, fuse=True),
    "conv 14->32 1x1x1": conv_case(14, 32, (1, 1, 1), 3, 5, 9, 11, fuse=True),
    "conv 3->2 1x1x1 tiny": conv_case(3, 2, (1, 1, 1), 3, 5, 9, 11, fuse=True),
    "conv 3->64 1x1x1 bias": conv_case(3, 64, (1, 1, 1), 2, 4, 8, 8, bias=True),
    "conv 128->64 3x3x3": conv_case(128, 64, (3, 3, 3), 1, 2, 7, 7),
    "conv 256->600 1x3x3": conv_case(256, 600, (1, 3, 3), 2, 1, 6, 6, fuse=True),
    "conv 512->921 1x1x1": conv_case(512, 921, (1, 1, 1), 2, 3, 2, 2, fuse=True),
    "conv2d 40->96 3x3 fp32 bias": conv_case(40, 96, (1, 3, 3), 2, 1, 5, 6, out_fp32=True, bias=True),
    "bn 24 pool222": bn_case(24, 0.2, (2, 2, 2), (2, 4, 8, 8)),
    "bn 8 pool122": bn_case(8, 0.01, (1, 2, 2), (3, 2, 6, 6), want_full=False),
    "bn 32 pool211": bn_case(32, 0.01, (2, 1, 1), (1, 4, 5, 7), want_full=False),
    "bn 21 relu": bn_case(21, 0.0, (1, 1, 1), (2, 3, 5, 7), want_pool=False),
    "bn 64 drop": bn_case(64, 0.2, (1, 1, 1), (2, 2, 4, 4), drop=0.25, want_pool=False),
    "bn 300 pool222": bn_case(300, 0.2, (2, 2, 2), (2, 2, 4, 4)),
    "upcat": up_case((2, 1, 3, 5), 16, True),
    "upsample": up_case((1, 2, 7, 7), 24, False),
    "idpool 222": idpool_case((2, 2, 6, 6), 64, (2, 2, 2), 0.0),
    "idpool drop": idpool_case((2, 1, 3, 3), 512, (1, 1, 1), 0.25),
    "convlstm cell": lstm_case,
    "misc": misc_case,
}

bad = 0
filt = [a for a in sys.argv[1:] if not a.startswith("x")]
reps = max([int(a[1:]) for a in sys.argv[1:] if a.startswith("x")] + [1])
todo = [(n, f) for n, f in CASES.items() if not filt or any(s in n for s in filt)] * reps
for name, fn in todo:
    outs = []
    for pat in (0, 0, 0x7FC07FC0, 0x7F807F80):
        poison(pat)
        outs.append(fn())
        torch.cuda.synchronize()
    ref, again, nan_run, inf_run = outs
    msgs = []
    for k in ref:
        noise = float((again[k].double() - ref[k].double()).norm() / (ref[k].double().norm() + 1e-30))
        for tag, other in (("nan", nan_run), ("inf", inf_run)):
            if not torch.isfinite(other[k]).all():
                msgs.append(f"{k} not finite under {tag}-poison")
            else:
                d = float((other[k].double() - ref[k].double()).norm() / (ref[k].double().norm() + 1e-30))
                if d > 4 * noise + 5e-4:   # a lone bf16-ulp flip (fp64 atomics order) stays below this
                    diff = (other[k].double() - ref[k].double()).abs().flatten()
                    nz = diff.nonzero().flatten()
                    msgs.append(f"{k} moves {d:.2e} under {tag}-poison (noise {noise:.1e}; {nz.numel()} of "
                                f"{diff.numel()} elements, first idx {nz[:6].tolist()}, max {float(diff.max()):.3e}, "
                                f"ref there {float(ref[k].flatten()[diff.argmax()]):.4e})")
    bad += len(msgs)
    print(f"{name:32s}", "clean" if not msgs else "; ".join(msgs), flush=True)
print("POISON OPS", "CLEAN" if bad == 0 else f"{bad} PROBLEMS")
