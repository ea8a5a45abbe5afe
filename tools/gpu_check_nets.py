"""On-device parity of the full networks / train step against the CPU oracle (oracle/vfd_oracle.py).

    python tools/gpu_check_nets.py [netg] [netd] [step] [lstm]
Prints relative errors; the thresholds mirror the pytest -m gpu suite.
"""
import sys
import time
import types
import torch

sys.path.insert(0, ".")
import vfd_gan_b200 as V  # noqa: E402
from vfd_gan_b200 import ops  # noqa: E402
from oracle import vfd_oracle as O  # noqa: E402


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


def dropout_masks(netg, seeds, shapes, p, dev):
    """Recover the Philox masks the CUDA path used: run bn_act_fwd on zeros with scale 0 / shift 1."""
    masks = []
    for seed, (N, C, D, H, W) in zip(seeds, shapes):
        y = torch.zeros(N, D, H, W, C, dtype=torch.bfloat16, device=dev)
        out = torch.empty_like(y)
        ops.bn_act_fwd(y, torch.zeros(C, device=dev), torch.ones(C, device=dev), 1.0, out, None, 1, 1, 1, p, seed)
        masks.append(out.float().permute(0, 4, 1, 2, 3).cpu().contiguous())
    return masks


def check_netg(B=2, D=16, S=32, ngf=32):
    dev = "cuda"
    torch.manual_seed(0)
    net = V.NetG(3, ngf)
    net.apply(V.weights_init)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.to(dev).train()
    x = torch.rand(B, 3, D, S, S) * 2 - 1
    gt = (torch.rand(B, 1, D, S, S) > 0.9).float()
    pred = net(x.to(dev))
    loss = V.weighted_bce(pred, gt.to(dev))
    loss.backward()
    seeds = net.last_dropout_seeds
    g = ngf
    shapes = [(B, 8 * g, D // 16, S // 16, S // 16), (B, 8 * g, D // 8, S // 8, S // 8), (B, 4 * g, D // 4, S // 4, S // 4),
              (B, 2 * g, D // 2, S // 2, S // 2)]
    masks = dropout_masks(net, seeds, shapes, 0.25, dev)
    print("dropout keep fractions", [round(float((m > 0).float().mean()), 4) for m in masks])
    ok = True
    for rb in (True, False):
        sdo = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
        po = O.netg_forward(sdo, x, True, masks, round_bf16=rb)
        lo = O.weighted_bce(po, gt)
        lo.backward()
        e_p = rel(pred, po)
        print(f"[netg oracle round_bf16={rb}] predict rel {e_p:.3e}  loss cuda {loss.item():.6f} oracle {lo.item():.6f}")
        worst = 0.0
        for k, p_ in net.named_parameters():
            if sdo[k].grad is None or k.endswith("conv.bias") or ".bias" in k and "bn" not in k:
                continue
            e = rel(p_.grad, sdo[k].grad)
            worst = max(worst, e)
            if e > 2e-2:
                print(f"   grad {k}: rel {e:.3e}")
        print(f"   worst param-grad rel {worst:.3e}")
        e_rm = max(rel(net.state_dict()[k], sdo[k]) for k in sd if "running" in k)
        print(f"   worst running-stat rel {e_rm:.3e}")
        if rb:
            ok &= e_p < 5e-3 and worst < 5e-2
    return ok


def check_netd(B=2, D=16, S=64):
    dev = "cuda"
    args = types.SimpleNamespace(nfr=D, isize=S)
    torch.manual_seed(1)
    net = V.NetD(args)
    net.apply(V.weights_init)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.to(dev).train()
    x = torch.rand(B, 3, D, S, S)
    y = torch.rand(B, 3, D, S, S) * 2 - 1
    s_cls, s_feat, t_cls, t_feat = net(x.to(dev), y.to(dev))
    loss = (torch.nn.functional.binary_cross_entropy(s_cls, torch.ones_like(s_cls)) +
            torch.nn.functional.binary_cross_entropy(t_cls, torch.ones_like(t_cls)))
    loss.backward()
    ok = True
    for rb in (True, False):
        sdo = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
        so, sfo, to, tfo = O.netd_forward(sdo, x, y, True, rb)
        lo = (torch.nn.functional.binary_cross_entropy(so, torch.ones_like(so)) +
              torch.nn.functional.binary_cross_entropy(to, torch.ones_like(to)))
        lo.backward()
        print(f"[netd oracle round_bf16={rb}] s_cls {rel(s_cls, so):.3e} t_cls {rel(t_cls, to):.3e} "
              f"s_feat {rel(s_feat, sfo):.3e} t_feat {rel(t_feat, tfo):.3e}")
        worst = 0.0
        for k, p_ in net.named_parameters():
            if sdo[k].grad is None or (k.endswith(".bias") and "bn" not in k and "linear" not in k):
                continue
            e = rel(p_.grad, sdo[k].grad)
            worst = max(worst, e)
            if e > 5e-2:
                print(f"   grad {k}: rel {e:.3e}")
        print(f"   worst param-grad rel {worst:.3e}")
        if rb:
            ok &= rel(s_feat, sfo) < 1e-2 and rel(t_feat, tfo) < 1e-2 and worst < 1e-1
    return ok


def check_step(B=4, D=16, S=64, steps=10):
    dev = "cuda"
    args = types.SimpleNamespace(nfr=D, isize=S)
    torch.manual_seed(0)
    netg = V.NetG()
    netd = V.NetD(args)
    netg.apply(V.weights_init)
    netd.apply(V.weights_init)
    netg.dropout.p = 0.0  # trajectory parity is checked without dropout
    oracle = O.OracleTrainer(netg.state_dict(), netd.state_dict())
    netg, netd = netg.to(dev), netd.to(dev)
    tr = V.GanTrainStep(netg, netd)
    ok = True
    t_cpu = 0.0
    for it in range(steps):
        inp, gt, gf, pf = O.synthetic_batch(B, D, S, seed=100 + it)
        tr.step(inp.to(dev), gt.to(dev), gf.to(dev), pf.to(dev))
        got = tr.losses_dict()
        t0 = time.time()
        want, _ = oracle.step(inp, gt, gf, pf, dropout_masks=[1.0, 1.0, 1.0, 1.0])
        t_cpu += time.time() - t0
        errs = {k: abs(got[k] - want[k]) / (abs(want[k]) + 1e-12) for k in want}
        worst = max(errs, key=errs.get)
        print(f"step {it}: err_g {got['g/err_g']:.5f}/{want['g/err_g']:.5f} err_d {got['d/err_d']:.5f}/"
              f"{want['d/err_d']:.5f} adv {got['g/err_g_adv']:.5f}/{want['g/err_g_adv']:.5f} worst {worst} "
              f"{errs[worst]:.2e}", flush=True)
        if it == steps - 1:
            ok &= errs[worst] < 1e-2
    print(f"oracle CPU time/step {t_cpu / steps:.2f}s")
    return ok


def check_lstm():
    dev = "cuda"
    torch.manual_seed(3)
    cell = V.ConvLSTMCell((8, 8), 16, 32, (3, 3), False)
    sd = {("" + k): v.clone() for k, v in cell.state_dict().items()}
    cell = cell.to(dev)
    x = torch.randn(2, 16, 8, 8)
    h = torch.randn(2, 32, 8, 8) * 0.5
    c = torch.randn(2, 32, 8, 8) * 0.5
    xg, hg, cg = (t.to(dev).requires_grad_(True) for t in (x, h, c))
    hn, cn = cell(xg, (hg, cg))
    (hn.sum() + (cn * cn).sum()).backward()
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xo, ho, co = (t.clone().requires_grad_(True) for t in (x, h, c))
    hno, cno = O.convlstm_cell(sdo, "", xo, ho, co, round_bf16=True)
    (hno.sum() + (cno * cno).sum()).backward()
    e = [rel(hn, hno), rel(cn, cno), rel(xg.grad, xo.grad), rel(hg.grad, ho.grad), rel(cg.grad, co.grad),
         rel(cell.conv.weight.grad, sdo["conv.weight"].grad)]
    print("[lstm] h c dx dh dc dW rel:", " ".join(f"{v:.2e}" for v in e))
    return max(e[:2]) < 1e-3 and max(e[2:]) < 2e-2


if __name__ == "__main__":
    which = sys.argv[1:] or ["netg", "netd", "lstm", "step"]
    ok = True
    for w in which:
        t0 = time.time()
        r = {"netg": check_netg, "netd": check_netd, "step": check_step, "lstm": check_lstm}[w]()
        print(f"== {w}: {'OK' if r else 'FAIL'} ({time.time() - t0:.1f}s)", flush=True)
        ok &= r
    sys.exit(0 if ok else 1)
