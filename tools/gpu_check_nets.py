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
    res = {}
    for rb in (True, False):
        sdo = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
        po = O.netg_forward(sdo, x, True, masks, round_bf16=rb)
        lo = O.weighted_bce(po, gt)
        lo.backward()
        res[rb] = (po.detach(), lo.item(), sdo)
    print(f"[netg] predict rel: cuda-fp32 {rel(pred, res[False][0]):.2e} cuda-matched {rel(pred, res[True][0]):.2e} "
          f"matched-fp32 {rel(res[True][0], res[False][0]):.2e}; loss cuda {loss.item():.6f} fp32 {res[False][1]:.6f}")
    ok = rel(pred, res[True][0]) < 5e-3
    ok &= report_grads(net, res)
    e_rm = max(rel(net.state_dict()[k], res[False][2][k]) for k in sd if "running" in k)
    print(f"   worst running-stat rel vs fp32 oracle {e_rm:.3e}")
    return ok and e_rm < 2e-2


def report_grads(net, res, verbose_over=None):
    """Parameter gradients: the CUDA path must sit inside the bf16 noise envelope, i.e. its distance to
    the fp32 oracle may not exceed twice the distance of the operand-matched oracle to the fp32 one."""
    ok = True
    worst = (0.0, None)
    for k, p_ in net.named_parameters():
        gf, gm = res[False][2][k].grad, res[True][2][k].grad
        if gf is None or (k.endswith(".bias") and ".bn." not in k and "linear" not in k):
            continue  # conv biases in front of a BatchNorm: gradient is identically zero (noise in torch)
        e_cf, e_cm, e_mf = rel(p_.grad, gf), rel(p_.grad, gm), rel(gm, gf)
        good = e_cf <= max(2.0 * e_mf, 2e-2)
        ok &= good
        if not good or e_cf > worst[0]:
            worst = max(worst, (e_cf, k))
        if not good:
            print(f"   grad {k}: cuda-fp32 {e_cf:.2e} cuda-matched {e_cm:.2e} matched-fp32 {e_mf:.2e}  FAIL")
    names = [k for k, _ in net.named_parameters() if k.endswith("weight") and "bn" not in k]
    for k in (names[0], names[len(names) // 2], names[-1]):
        p_ = dict(net.named_parameters())[k]
        print(f"   grad {k}: cuda-fp32 {rel(p_.grad, res[False][2][k].grad):.2e} cuda-matched "
              f"{rel(p_.grad, res[True][2][k].grad):.2e} matched-fp32 {rel(res[True][2][k].grad, res[False][2][k].grad):.2e}")
    print(f"   worst cuda-fp32 param-grad rel {worst[0]:.3e} ({worst[1]}) -> {'ok' if ok else 'FAIL'}")
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
    res = {}
    for rb in (True, False):
        sdo = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
        outs = O.netd_forward(sdo, x, y, True, rb)
        lo = (torch.nn.functional.binary_cross_entropy(outs[0], torch.ones_like(outs[0])) +
              torch.nn.functional.binary_cross_entropy(outs[2], torch.ones_like(outs[2])))
        lo.backward()
        res[rb] = ([o.detach() for o in outs], lo.item(), sdo)
    ok = True
    for name, got, i in (("s_cls", s_cls, 0), ("s_feat", s_feat, 1), ("t_cls", t_cls, 2), ("t_feat", t_feat, 3)):
        e_cf, e_cm, e_mf = rel(got, res[False][0][i]), rel(got, res[True][0][i]), rel(res[True][0][i], res[False][0][i])
        good = e_cf <= max(2.0 * e_mf, 5e-3)
        ok &= good
        print(f"[netd] {name}: cuda-fp32 {e_cf:.2e} cuda-matched {e_cm:.2e} matched-fp32 {e_mf:.2e} {'ok' if good else 'FAIL'}")
    ok &= report_grads(net, res)
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
    which = sys.argv[1:] or ["netd", "lstm", "step", "netg"]
    ok = True
    for w in which:
        t0 = time.time()
        r = {"netg": check_netg, "netd": check_netd, "step": check_step, "lstm": check_lstm}[w]()
        print(f"== {w}: {'OK' if r else 'FAIL'} ({time.time() - t0:.1f}s)", flush=True)
        ok &= r
    sys.exit(0 if ok else 1)
