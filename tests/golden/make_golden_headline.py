"""Generates the fixtures that pin the HEADLINE configuration (BASELINE configs[1]: 16x3x112x112 clips) against the
reference's own modules, imported from /root/reference in the build container.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_headline.py

  headline_traj_112.pt   3 optimize_params steps (models/mygannet.py:350-366 restated around the reference's NetG /
                         NetD, dropout off) at B=2, 16x3x112x112: the 12 logged losses per step, and for step 0 the
                         prediction (bf16), every BatchNorm weight/bias gradient, the conv_last / first / dominant
                         (uconv1) conv weight gradients and the Frobenius norm of every other parameter gradient,
                         for NetG (from err_g) and NetD (from err_d).
  headline_step1_b32.pt  the 12 losses of the FIRST step of exactly bench.py's configuration (B=32, weights from
                         torch.manual_seed(0), data from torch.Generator().manual_seed(1), dropout off), forward only.

112 is not a size the reference's NetD accepts (its Linear layers hard-code the 128 / 16 geometry), so
both Linears and TDisc's global pool are re-created for 112 exactly as vfd_gan_b200.NetD(isize=112) sizes them; every
conv / BatchNorm is the reference's.
"""
import os
import sys
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import make_ref  # noqa: E402
from oracle.vfd_oracle import synthetic_batch  # noqa: E402  (only the seeded input generator)

R = make_ref.import_ref("/root/reference")
mg, lu = R.mygannet, R.utils
D, S = 16, 112
KEYS = ("g/err_g", "g/err_g_adv", "g/err_g_adv_s", "g/err_g_adv_t", "g/err_g_con", "d/err_d_real_s", "d/err_d_real_t",
        "d/err_d_fake_s", "d/err_d_fake_t", "d/err_d_real", "d/err_d_fake", "d/err_d")
FULL_GRADS = ("conv_last.weight", "dconv1.conv.spatial_conv.weight", "uconv1.conv.spatial_conv.weight",
              "uconv1.conv.temporal_conv.weight", "spatdisc.dconv1.conv.spatial_conv.weight",
              "tempdisc.dconv1.conv.temporal_conv.weight", "spatdisc.linear.weight", "tempdisc.linear.weight")


def reference_nets_112():
    """Same RNG call sequence as tests/helpers.py::build_headline_nets."""
    torch.manual_seed(0)
    netg = mg.NetG()
    netd = mg.NetD(types.SimpleNamespace(nfr=D, isize=128))
    netd.tempdisc.gpool = nn.AvgPool3d((1, S, S), stride=1)
    netd.spatdisc.linear = nn.Linear(32 * 32 * (S // 64) ** 2, 1)
    netd.tempdisc.linear = nn.Linear(32 * 4 * (D // 8), 1)
    netg.apply(lu.weights_init)
    netd.apply(lu.weights_init)
    netg.dropout.p = 0.0
    return netg.train(), netd.train()


def one_step(netg, netd, opt_g, opt_d, batch, backward=True):
    inp, gt, gt_flow, pre_flow = batch
    B = inp.shape[0]
    bce = nn.BCELoss()
    predict = netg(inp)
    pre_3ch, gt_3ch = lu.gray2rgb(predict.detach()), lu.gray2rgb(gt.detach())
    s_pr, s_fr, t_pr, t_fr = netd(gt_3ch, gt_flow.detach())
    s_pf, s_ff, t_pf, t_ff = netd(pre_3ch.detach(), pre_flow.detach())
    adv_s, adv_t = lu.l2_loss(s_fr, s_ff), lu.l2_loss(t_fr, t_ff)
    adv = adv_s + adv_t
    con = lu.weighted_bce(predict, gt)
    err_g = adv * 1 + con * 10
    ones, zeros = torch.ones(B), torch.zeros(B)
    e_rs, e_rt, e_fs, e_ft = bce(s_pr, ones), bce(t_pr, ones), bce(s_pf, zeros), bce(t_pf, zeros)
    real, fake = (e_rs + e_rt) * 0.5, (e_fs + e_ft) * 0.5
    err_d = (real + fake) * 0.5
    grads = None
    if backward:
        opt_g.zero_grad()
        err_g.backward(retain_graph=True)
        g_grads = {k: p.grad.detach().clone() for k, p in netg.named_parameters()}
        opt_g.step()
        opt_d.zero_grad()
        err_d.backward()
        d_grads = {k: p.grad.detach().clone() for k, p in netd.named_parameters()}
        opt_d.step()
        grads = (g_grads, d_grads)
    vals = [err_g, adv, adv_s, adv_t, con, e_rs, e_rt, e_fs, e_ft, real, fake, err_d]
    return dict(zip(KEYS, (float(v) for v in vals))), predict.detach(), grads


def pack_grads(grads):
    out = {"full": {}, "norm": {}}
    for k, g in grads.items():
        if k in FULL_GRADS or ".bn." in k:
            out["full"][k] = g
        out["norm"][k] = float(g.norm())
    return out


def trajectory():
    B = 2
    netg, netd = reference_nets_112()
    init_check = {"g_first": netg.dconv1.conv.spatial_conv.weight.detach().flatten()[:8].clone(),
                  "d_lin": netd.spatdisc.linear.weight.detach().flatten()[:8].clone(),
                  "t_lin": netd.tempdisc.linear.weight.detach().flatten()[:8].clone()}
    opt_d = torch.optim.Adam(netd.parameters(), lr=2e-5, betas=(0.5, 0.999))
    opt_g = torch.optim.Adam(netg.parameters(), lr=2e-5, betas=(0.5, 0.999))
    traj, step0 = [], None
    for it in range(3):
        losses, predict, (gg, gd) = one_step(netg, netd, opt_g, opt_d, synthetic_batch(B, D, S, seed=200 + it))
        traj.append(losses)
        if it == 0:
            step0 = {"predict": predict.bfloat16(), "g": pack_grads(gg), "d": pack_grads(gd)}
        print(it, losses["g/err_g"], losses["d/err_d"], flush=True)
    torch.save({"traj": traj, "step0": step0, "init_check": init_check,
                "config": {"B": B, "D": D, "S": S, "seed": 0, "data_seed0": 200}},
               os.path.join(HERE, "headline_traj_112.pt"))


def bench_step1():
    """bench.py's nets (vfd_gan_b200 constructors are the reference's, bit-equal init: tests/test_reference_surface.py)
    loaded into the reference modules; bench.py's rank-0 data."""
    import vfd_gan_b200 as V
    B = 32
    torch.manual_seed(0)
    vg, vd = V.NetG(), V.NetD(types.SimpleNamespace(nfr=D, isize=S))
    vg.apply(V.weights_init)
    vd.apply(V.weights_init)
    netg, netd = reference_nets_112()
    netg.load_state_dict(vg.state_dict())
    netd.load_state_dict(vd.state_dict())
    g = torch.Generator().manual_seed(1)
    shp3, shp1 = (B, 3, D, S, S), (B, 1, D, S, S)
    inp = torch.rand(shp3, generator=g) * 2 - 1
    gt = (torch.rand(shp1, generator=g) > 0.9).float()
    gf = torch.rand(shp3, generator=g) * 2 - 1
    pf = torch.rand(shp3, generator=g) * 2 - 1
    with torch.no_grad():
        losses, predict, _ = one_step(netg, netd, None, None, (inp, gt, gf, pf), backward=False)
    print("bench step 1:", losses, flush=True)
    torch.save({"losses": losses, "predict_clip_means": predict.mean(dim=(1, 2, 3, 4)),
                "config": {"B": B, "D": D, "S": S, "weights_seed": 0, "data_seed": 1, "dropout": 0.0}},
               os.path.join(HERE, "headline_step1_b32.pt"))


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    trajectory()
    bench_step1()
