"""Generates tests/golden/*.pt by running the REFERENCE's own modules (imported from /root/reference,
build container only) on seeded inputs. The fixtures pin the oracle (tests/test_oracle.py) and the
CUDA path (tests/test_parity_gpu.py) on machines where the reference is absent.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Fixtures (small on purpose; weights are re-created from the seed, not stored, wherever the nets are big):
  st_conv_small.pt    SpatioTemporalConv(8->16,k3) state_dict + input + train-mode output + input/weight grads
  convlstm_cell.pt    ConvLSTMCell(16,32,(3,3)) state_dict + inputs + (h,c)
  losses.pt           l2_loss / weighted_bce / BCELoss values on seeded tensors
  netg_netd_small.pt  NetG(ngf=8) + SDisc/TDisc(ndf=8) outputs (train mode, dropout off); weights/inputs from seed 14
  step_traj_cfg1.pt   12 logged losses over 10 optimize_params steps of the full-size NetG/NetD at
                      BASELINE config 1 (B=4, 16x3x64x64; weights from torch.manual_seed(0) + weights_init,
                      NetD Linears sized for isize=64 as in SURVEY.md D4), dropout disabled.
"""
import os
import sys
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
for n in ("matplotlib", "matplotlib.pyplot", "skimage", "skimage.transform"):
    sys.modules.setdefault(n, types.ModuleType(n))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.modules["matplotlib"].rc = lambda *a, **k: None
sys.modules["matplotlib.pyplot"].figure = lambda *a, **k: None
sys.modules["skimage"].transform = sys.modules["skimage.transform"]
sys.modules["skimage.transform"].resize = None
sys.path.insert(0, "/root/reference")

from models.spatiotempconv import SpatioTemporalConv  # noqa: E402
from models.convlstm import ConvLSTMCell  # noqa: E402
from models.mygannet import NetG, NetD, SDisc, TDisc  # noqa: E402
from lib.utils import weights_init, l2_loss, weighted_bce, gray2rgb  # noqa: E402
from oracle.vfd_oracle import synthetic_batch  # noqa: E402  (only the seeded input generator)


def sd_clone(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def main():
    # ---- SpatioTemporalConv
    torch.manual_seed(11)
    m = SpatioTemporalConv(8, 16, 3, padding=1)
    m.apply(weights_init)
    sd0 = sd_clone(m)
    x = torch.randn(2, 8, 4, 12, 12, requires_grad=True)
    gy = torch.randn(2, 16, 4, 12, 12)
    m.train()
    y = m(x)
    y.backward(gy)
    torch.save({"sd": sd0, "x": x.detach(), "gy": gy, "y": y.detach(), "gx": x.grad,
                "gw_spatial": m.spatial_conv.weight.grad, "gw_temporal": m.temporal_conv.weight.grad,
                "gb_temporal": m.temporal_conv.bias.grad, "g_bn_w": m.bn.weight.grad, "g_bn_b": m.bn.bias.grad,
                "sd_after": sd_clone(m)}, os.path.join(HERE, "st_conv_small.pt"))

    # ---- ConvLSTMCell
    torch.manual_seed(12)
    cell = ConvLSTMCell((8, 8), 16, 32, (3, 3), True)
    xi, h, c = torch.randn(2, 16, 8, 8), torch.randn(2, 32, 8, 8) * 0.5, torch.randn(2, 32, 8, 8) * 0.5
    hn, cn = cell(xi, (h, c))
    torch.save({"sd": sd_clone(cell), "x": xi, "h": h, "c": c, "h_next": hn.detach(), "c_next": cn.detach()},
               os.path.join(HERE, "convlstm_cell.pt"))

    # ---- losses
    torch.manual_seed(13)
    a, b = torch.randn(3, 5, 4, 6, 6), torch.randn(3, 5, 4, 6, 6)
    p = torch.rand(3, 1, 4, 6, 6)
    p[0, 0, 0, 0, 0], p[0, 0, 0, 0, 1] = 0.0, 1e-9
    t = (torch.rand(3, 1, 4, 6, 6) > 0.7).float()
    t[0, 0, 0, 0, 0] = 1.0
    q = torch.rand(7)
    torch.save({"a": a, "b": b, "l2": l2_loss(a, b), "p": p, "t": t, "wbce": weighted_bce(p, t),
                "wbce_pw3": weighted_bce(p, t, pos_weight=3), "q": q,
                "bce_ones": nn.BCELoss()(q, torch.ones(7)), "bce_zeros": nn.BCELoss()(q, torch.zeros(7))},
               os.path.join(HERE, "losses.pt"))

    # ---- small nets (train mode, dropout off). Weights and inputs are NOT stored: the test re-creates
    #      them with the same torch.manual_seed(14) call sequence (module construction order and init
    #      RNG consumption of vfd_gan_b200's modules equal the reference's; tests/test_reference_surface.py).
    torch.manual_seed(14)
    g = NetG(3, 8)
    g.apply(weights_init)
    g.dropout.p = 0.0
    g.train()
    xg = torch.rand(2, 3, 16, 32, 32) * 2 - 1
    sdisc = SDisc(3, 16, ndf=8, kernel=(1, 3, 3), padding=(0, 1, 1))      # needs isize 128 (2x2 final map)
    tdisc = TDisc(3, 32, ndf=8, kernel=(3, 1, 1), padding=(1, 0, 0))      # needs nfr 16
    sdisc.apply(weights_init)
    tdisc.apply(weights_init)
    xs = torch.rand(1, 3, 16, 128, 128)
    xt = torch.rand(2, 3, 16, 32, 32) * 2 - 1
    init_check = {"g": g.dconv1.conv.spatial_conv.weight.detach().flatten()[:8].clone(),
                  "s": sdisc.dconv1.conv.spatial_conv.weight.detach().flatten()[:8].clone(),
                  "t": tdisc.linear.weight.detach().flatten()[:8].clone(), "xg": xg.flatten()[:8].clone(),
                  "xs": xs.flatten()[:8].clone(), "xt": xt.flatten()[:8].clone()}
    pred = g(xg)
    s_cls, s_feat = sdisc(xs)
    t_cls, t_feat = tdisc(xt)
    torch.save({"init_check": init_check, "predict": pred.detach(),
                "g_running": {k: v.clone() for k, v in g.state_dict().items() if k.startswith("uconv1.") and "running" in k},
                "s_cls": s_cls.detach(), "s_feat": s_feat.detach(), "t_cls": t_cls.detach(),
                "t_feat": t_feat.detach()}, os.path.join(HERE, "netg_netd_small.pt"))

    # ---- 10-step trajectory of the full-size nets at config 1 (optimize_params restated around the
    #      reference modules: MyGAN itself hard-codes 'cuda', models/mygannet.py:239-261)
    B, D, S = 4, 16, 64
    torch.manual_seed(0)
    netg = NetG()
    args = types.SimpleNamespace(nfr=D, isize=S)
    netd = NetD(args)
    netd.spatdisc.linear = nn.Linear(32 * 32 * (S // 64) ** 2, 1)   # generalised Linear sizes;
    netd.tempdisc.linear = nn.Linear(32 * 4 * (D // 8), 1)          # identical to the reference at 128/16
    netg.apply(weights_init)
    netd.apply(weights_init)
    netg.dropout.p = 0.0
    netg.train()
    netd.train()
    init_check = {"g_first": netg.dconv1.conv.spatial_conv.weight.detach().flatten()[:8].clone(),
                  "d_first": netd.spatdisc.dconv1.conv.spatial_conv.weight.detach().flatten()[:8].clone(),
                  "d_lin": netd.spatdisc.linear.weight.detach().flatten()[:8].clone()}
    opt_d = torch.optim.Adam(netd.parameters(), lr=2e-5, betas=(0.5, 0.999))
    opt_g = torch.optim.Adam(netg.parameters(), lr=2e-5, betas=(0.5, 0.999))
    bce = nn.BCELoss()
    ones, zeros = torch.ones(B), torch.zeros(B)
    traj = []
    for it in range(10):
        inp, gt, gt_flow, pre_flow = synthetic_batch(B, D, S, seed=100 + it)
        predict = netg(inp)
        pre_3ch, gt_3ch = gray2rgb(predict.detach()), gray2rgb(gt.detach())
        s_pr, s_fr, t_pr, t_fr = netd(gt_3ch, gt_flow.detach())
        s_pf, s_ff, t_pf, t_ff = netd(pre_3ch.detach(), pre_flow.detach())
        opt_g.zero_grad()
        adv_s, adv_t = l2_loss(s_fr, s_ff), l2_loss(t_fr, t_ff)
        adv = adv_s + adv_t
        con = weighted_bce(predict, gt)
        err_g = adv * 1 + con * 10
        err_g.backward(retain_graph=True)
        opt_g.step()
        opt_d.zero_grad()
        e_rs, e_rt, e_fs, e_ft = bce(s_pr, ones), bce(t_pr, ones), bce(s_pf, zeros), bce(t_pf, zeros)
        real, fake = (e_rs + e_rt) * 0.5, (e_fs + e_ft) * 0.5
        err_d = (real + fake) * 0.5
        err_d.backward()
        opt_d.step()
        traj.append({"g/err_g": err_g.item(), "g/err_g_adv": adv.item(), "g/err_g_adv_s": adv_s.item(),
                     "g/err_g_adv_t": adv_t.item(), "g/err_g_con": con.item(), "d/err_d_real_s": e_rs.item(),
                     "d/err_d_real_t": e_rt.item(), "d/err_d_fake_s": e_fs.item(), "d/err_d_fake_t": e_ft.item(),
                     "d/err_d_real": real.item(), "d/err_d_fake": fake.item(), "d/err_d": err_d.item()})
        print(it, traj[-1]["g/err_g"], traj[-1]["d/err_d"], flush=True)
    torch.save({"traj": traj, "init_check": init_check, "predict_mean_last": float(predict.mean()),
                "config": {"B": B, "D": D, "S": S, "seed": 0, "data_seed0": 100}},
               os.path.join(HERE, "step_traj_cfg1.pt"))


if __name__ == "__main__":
    main()
