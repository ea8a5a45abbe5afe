"""Generates tests/golden/*.pt by running the REFERENCE's own modules (imported from /root/reference,
build container only) on seeded inputs. The fixtures pin the oracle (tests/test_oracle.py) and the
CUDA path (tests/test_parity_gpu.py) on machines where the reference is absent.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Fixtures (small on purpose; weights are re-created from the seed, not stored, wherever the nets are big):
  st_conv_small.pt    SpatioTemporalConv(8->16,k3) state_dict + input + train-mode output + input/weight grads
  convlstm_cell.pt    ConvLSTMCell(16,32,(3,3)) state_dict + inputs + (h,c)
  losses.pt           l2_loss / weighted_bce / BCELoss values on seeded tensors
  netg_netd_small.pt  NetG(ngf=8) + SDisc/TDisc(ndf=8) outputs (train mode, dropout off); weights/inputs from seed 14
  composed_small.pt   builder-defined compositions made of reference modules: NetG(3,8) + a
                      ConvLSTMCell unrolled over the latent (config 3), NetG(3,8) -> gray2rgb -> second encoder with
                      per-clip latent scores / min-max scaling / AUC over 16 clips (config 5)
  stcnn_small.pt      models/mystcnn.py AutoEncoder: predict on a seeded clip + 3 BCELoss/Adam steps (config 4)
  flow_small.pt       lib/utils.py video_to_flow (the reference function itself: cv2 Farneback + HSV encoding) on a
                      seeded smooth clip, as uint8 levels, plus cv2.calcOpticalFlowFarneback fields of two frame pairs
  eval_small.pt       threshold + morphology_proc through cv2 (lib/utils.py:139-152) and lib/evaluate.py's
                      roc / pr / f1_score through sklearn on seeded masks and scores
  step_traj_cfg1.pt   12 logged losses over 10 optimize_params steps of the full-size NetG/NetD at
                      BASELINE config 1 (B=4, 16x3x64x64; weights from torch.manual_seed(0) + weights_init,
                      NetD Linears sized for isize=64 as in SURVEY.md D4), dropout disabled.
"""
import os
import sys
import types

import numpy as np
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
from models.mygannet import NetG, NetD, SDisc, TDisc, NetgConv  # noqa: E402
from models.mystcnn import AutoEncoder  # noqa: E402
from lib.utils import weights_init, l2_loss, weighted_bce, gray2rgb  # noqa: E402
from oracle.vfd_oracle import synthetic_batch  # noqa: E402  (only the seeded input generator)


def sd_clone(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


# per-clip contrast of the 16 scoring clips (a fixed shuffle of 16 levels, so every batch mixes them differently)
SCORE_AMPS = [0.1 + 0.9 * k / 15 for k in (7, 0, 12, 3, 15, 9, 1, 5, 10, 14, 2, 6, 4, 13, 8, 11)]


def composed_fixture():
    """Compositions of reference modules only; the module construction / init order equals
    vfd_gan_b200.composed's, so the tests re-create the weights from the seed."""
    out = {}
    # ---- config 3: NetG + ConvLSTM over the latent (models/convlstm.py:199-201 idiom, unrolled by hand because
    #      ConvLSTM.forward's init_hidden hard-codes .cuda(), models/convlstm.py:60-62)
    torch.manual_seed(21)
    g = NetG(3, 8)
    cell = ConvLSTMCell((2, 2), 128, 128, (3, 3), False)
    g.apply(weights_init)
    g.dropout.p = 0.0
    g.train()
    x = torch.rand(2, 3, 32, 32, 32) * 2 - 1
    acts = {}
    g.dconv5.register_forward_hook(lambda m, i, o: acts.__setitem__("latent", o))

    def lstm_hook(m, inp):          # replace uconv5's input by the ConvLSTM output
        lat = inp[0].transpose(1, 2)                        # (B, T, C, h, w)
        h = torch.zeros(lat.shape[0], 128, 2, 2)
        c = torch.zeros(lat.shape[0], 128, 2, 2)
        hs = []
        for t in range(lat.shape[1]):
            h, c = cell(lat[:, t], (h, c))
            hs.append(h)
        acts["lstm"] = torch.stack(hs, dim=1).transpose(1, 2)
        return (acts["lstm"],)
    hook = g.uconv5.register_forward_pre_hook(lstm_hook)
    pred = g(x)
    gy = torch.randn(pred.shape)
    pred.backward(gy)
    hook.remove()
    out["lstm"] = {"init_check": {"g": g.dconv1.conv.spatial_conv.weight.detach().flatten()[:8].clone(),
                                  "cell": cell.conv.weight.detach().flatten()[:8].clone(), "x": x.flatten()[:8].clone()},
                   "predict": pred.detach(), "latent": acts["lstm"].detach(), "gy": gy,
                   "g_cell": cell.conv.weight.grad[::16].clone(),            # row samples keep the fixture small
                   "g_dconv5_t": g.dconv5.conv.temporal_conv.weight.grad[::4].clone(),
                   "g_uconv5_s": g.uconv5.conv.spatial_conv.weight.grad[::4].clone()}

    # ---- config 5: enc-dec-enc scoring sweep (definitions: models/ganomaly.py:372,396)
    torch.manual_seed(22)
    g = NetG(3, 8)
    enc = nn.ModuleList([NetgConv(3, 8), NetgConv(8, 16), NetgConv(16, 32), NetgConv(32, 64), NetgConv(64, 128)])
    g.apply(weights_init)
    enc.apply(weights_init)
    g.dropout.p = 0.0
    g.train()
    enc.train()
    latents = {}
    g.dconv5.register_forward_hook(lambda m, i, o: latents.__setitem__("i", o))
    scores, preds = [], []
    with torch.no_grad():
        for b in range(4):                                  # 4 batches x 4 clips; BN uses batch statistics
            gen = torch.Generator().manual_seed(500 + b)
            xb = torch.rand(4, 3, 16, 32, 32, generator=gen) * 2 - 1
            amp = torch.tensor([SCORE_AMPS[4 * b + j] for j in range(4)]).view(4, 1, 1, 1, 1)
            xb = xb * amp                                   # distinct contrast per clip spreads the scores out
            p = g(xb)
            h = gray2rgb(p)
            for i, blk in enumerate(enc):
                h = blk(h)
                if i < 4:
                    h = g.avgpool(h)
            li, lo = latents["i"], h
            scores.append(torch.mean(torch.pow(li - lo, 2).flatten(1), dim=1))     # ganomaly.py:372 over non-batch dims
            if b == 0:
                first = {"predict": p.clone(), "latent_i": li.clone(), "latent_o": lo.clone(),
                         "l_enc": l2_loss(lo, li), "l_con": nn.L1Loss()(p, xb[:, :1])}
    raw = torch.cat(scores)
    scaled = (raw - torch.min(raw)) / (torch.max(raw) - torch.min(raw))             # ganomaly.py:396
    # anomalous = the lowest-contrast clip of each batch: its score cluster sits between the other two, so the
    # area is well away from 0 / 1 and no positive-negative pair is closer than 25 % (robust to bf16 compute)
    labels = torch.tensor([int(SCORE_AMPS[i] == min(SCORE_AMPS[i // 4 * 4:i // 4 * 4 + 4])) for i in range(16)])
    from sklearn.metrics import roc_curve, auc
    fpr, tpr, _ = roc_curve(labels.numpy(), scaled.numpy())
    out["score"] = {"init_check": {"g": g.dconv1.conv.spatial_conv.weight.detach().flatten()[:8].clone(),
                                   "enc": enc[0].conv.spatial_conv.weight.detach().flatten()[:8].clone()},
                    "first": first, "raw": raw, "scaled": scaled, "labels": labels, "auc": float(auc(fpr, tpr))}
    torch.save(out, os.path.join(HERE, "composed_small.pt"))
    print("composed: scores", raw.tolist(), "auc", out["score"]["auc"])


def stcnn_fixture():
    torch.manual_seed(15)
    m = AutoEncoder()
    m.apply(weights_init)
    for blk in m.children():
        if hasattr(blk, "dropout"):
            blk.dropout.p = 0.0
    m.train()
    init_check = {"first": m.down_sep1.spaceconv.weight.detach().flatten()[:8].clone(),
                  "last": m.conv_last.weight.detach().flatten()[:8].clone()}
    gen = torch.Generator().manual_seed(600)
    x = torch.rand(2, 3, 16, 32, 32, generator=gen) * 2 - 1
    gt = (torch.rand(2, 1, 16, 32, 32, generator=gen) > 0.9).float()
    opt = torch.optim.Adam(m.parameters(), lr=2e-5, betas=(0.5, 0.999))      # lib/train_stcnn.py:91
    bce = nn.BCELoss()
    losses, first = [], None
    for it in range(3):                                                       # lib/train_stcnn.py:104-109
        opt.zero_grad()
        predict = m(x)
        err = bce(predict, gt)
        err.backward()
        if it == 0:
            first = {"predict": predict.detach().clone(),
                     "g_first": m.down_sep1.spaceconv.weight.grad.clone(),
                     "g_up4_last": m.up_sep4.conv_last.weight.grad.clone(),
                     "g_down4_conv_b": m.down_sep4.conv.bias.grad.clone(),
                     "rm_bn2": m.up_sep4.bn2.running_mean.clone()}
        opt.step()
        losses.append(err.item())
        print("stcnn", it, losses[-1], flush=True)
    torch.save({"init_check": init_check, "first": first, "losses": losses}, os.path.join(HERE, "stcnn_small.pt"))


def eval_fixture():
    import cv2
    import numpy as np
    from sklearn.metrics import roc_curve, auc, f1_score, precision_recall_curve
    gen = torch.Generator().manual_seed(700)
    out = {}
    cases = []
    for shape, dens in (((2, 1, 16, 24, 40), 0.75), ((1, 1, 4, 7, 5), 0.9), ((2, 1, 16, 32, 33), 0.6)):
        p = torch.rand(shape, generator=gen)
        p = (p < dens).float() * (0.5 + 0.5 * torch.rand(shape, generator=gen)) + (p >= dens).float() * 0.5 * torch.rand(shape, generator=gen)
        t = (p > torch.Tensor([0.5])).float() * 1                                     # lib/utils.py:149-152
        kernel = np.ones((5, 5), np.uint8)                                            # lib/utils.py:139-147
        m = np.stack([np.stack([cv2.morphologyEx(i, cv2.MORPH_OPEN, kernel) for i in v]) for v in t.numpy()])
        cases.append({"predict": p, "t_pre": t, "m_pre": torch.from_numpy(m)})
    out["morph"] = cases
    # binary-mask metrics exactly as MyGAN.test feeds lib/evaluate.py (models/mygannet.py:444-448)
    gts = (torch.rand(20000, generator=gen) > 0.85).int().numpy()
    noise = torch.rand(20000, generator=gen).numpy()
    pred = ((gts == 1) & (noise > 0.3) | (gts == 0) & (noise > 0.9)).astype(np.float32)
    fpr, tpr, _ = roc_curve(gts, pred)
    precision, recall, _ = precision_recall_curve(gts, pred)
    sc = pred.copy()
    sc[sc >= 0.20] = 1
    sc[sc < 0.20] = 0
    out["binary"] = {"gts": torch.from_numpy(gts), "pred": torch.from_numpy(pred), "roc": float(auc(fpr, tpr)),
                     "pr": float(auc(recall, precision)), "f1": float(f1_score(gts, sc))}
    # continuous per-clip scores with ties
    n = 3000
    lab = (torch.rand(n, generator=gen) > 0.7).int().numpy()
    s = (torch.rand(n, generator=gen) * 0.6 + torch.from_numpy(lab).float() * 0.25 * torch.rand(n, generator=gen)).numpy()
    s = np.round(s * 200) / 200                                                       # plenty of ties
    s[:5] = 0.0
    s[5] = -0.0
    fpr, tpr, _ = roc_curve(lab, s)
    precision, recall, _ = precision_recall_curve(lab, s)
    out["scores"] = {"labels": torch.from_numpy(lab), "scores": torch.from_numpy(s.astype(np.float32)),
                     "roc": float(auc(fpr, tpr)), "pr": float(auc(recall, precision))}
    torch.save(out, os.path.join(HERE, "eval_small.pt"))
    print("eval:", out["binary"]["roc"], out["binary"]["pr"], out["binary"]["f1"], out["scores"]["roc"])


def flow_clip(B, D, S, seed):
    """Seeded smooth clip in [-1, 1] (blurred noise + per-frame brightness drift) -- shared with the tests."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(B, 3, D, S + 16, S + 16, generator=g)
    vid = torch.nn.functional.avg_pool3d(base, (1, 9, 9), stride=1, padding=(0, 4, 4))[:, :, :, 8:8 + S, 8:8 + S]
    vid = (vid - vid.min()) / (vid.max() - vid.min())
    return (vid * 0.8 + 0.2 * torch.rand(B, 3, D, 1, 1, generator=g)) * 2 - 1


def flow_fixture():
    import warnings
    import cv2
    from lib.utils import video_to_flow
    from oracle import flow_oracle as FO
    warnings.simplefilter("ignore")                # np.uint8() of out-of-range floats warns in numpy 2
    out = {}
    for name, (B, D, S, seed) in {"s64": (2, 4, 64, 800), "s112": (1, 3, 112, 801), "s128": (1, 3, 128, 802)}.items():
        vid = flow_clip(B, D, S, seed)
        ref = video_to_flow(vid)                                               # the reference function, on CPU
        levels = torch.round((ref + 1) * 0.5 * 255).to(torch.uint8)
        assert float(((levels.float() / 255) * 2 - 1 - ref).abs().max()) < 1e-6
        grey = FO.gray_frames(vid.numpy())
        raw = [cv2.calcOpticalFlowFarneback(grey[0, i], grey[0, i + 1], None, 0.5, 3, 15, 3, 5, 1.2, 0)
               for i in range(2)]
        out[name] = {"cfg": (B, D, S, seed), "levels": levels, "cv2_flow_b0": torch.from_numpy(np.stack(raw))}
        print("flow", name, tuple(ref.shape), float(ref.mean()), flush=True)
    torch.save(out, os.path.join(HERE, "flow_small.pt"))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "flow":
        flow_fixture()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "new":       # only the fixtures added after the first set
        composed_fixture()
        stcnn_fixture()
        eval_fixture()
        return
    composed_fixture()
    stcnn_fixture()
    eval_fixture()
    flow_fixture()
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
