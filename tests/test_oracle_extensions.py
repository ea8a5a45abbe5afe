"""CPU: the oracle's restatements of the builder-defined compositions (configs 3 / 5), the STCNN (config 4) and
the MyGAN.test host detours against the fixtures produced from the reference's own modules, cv2 and sklearn
(tests/golden/make_golden.py). fp32 on both sides -> tight tolerances."""
import numpy as np
import torch

from oracle import vfd_oracle as O
from helpers import golden, rel, build_lstm_net, build_enc_dec_enc, build_stcnn, score_batch


def test_netg_lstm_composition_matches_reference_fixture():
    f = golden("composed_small.pt")["lstm"]
    g, x = build_lstm_net()
    assert torch.equal(g.dconv1.conv.spatial_conv.weight.detach().flatten()[:8], f["init_check"]["g"])
    assert torch.equal(g.clstm.cell_list[0].conv.weight.detach().flatten()[:8], f["init_check"]["cell"])
    assert torch.equal(x.flatten()[:8], f["init_check"]["x"])
    sd = {k: v.clone() for k, v in g.state_dict().items()}
    for k in sd:
        if sd[k].is_floating_point() and "running" not in k:
            sd[k].requires_grad_(True)
    pred, latent = O.netg_lstm_forward(sd, x, True, [1.0] * 4, return_latent=True)
    assert torch.allclose(pred, f["predict"], atol=1e-5) and rel(latent, f["latent"]) < 1e-5
    pred.backward(f["gy"])
    assert rel(sd["clstm.cell_list.0.conv.weight"].grad[::16], f["g_cell"]) < 1e-4
    assert rel(sd["dconv5.conv.temporal_conv.weight"].grad[::4], f["g_dconv5_t"]) < 1e-4
    assert rel(sd["uconv5.conv.spatial_conv.weight"].grad[::4], f["g_uconv5_s"]) < 1e-4


def test_enc_dec_enc_scores_match_reference_fixture():
    f = golden("composed_small.pt")["score"]
    m = build_enc_dec_enc()
    assert torch.equal(m.netg.dconv1.conv.spatial_conv.weight.detach().flatten()[:8], f["init_check"]["g"])
    assert torch.equal(m.encoder2.dconv1.conv.spatial_conv.weight.detach().flatten()[:8], f["init_check"]["enc"])
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    raw = []
    with torch.no_grad():
        for b in range(4):
            xb = score_batch(b)
            p, li, lo = O.enc_dec_enc_forward(sd, xb, True, [1.0] * 4)
            raw.append(O.anomaly_scores(li, lo))
            if b == 0:
                first = f["first"]
                assert torch.allclose(p, first["predict"], atol=1e-5)
                assert rel(li, first["latent_i"]) < 1e-5 and rel(lo, first["latent_o"]) < 1e-5
                assert torch.allclose(O.l2_loss(lo, li), first["l_enc"], rtol=1e-5)
                assert torch.allclose(O.l1_loss(p, xb[:, :1]), first["l_con"], rtol=1e-5)
    raw = torch.cat(raw)
    assert torch.allclose(raw, f["raw"], rtol=1e-4)
    assert torch.allclose(O.minmax_scale(raw), f["scaled"], atol=1e-4)
    assert abs(O.evaluate(f["labels"].numpy(), O.minmax_scale(raw).numpy(), "roc") - f["auc"]) < 5e-4


def test_stcnn_matches_reference_fixture():
    f = golden("stcnn_small.pt")
    m, x, gt = build_stcnn()
    assert torch.equal(m.down_sep1.spaceconv.weight.detach().flatten()[:8], f["init_check"]["first"])
    assert torch.equal(m.conv_last.weight.detach().flatten()[:8], f["init_check"]["last"])
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    params = [k for k in sd if sd[k].is_floating_point() and "running" not in k]
    for k in params:
        sd[k].requires_grad_(True)
    opt = torch.optim.Adam([sd[k] for k in params], lr=2e-5, betas=(0.5, 0.999))
    for it, want in enumerate(f["losses"]):
        opt.zero_grad()
        predict = O.autoencoder_forward(sd, x, True)
        err = torch.nn.functional.binary_cross_entropy(predict, gt)
        err.backward()
        if it == 0:
            first = f["first"]
            assert torch.allclose(predict, first["predict"], atol=1e-5)
            assert rel(sd["down_sep1.spaceconv.weight"].grad, first["g_first"]) < 1e-4
            assert rel(sd["up_sep4.conv_last.weight"].grad, first["g_up4_last"]) < 1e-4
            assert rel(sd["down_sep4.conv.bias"].grad, first["g_down4_conv_b"]) < 1e-4
            assert torch.allclose(sd["up_sep4.bn2.running_mean"], first["rm_bn2"], atol=1e-6)
        opt.step()
        assert abs(err.item() - want) <= 2e-4 * abs(want), (it, err.item(), want)


def test_threshold_and_opening_match_cv2_fixture():
    for case in golden("eval_small.pt")["morph"]:
        t = O.threshold(case["predict"])
        assert torch.equal(t, case["t_pre"])
        assert torch.equal(O.morphology_proc(t), case["m_pre"])


def test_metrics_match_sklearn_fixture():
    from vfd_gan_b200.evaluate import binary_metrics_from_counts
    f = golden("eval_small.pt")
    b = f["binary"]
    gts, pred = b["gts"].numpy(), b["pred"].numpy()
    for metric, key in (("roc", "roc"), ("pr", "pr"), ("f1_score", "f1")):
        assert abs(O.evaluate(gts, pred, metric) - b[key]) < 1e-12
    tp = int(((gts == 1) & (pred >= 0.2)).sum())
    fp = int(((gts == 0) & (pred >= 0.2)).sum())
    fn = int(((gts == 1) & (pred < 0.2)).sum())
    tn = int(((gts == 0) & (pred < 0.2)).sum())
    got = binary_metrics_from_counts(tp, fp, fn, tn)            # host formulas over the device kernel's counts
    for key in ("roc", "pr", "f1"):
        assert abs(got[key] - b[key]) < 1e-9, key
    s = f["scores"]
    assert abs(O.evaluate(s["labels"].numpy(), s["scores"].numpy(), "roc") - s["roc"]) < 1e-12
    # tie-aware Mann-Whitney statement of the same area (what vfd_roc_auc computes), in numpy
    lab, sc = s["labels"].numpy().astype(bool), s["scores"].numpy().astype(np.float64)
    neg = np.sort(sc[~lab])
    below = np.searchsorted(neg, sc[lab], side="left")
    upto = np.searchsorted(neg, sc[lab], side="right")
    area = float((below + 0.5 * (upto - below)).sum() / (lab.sum() * (~lab).sum()))
    assert abs(area - s["roc"]) < 1e-12
    assert abs(O.evaluate(s["labels"].numpy(), s["scores"].numpy(), "pr") - s["pr"]) < 1e-12


def test_binary_metric_formulas_on_edge_cases():
    """binary_metrics_from_counts against the sklearn calls of lib/evaluate.py on masks with no predicted
    positives, all predicted positives, a perfect mask and an inverted one."""
    import warnings
    from vfd_gan_b200.evaluate import binary_metrics_from_counts
    rng = np.random.default_rng(3)
    gts = (rng.random(4000) > 0.7).astype(np.int32)
    cases = {"none": np.zeros(4000, np.float32), "all": np.ones(4000, np.float32), "perfect": gts.astype(np.float32),
             "inverted": 1.0 - gts.astype(np.float32), "noisy": ((gts == 1) ^ (rng.random(4000) > 0.8)).astype(np.float32)}
    for name, pred in cases.items():
        tp = int(((gts == 1) & (pred >= 0.2)).sum())
        fp = int(((gts == 0) & (pred >= 0.2)).sum())
        fn = int(((gts == 1) & (pred < 0.2)).sum())
        tn = int(((gts == 0) & (pred < 0.2)).sum())
        got = binary_metrics_from_counts(tp, fp, fn, tn)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for metric, key in (("roc", "roc"), ("pr", "pr"), ("f1_score", "f1")):
                assert abs(got[key] - O.evaluate(gts, pred, metric)) < 1e-9, (name, key, got[key])
