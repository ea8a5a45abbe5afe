"""GPU (-m gpu): the scoring kernels, the builder-defined compositions (BASELINE configs 3 and 5), the STCNN
(config 4) and the device versions of MyGAN.test's host detours, against the golden fixtures (reference modules /
cv2 / sklearn, tests/golden/make_golden.py) and the CPU oracle. Tolerances as in test_parity_gpu.py: fp32 kernels
1e-5, bf16-stored tensors against the operand-matched oracle 3e-3, network outputs against the fp32 fixtures
1e-2 .. 2e-2; index / count / mask work is bit-exact."""
import types

import pytest
import torch
import torch.nn.functional as F

import vfd_gan_b200 as V
from vfd_gan_b200 import ops
from oracle import vfd_oracle as O
from helpers import golden, rel, build_lstm_net, build_enc_dec_enc, build_stcnn, score_batch

pytestmark = pytest.mark.gpu
DEV = "cuda"


# ------------------------------------------------------------------------------------------ reductions
def test_latent_score_and_l2_gradients():
    torch.manual_seed(0)
    a = torch.randn(5, 20, 2, 3, 3)
    b = torch.randn(5, 20, 2, 3, 3)
    ac = ops.PackFn.apply(a.to(DEV), 0).requires_grad_(True)
    bc = ops.PackFn.apply(b.to(DEV), 0).requires_grad_(True)          # 20 -> 24 padded channels
    loss, scores = V.latent_l2_and_scores(ac, bc, 20)                  # l2_loss(latent_o, latent_i) + per-clip means
    ar, br = a.bfloat16().float().requires_grad_(True), b.bfloat16().float().requires_grad_(True)
    want = O.l2_loss(br, ar)
    assert abs(float(loss) - float(want)) <= 1e-5 * float(want)
    assert torch.allclose(scores.cpu(), O.anomaly_scores(ar, br).detach(), rtol=1e-5)
    (loss * 3.0).backward()
    (want * 3.0).backward()
    ga, gb = torch.empty_like(a, device=DEV), torch.empty_like(b, device=DEV)
    ops.unpack_ncdhw(ac.grad, ga)
    ops.unpack_ncdhw(bc.grad, gb)
    assert rel(ga, ar.grad) < 4e-3 and rel(gb, br.grad) < 4e-3        # gradients are stored in bf16
    assert float(bc.grad[..., 20:].abs().max()) == 0.0


def test_l1_and_bce_losses():
    torch.manual_seed(1)
    for n in (1, 7, 4096 + 3):
        a, b = torch.randn(n), torch.randn(n)
        b[0] = a[0]                                                    # sign(0) = 0
        ad = a.to(DEV).requires_grad_(True)
        ar = a.clone().requires_grad_(True)
        got = ops.L1LossFn.apply(ad, b.to(DEV))
        want = O.l1_loss(ar, b)
        assert abs(float(got) - float(want)) <= 1e-5 * float(want) + 1e-7
        (got * 2).backward()
        (want * 2).backward()
        assert torch.allclose(ad.grad.cpu(), ar.grad, atol=1e-7)
        p = torch.rand(n)
        p[0] = 0.0 if n > 1 else 0.3                                   # log clamp at -100
        t = (torch.rand(n) > 0.5).float()
        pd, pr = p.to(DEV).requires_grad_(True), p.clone().requires_grad_(True)
        got = ops.BceLossFn.apply(pd, t.to(DEV))
        want = F.binary_cross_entropy(pr, t)
        assert abs(float(got) - float(want)) <= 1e-5 * float(want)
        got.backward()
        want.backward()
        assert torch.allclose(pd.grad.cpu(), pr.grad, rtol=1e-4, atol=1e-6)


def test_score_scaling():
    raw = torch.tensor([0.5, 2.0, 1.25, 0.75], device=DEV)
    mm = torch.stack([raw.min(), raw.max()])
    out = torch.empty_like(raw)
    ops.score_scale(raw, mm, out)
    assert torch.equal(out.cpu(), O.minmax_scale(raw.cpu()))
    per_clip = torch.tensor([4.0, 1.0, 9.0], dtype=torch.float64, device=DEV)
    scores = torch.empty(3, device=DEV)
    mm2 = torch.tensor([float("inf"), float("-inf")], device=DEV)
    ops.score_finalize(per_clip, 0.5, scores, mm2)
    assert scores.tolist() == [2.0, 0.5, 4.5] and mm2.tolist() == [0.5, 4.5]


# ------------------------------------------------------------------------------------------ config 3
def test_netg_lstm_against_reference_fixture_and_oracle():
    f = golden("composed_small.pt")["lstm"]
    g, x = build_lstm_net()
    sd = {k: v.clone() for k, v in g.state_dict().items()}
    g = g.to(DEV).train()
    xc = ops.PackFn.apply(x.to(DEV), 0)
    logits, latent = g.forward_cl(xc)
    pred = ops.SigmoidHeadFn.apply(logits)
    assert pred.shape == f["predict"].shape and rel(pred, f["predict"]) < 2e-2
    assert rel(ops.UnpackFn.apply(latent, 128), f["latent"]) < 3e-2
    pred.backward(f["gy"].to(DEV))
    res = {}
    for rb in (True, False):
        sdo = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
        po, lo = O.netg_lstm_forward(sdo, x, True, [1.0] * 4, round_bf16=rb, return_latent=True)
        po.backward(f["gy"])
        res[rb] = (sdo, po.detach(), lo.detach())
    assert rel(pred, res[True][1]) < 1e-2
    # the ConvLSTM output is small (o * tanh(c)) and sits behind five BatchNorms over 16 samples: judge it
    # by the bf16 envelope like the gradients (distance to fp32 <= 2x the operand-matched oracle's distance)
    lat = ops.UnpackFn.apply(latent, 128)
    assert rel(lat, f["latent"]) <= max(2.0 * rel(res[True][2], f["latent"]), 2e-2)
    gw = g.clstm.cell_list[0].conv.weight.grad
    gf, gm = res[False][0]["clstm.cell_list.0.conv.weight"].grad, res[True][0]["clstm.cell_list.0.conv.weight"].grad
    assert rel(gw, gf) <= max(2.0 * rel(gm, gf), 2e-2)
    assert rel(gw[::16], f["g_cell"]) <= max(2.0 * rel(gm[::16], f["g_cell"]), 2e-2)
    for name, key in (("dconv5.conv.temporal_conv.weight", "g_dconv5_t"), ("uconv5.conv.spatial_conv.weight", "g_uconv5_s")):
        got = dict(g.named_parameters())[name].grad
        assert rel(got[::4], f[key]) <= max(2.0 * rel(res[True][0][name].grad[::4], f[key]), 2e-2), name


def test_fused_convlstm_step_equals_the_two_kernel_path():
    """models/convlstm.py:46-58 with the cell update in the gate conv's epilogue (vfd_convlstm_step_fwd, hidden sizes
    that are multiples of 64) against the gate conv + cell kernel pair that the reference fixture pins: hidden states,
    and the gradients of the input and of the gate weights / bias through a 3-step unroll."""
    torch.manual_seed(4)
    N, T, H, W, cin, hid = 2, 3, 8, 8, 64, 64
    lstm = V.ConvLSTM((H, W), cin, hid, (3, 3), 1, batch_first=True, bias=True).to(DEV)
    x = (torch.randn(N, T, H, W, cin, device=DEV) * 0.5).bfloat16()
    gout = torch.randn(N, T, H, W, hid, device=DEV).bfloat16()
    res = {}
    for fused in (True, False):
        ops.LSTM_FUSED = fused
        try:
            xin = x.clone().requires_grad_(True)
            lstm.zero_grad()
            calls = ops._lib.LAUNCHES
            out = lstm.forward_cl(xin)
            out.backward(gout)
            res[fused] = (out.detach().float(), xin.grad.float(), lstm.cell_list[0].conv.weight.grad.clone(),
                          lstm.cell_list[0].conv.bias.grad.clone())
        finally:
            ops.LSTM_FUSED = True
    for a, b, tol in zip(res[True], res[False], (4e-3, 1e-2, 3e-3, 3e-3)):
        assert a.shape == b.shape and rel(a, b) < tol, (rel(a, b), tol)
    assert ops.lstm_step_fusable(lstm.cell_list[0].conv.weight, cin + hid)
    small = V.ConvLSTM((H, W), 8, 8, (3, 3), 1, batch_first=True, bias=True).to(DEV)
    assert not ops.lstm_step_fusable(small.cell_list[0].conv.weight, 16)        # falls back to the two-kernel path


def test_gan_step_with_convlstm_bottleneck_tracks_oracle():
    """BASELINE config 3 in miniature: 32-frame clips, NetG + ConvLSTM bottleneck (T = 2), full GAN step."""
    B, D, S = 2, 32, 64
    torch.manual_seed(5)
    netg = V.NetGLstm(3, 32, isize=S)
    netd = V.NetD(types.SimpleNamespace(nfr=D, isize=S))
    netg.apply(V.weights_init)
    netd.apply(V.weights_init)
    netg.dropout.p = 0.0
    oracle = O.OracleTrainer(netg.state_dict(), netd.state_dict(), netg_fn=O.netg_lstm_forward)
    w0 = netg.clstm.cell_list[0].conv.weight.detach().clone()
    step = V.GanTrainStep(netg.to(DEV), netd.to(DEV), graph=False)
    for it in range(2):
        batch = O.synthetic_batch(B, D, S, seed=40 + it)
        step.step(*(t.to(DEV) for t in batch))
        got = step.losses_dict()
        want, _ = oracle.step(*batch, dropout_masks=[1.0] * 4)
        for k in want:
            tol = 2e-2 if "adv" in k else 1e-2
            assert abs(got[k] - want[k]) <= tol * abs(want[k]) + 1e-5, (it, k, got[k], want[k])
    # the bottleneck trains: two Adam steps move every weight by about lr * sign(grad), so compare the updates
    # (elements with a near-zero gradient may flip sign under bf16 compute; most must agree)
    du = netg.clstm.cell_list[0].conv.weight.detach().cpu() - w0
    do = oracle.pg["clstm.cell_list.0.conv.weight"].detach() - w0
    assert float(du.abs().max()) > 1e-5
    assert float((du * do).sum() / (du.norm() * do.norm())) > 0.9


# ------------------------------------------------------------------------------------------ config 5
def test_anomaly_score_sweep_against_reference_fixture():
    f = golden("composed_small.pt")["score"]
    m = build_enc_dec_enc()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.to(DEV).train()
    scorer = V.AnomalyScorer(m)
    matched = []
    for b in range(4):
        xb = score_batch(b)
        scorer.score_batch(xb.to(DEV))
        with torch.no_grad():
            _, li, lo = O.enc_dec_enc_forward(sd, xb, True, [1.0] * 4, round_bf16=True)
        matched.append(O.anomaly_scores(li, lo))
    scaled, raw = scorer.finish()
    matched = torch.cat(matched)
    raw_c, scaled_c = raw.cpu(), scaled.cpu()
    assert torch.allclose(raw_c, matched, rtol=1e-2), (raw_c, matched)       # operand-matched oracle
    assert torch.allclose(raw_c, f["raw"], rtol=3e-2), (raw_c, f["raw"])     # fp32 reference composition
    assert torch.allclose(scaled_c, O.minmax_scale(raw_c), atol=1e-6)
    assert torch.allclose(scaled_c, f["scaled"], atol=3e-2)
    # ranking: identical wherever the reference separates two clips by more than the bf16 noise (2 %)
    ref = f["raw"]
    for i in range(16):
        for j in range(16):
            if ref[i] > 1.02 * ref[j]:
                assert raw_c[i] > raw_c[j], (i, j)
    # AUC to 3 decimals: device kernel on the device scores vs sklearn on the reference scores
    labels = f["labels"].float()
    area = V.evaluate.roc_auc(labels.to(DEV), scaled)
    assert round(float(area[0]), 3) == round(f["auc"], 3)
    assert int(area[1]) == int(labels.sum()) and int(area[2]) == 16 - int(labels.sum())
    assert abs(O.evaluate(labels.numpy(), scaled_c.numpy(), "roc") - float(area[0])) < 1e-12


def test_enc_dec_enc_training_losses_have_gradients():
    """l_con (L1) and l_enc (latent L2) of the composition back-propagate into both encoders and the decoder."""
    m = build_enc_dec_enc().to(DEV).train()
    x = score_batch(1).to(DEV)
    predict, li, lo = m.forward_cl(ops.PackFn.apply(x, 0))
    l_enc, _ = V.latent_l2_and_scores(li, lo, m.latent_channels)
    l_con = ops.L1LossFn.apply(predict, x[:, :1])
    (l_con * 50 + l_enc).backward()
    for name in ("netg.dconv1.conv.spatial_conv.weight", "netg.uconv1.conv.temporal_conv.weight",
                 "encoder2.dconv5.conv.temporal_conv.weight", "netg.conv_last.weight"):
        g = dict(m.named_parameters())[name].grad
        assert g is not None and torch.isfinite(g).all() and float(g.abs().max()) > 0, name


# ------------------------------------------------------------------------------------------ config 4
def test_stcnn_against_reference_fixture():
    f = golden("stcnn_small.pt")
    m, x, gt = build_stcnn()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.to(DEV)
    tr = V.StcnnTrainStep(m)
    losses = []
    for it in range(3):
        err = tr.step(x.to(DEV), gt.to(DEV))
        losses.append(float(err))
        if it == 0:
            first = f["first"]
            with torch.no_grad():
                pm = O.autoencoder_forward({k: v.clone() for k, v in sd.items()}, x, True, round_bf16=True)
            assert rel(tr.predict, pm) < 1e-2 and rel(tr.predict, first["predict"]) < 2e-2
            assert rel(m.up_sep4.bn2.running_mean, first["rm_bn2"]) < 3e-2
    for got, want in zip(losses, f["losses"]):
        assert abs(got - want) <= 1e-2 * abs(want), (losses, f["losses"])


def test_stcnn_gradients_inside_bf16_envelope():
    f = golden("stcnn_small.pt")["first"]
    m, x, gt = build_stcnn()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.to(DEV).train()
    pred = m(x.to(DEV))
    ops.BceLossFn.apply(pred, gt.to(DEV)).backward()
    sdo = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    F.binary_cross_entropy(O.autoencoder_forward(sdo, x, True, round_bf16=True), gt).backward()
    for name, key in (("down_sep1.spaceconv.weight", "g_first"), ("up_sep4.conv_last.weight", "g_up4_last"),
                      ("down_sep4.conv.bias", "g_down4_conv_b")):
        got, matched = dict(m.named_parameters())[name].grad, sdo[name].grad
        assert rel(got, f[key]) <= max(2.0 * rel(matched, f[key]), 2e-2), (name, rel(got, f[key]), rel(matched, f[key]))


def test_stcnn_dropout_path_runs():
    m, x, gt = build_stcnn()
    for blk in m.children():
        if hasattr(blk, "dropout"):
            blk.dropout.p = 0.25
    m = m.to(DEV).train()
    xc = ops.PackFn.apply(x.to(DEV), 0)
    # the dropout op itself: deterministic in the seed, inverted scaling, keep rate 1 - p
    t = torch.ones(2, 4, 16, 16, 64, dtype=torch.bfloat16, device=DEV)
    d1, d2, d3 = (ops.IdentityPoolFn.apply(t, (1, 1, 1), 0.25, s) for s in (11, 11, 12))
    assert torch.equal(d1, d2) and not torch.equal(d1, d3)
    assert set(d1.float().unique().tolist()) == {0.0, float(torch.tensor(1 / 0.75).bfloat16())}
    assert abs(float((d1 > 0).float().mean()) - 0.75) < 0.01
    # whole net: same seeds -> same masks (run-to-run differences are fp32 atomics order amplified by the
    # 8-sample BatchNorms at the bottleneck); other seeds -> visibly different output
    a = m.forward_cl(xc, dropout_seeds=[1, 2, 3, 4])
    b = m.forward_cl(xc, dropout_seeds=[1, 2, 3, 4])
    c = m.forward_cl(xc, dropout_seeds=[5, 6, 7, 8])
    assert torch.isfinite(a).all() and torch.isfinite(c).all()
    print("stcnn dropout: same seeds", rel(a, b), "other seeds", rel(a, c))
    assert rel(a, b) < 0.5 * rel(a, c)
    a.float().sum().backward()
    assert torch.isfinite(m.down_sep1.spaceconv.weight.grad).all()


# ------------------------------------------------------------------------------------------ MyGAN.test detours
def test_threshold_and_opening_bit_exact_against_cv2_fixture():
    for case in golden("eval_small.pt")["morph"]:
        t, m = V.evaluate.threshold_open(case["predict"].to(DEV))
        assert torch.equal(t.cpu(), case["t_pre"])
        assert torch.equal(m.cpu(), case["m_pre"])
        assert torch.equal(V.evaluate.morphology_proc(case["t_pre"].to(DEV)).cpu(), case["m_pre"])


def test_threshold_and_opening_at_full_size_properties():
    """BASELINE-size masks (32 x 16 x 112 x 112): opening is idempotent and anti-extensive, and matches the
    oracle restatement."""
    torch.manual_seed(3)
    p = torch.rand(32, 1, 16, 112, 112, device=DEV)
    p = F.avg_pool3d(p, 3, stride=1, padding=1)                  # blobs, so the opening keeps something
    t, m = V.evaluate.threshold_open(p, 0.5)
    assert float(m.sum()) > 0 and bool((m <= t).all())
    assert torch.equal(V.evaluate.morphology_proc(m), m)
    assert torch.equal(m[:2].cpu(), O.morphology_proc(O.threshold(p[:2].cpu())))


def test_confusion_counts_and_binary_metrics_against_sklearn_fixture():
    b = golden("eval_small.pt")["binary"]
    gts, pred = b["gts"].float().to(DEV), b["pred"].to(DEV)
    counts = V.evaluate.confusion_counts(gts, pred, 0.20)
    half = gts.numel() // 2 + 1                                   # accumulate over two ragged batches as well
    c2 = V.evaluate.confusion_counts(gts[:half], pred[:half], 0.20)
    c2 = V.evaluate.confusion_counts(gts[half:].clone(), pred[half:].clone(), 0.20, c2)
    assert torch.equal(counts, c2) and int(counts.sum()) == gts.numel()
    got = V.evaluate.binary_metrics_from_counts(*counts.tolist())
    for key in ("roc", "pr", "f1"):
        assert abs(got[key] - b[key]) < 1e-9, key


def test_roc_auc_kernel_against_sklearn_fixture():
    s = golden("eval_small.pt")["scores"]
    out = V.evaluate.roc_auc(s["labels"].to(DEV), s["scores"].to(DEV))
    assert abs(float(out[0]) - s["roc"]) < 1e-12
    assert abs(float(out[3]) - s["pr"]) < 1e-12                   # lib/evaluate.py 'pr': auc(recall, precision)
    assert int(out[1]) == int(s["labels"].sum())
    torch.manual_seed(9)
    for n in (2, 1000, 1025, 16384):                              # padding / several elements per thread
        lab = (torch.rand(n) > 0.5).float()
        lab[0], lab[1] = 0.0, 1.0
        sc = torch.randn(n).round(decimals=1)
        got = V.evaluate.roc_auc(lab.to(DEV), sc.to(DEV))
        assert abs(float(got[0]) - O.evaluate(lab.numpy(), sc.numpy(), "roc")) < 1e-12, n
        assert abs(float(got[3]) - O.evaluate(lab.numpy(), sc.numpy(), "pr")) < 1e-12, n
    with pytest.raises(RuntimeError):
        V.evaluate.roc_auc(torch.zeros(20000, device=DEV), torch.zeros(19999, device=DEV))


def _areas(lab, sc):
    return O.evaluate(lab.numpy(), sc.numpy(), "roc"), O.evaluate(lab.numpy(), sc.numpy(), "pr")


def test_voxel_level_roc_pr_against_sklearn_at_1e6():
    """lib/evaluate.py's roc / pr on a million voxels (test.py:186-202 hands it every voxel of the test set): the
    multi-block sort + run-length pass against sklearn on the same arrays. ROC pair counts are exact integers, so the
    area differs from sklearn's trapezoid sum only by its float rounding (1e-12); the PR trapezoids are summed in a
    different order than np.trapz (1e-10)."""
    g = torch.Generator().manual_seed(11)
    n = 1_000_003
    lab = (torch.rand(n, generator=g) > 0.9).float()
    cases = {
        "continuous": torch.sigmoid(torch.randn(n, generator=g) + 1.5 * lab),
        "quantised (heavy ties)": (torch.sigmoid(torch.randn(n, generator=g) + 1.5 * lab) * 255).round() / 255,
        "signed with zeros": (torch.randn(n, generator=g) + lab).round(decimals=2) * (torch.rand(n, generator=g) > 0.2),
    }
    cases["signed with zeros"][::7] *= -1.0                       # -0.0 and +0.0 are one score
    for name, sc in cases.items():
        sc = sc.float()
        want_roc, want_pr = _areas(lab, sc)
        got = V.evaluate.roc_auc(lab.to(DEV), sc.to(DEV)).tolist()
        assert abs(got[0] - want_roc) < 1e-12, (name, got[0], want_roc)
        assert abs(got[3] - want_pr) < 1e-10, (name, got[3], want_pr)
        assert got[1] == float(lab.sum()) and got[2] == n - float(lab.sum())
        again = V.evaluate.roc_auc(lab.to(DEV), sc.to(DEV)).tolist()
        assert again == got, name                                 # deterministic: no floating-point atomics


def test_voxel_level_roc_edges_and_both_kernels_agree():
    g = torch.Generator().manual_seed(12)
    n = 16384                                                     # the size both kernels accept
    lab = (torch.rand(n, generator=g) > 0.5).float()
    sc = torch.randn(n, generator=g).round(decimals=1)
    small = torch.empty(4, dtype=torch.float64, device=DEV)
    large = torch.empty(4, dtype=torch.float64, device=DEV)
    from vfd_gan_b200 import _lib
    ws = torch.empty(int(_lib.lib().vfd_roc_auc_large_workspace(n)), dtype=torch.uint8, device=DEV)
    ops.roc_auc_op(sc.to(DEV), lab.to(DEV), small)
    ops.roc_auc_large_op(sc.to(DEV), lab.to(DEV), large, ws)
    assert small[1:3].tolist() == large[1:3].tolist()
    assert abs(float(small[0] - large[0])) < 1e-13 and abs(float(small[3] - large[3])) < 1e-12
    for n in (16385, 2048 * 3 + 1 + 16384, 50_000):               # one past the single-block limit, ragged tiles
        lab = (torch.rand(n, generator=g) > 0.7).float()
        sc = torch.rand(n, generator=g).round(decimals=3)
        want_roc, want_pr = _areas(lab, sc)
        got = V.evaluate.roc_auc(lab.to(DEV), sc.to(DEV)).tolist()
        assert abs(got[0] - want_roc) < 1e-12 and abs(got[3] - want_pr) < 1e-10, n
    n = 40_000
    lab = (torch.arange(n) % 3 == 0).float()
    assert float(V.evaluate.roc_auc(lab.to(DEV), torch.full((n,), 0.25, device=DEV))[0]) == 0.5   # one big tie
    sep = V.evaluate.roc_auc(lab.to(DEV), (lab * 2 - 1 + 0.1 * torch.rand(n, generator=g)).to(DEV)).tolist()
    assert sep[0] == 1.0 and abs(sep[3] - 1.0) < 1e-12                                          # separable
    one_class = V.evaluate.roc_auc(torch.zeros(n, device=DEV), torch.rand(n, device=DEV)).tolist()
    assert one_class[0] != one_class[0] and one_class[1] == 0.0 and one_class[2] == float(n)   # NaN area, counts kept
    with pytest.raises(RuntimeError):
        ops.roc_auc_large_op(torch.zeros(n, device=DEV), torch.zeros(n, device=DEV), large, ws[:1024])


def test_voxel_level_roc_properties_at_1e8():
    """Size-independent properties at the size of a whole test sweep (1e8 voxels = 62 clips of 16 x 112 x 112 x 8):
    swapping the classes gives exactly 1 - AUC (integer pair counts), a strictly increasing map of the scores changes
    nothing, and a subsample estimate agrees with sklearn on that subsample."""
    g = torch.Generator(device=DEV).manual_seed(13)
    n = 100_000_000
    lab = (torch.rand(n, device=DEV, generator=g) > 0.95).float()
    sc = torch.sigmoid(torch.randn(n, device=DEV, generator=g) + 2.0 * lab)
    a = V.evaluate.roc_auc(lab, sc).tolist()
    b = V.evaluate.roc_auc(1.0 - lab, sc).tolist()
    assert a[1] == b[2] and a[2] == b[1] and a[1] + a[2] == n
    assert abs(a[0] + b[0] - 1.0) < 1e-15
    c = V.evaluate.roc_auc(lab, sc * 3.0 - 1.0).tolist()         # float rounding may merge a few neighbours into ties
    assert abs(c[0] - a[0]) < 1e-9 and abs(c[3] - a[3]) < 1e-6
    # analytic value: P(sigmoid(z1 + 2) > sigmoid(z0)) = Phi(2 / sqrt(2)) = 0.92135...
    assert abs(a[0] - 0.9213503964748574) < 5e-4
    idx = torch.randint(0, n, (2_000_000,), device=DEV, generator=g)
    want_roc, want_pr = _areas(lab[idx].cpu(), sc[idx].cpu())
    sub = V.evaluate.roc_auc(lab[idx], sc[idx]).tolist()
    assert abs(sub[0] - want_roc) < 1e-12 and abs(sub[3] - want_pr) < 1e-10
    assert abs(sub[0] - a[0]) < 2e-3


def test_voxel_curve_accumulator_follows_test_py():
    """test.py:175-202: per batch append gt / predict voxels, then lib/evaluate.py's roc / pr / f1_score over all."""
    g = torch.Generator().manual_seed(14)
    acc = V.evaluate.VoxelCurveAccumulator(DEV, capacity=1000)    # forces two growths
    gts, predicts = [], []
    for b in range(3):
        gt = (torch.rand(2, 1, 4, 20, 20, generator=g) > 0.8).float()
        predict = torch.sigmoid(torch.randn(2, 1, 4, 20, 20, generator=g) + 2 * gt)
        acc.add(gt.to(DEV), predict.to(DEV))
        gts.append(gt.permute(0, 2, 3, 4, 1).numpy())
        predicts.append(predict.permute(0, 2, 3, 4, 1).numpy())
    import numpy as np
    gts = np.asarray(np.stack(gts), dtype=np.int32).flatten()
    predicts = np.asarray(np.stack(predicts)).flatten()
    got = acc.result()
    assert abs(got["roc"] - O.evaluate(gts, predicts, "roc")) < 1e-12
    assert abs(got["pr"] - O.evaluate(gts, predicts, "pr")) < 1e-12
    assert abs(got["f1"] - O.evaluate(gts, predicts, "f1_score")) < 1e-12


def test_empty_batches_are_no_ops():
    """Zero clips: every new entry point returns without launching (like the conv path, test_parity_gpu)."""
    e5 = torch.empty(0, 1, 16, 24, 24, device=DEV)
    t, m = V.evaluate.threshold_open(e5)
    assert t.shape == e5.shape and m.shape == e5.shape
    counts = V.evaluate.confusion_counts(torch.empty(0, device=DEV), torch.empty(0, device=DEV), 0.2)
    assert counts.tolist() == [0, 0, 0, 0]
    lat = torch.empty(0, 1, 2, 2, 128, dtype=torch.bfloat16, device=DEV)
    assert V.anomaly_scores(lat, lat, 128).shape == (0,)
    assert V.video_to_flow(torch.empty(0, 3, 16, 32, 32, device=DEV)).shape == (0, 3, 16, 32, 32)
    out = V.evaluate.roc_auc(torch.empty(0, device=DEV), torch.empty(0, device=DEV))
    assert out[0].isnan() and out[1] == 0 and out[2] == 0


def test_gan_evaluator_against_oracle_composition():
    """MyGAN.test on the device (models/mygannet.py:369-475) vs the same loop written with the oracle pieces
    (generator, threshold, cv2-pinned opening, cv2-pinned flow, discriminator, sklearn metrics). About 1-2 % of
    the voxels of a random-init generator sit within bf16 noise of the 0.5 threshold, hence the 2e-2 tolerance on
    the mask metrics."""
    import numpy as np
    from oracle import flow_oracle as FO
    from helpers import build_cfg1_nets, flow_clip
    B, D, S = 2, 16, 64
    netg, netd = build_cfg1_nets()
    sd_g = {k: v.clone() for k, v in netg.state_dict().items()}
    sd_d = {k: v.clone() for k, v in netd.state_dict().items()}
    ev = V.evaluate.GanEvaluator(netg.to(DEV).train(), netd.to(DEV).train())
    want = {k: 0.0 for k in V.evaluate.TEST_KEYS}
    gts, preds = [], []
    bce = torch.nn.BCELoss()
    for it in range(2):
        inp = flow_clip(B, D, S, 60 + it)
        gt = (flow_clip(B, D, S, 70 + it)[:, :1] > 0.1).float()
        ev.add_batch(inp.to(DEV), gt.to(DEV))
        with torch.no_grad():
            p = O.netg_forward(sd_g, inp, True, [1.0] * 4)
            m = O.morphology_proc(O.threshold(p))
            gt3, pre3 = O.gray2rgb(gt), O.gray2rgb(p)
            gt_flow = torch.from_numpy(FO.video_to_flow(gt3.numpy())[0])
            pre_flow = torch.from_numpy(FO.video_to_flow(pre3.numpy())[0])
            s_pr, s_fr, t_pr, t_fr = O.netd_forward(sd_d, gt3, gt_flow, True)
            s_pf, s_ff, t_pf, t_ff = O.netd_forward(sd_d, pre3, pre_flow, True)
            adv_s, adv_t, con = O.l2_loss(s_fr, s_ff), O.l2_loss(t_fr, t_ff), O.weighted_bce(p, gt)
            ones, zeros = torch.ones(B), torch.zeros(B)
            e = [bce(s_pr, ones), bce(t_pr, ones), bce(s_pf, zeros), bce(t_pf, zeros)]
            real, fake = (e[0] + e[1]) * 0.5, (e[2] + e[3]) * 0.5
            vals = e + [real, fake, (real + fake) * 0.5, adv_s, adv_t, adv_s + adv_t, con, adv_t * 1 + con * 10]
        for k, v in zip(V.evaluate.TEST_KEYS, vals):
            want[k] += float(v) / 2
        gts.append(gt.numpy().astype(np.int32).ravel())
        preds.append(m.numpy().ravel())
    errors, scores = ev.result()
    gts, preds = np.concatenate(gts), np.concatenate(preds)
    assert 0.02 < preds.mean() < 0.98                       # a non-trivial mask survives the opening
    for metric, key in (("roc", "score/roc"), ("pr", "score/pr"), ("f1_score", "score/f1")):
        assert abs(scores[key] - O.evaluate(gts, preds, metric)) < 2e-2, (key, scores[key])
    for k in V.evaluate.TEST_KEYS:
        tol = 5e-2 if ("adv" in k or "/err_g/" in k) else 2e-2
        assert np.isfinite(want[k]), k
        assert abs(errors[k] - want[k]) <= tol * abs(want[k]) + 1e-4, (k, errors[k], want[k])
