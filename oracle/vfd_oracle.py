"""CPU oracle for the vfd_gan ``mygan`` hot path -- TEST INFRASTRUCTURE ONLY.

This file restates, as plain functions over a ``state_dict``, what the reference computes with its
``nn.Module`` classes. It is the checker for the CUDA path: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import
it; nothing under ``vfd_gan_b200/`` does, and the product path fails loudly without its CUDA
extension instead of falling back to this.

Where the arithmetic lives: the reference has no arithmetic of its own -- every op is a PyTorch
library call (pinned torch==1.3.1, Pipfile:14; the semantics relied on here are unchanged up to the
torch 2.11 in this image: cross-correlation Conv3d with zero padding, BatchNorm3d eps 1e-5 /
momentum 0.1 / biased batch variance / unbiased running variance, trilinear Upsample with
align_corners, BCELoss log clamp at -100, Adam without amsgrad). The oracle therefore calls the same
``torch.nn.functional`` primitives on CPU fp32; each function cites the reference lines it follows.

Pinning: the reference ships no tests or golden vectors. The oracle is
pinned instead against outputs of the reference's own modules executed in the build container
(``tests/test_oracle_vs_reference.py`` when /root/reference is present) and against the fixtures
under ``tests/golden/`` generated from those modules by ``tests/golden/make_golden.py``.

``round_bf16=True`` makes the oracle "operand matched": it rounds to bfloat16 at exactly the points
where the CUDA path stores bf16 (conv operands and stored activations), so the remaining
difference is fp32 accumulation order -- this is the oracle the per-layer 1e-3 gates use.
"""
import math

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
def _r(t, on):
    """round-trip through bfloat16 when operand matching is on"""
    return t.bfloat16().float() if on else t


def intermed_channels(cin, cout, k):
    """models/spatiotempconv.py:44-45"""
    kt, kh, kw = k
    return int(math.floor((kt * kh * kw * cin * cout) / (kh * kw * cin + kt * cout)))


def batch_norm(sd, prefix, x, train, pre_bias=None):
    """nn.BatchNorm3d forward incl. running-stat side effects (models/spatiotempconv.py:51,63;
    models/mygannet.py:19,25,109,114). ``sd`` buffers are updated in place when ``train``.
    ``pre_bias`` (operand-matched mode only): a conv bias that was left out of ``x``; a per-channel
    constant cancels in training-mode normalisation and only shifts ``running_mean``."""
    rm, rv = sd.get(prefix + ".running_mean"), sd.get(prefix + ".running_var")
    if train and pre_bias is not None and rm is not None:
        tmp = rm.clone()  # autograd keeps the tensor handed to batch_norm; update a copy of it
        out = F.batch_norm(x, tmp, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], training=True, momentum=0.1,
                           eps=1e-5)
        rm.copy_(tmp + 0.1 * pre_bias.detach())
    else:
        out = F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], training=train or rm is None,
                           momentum=0.1, eps=1e-5)
    if train and prefix + ".num_batches_tracked" in sd:
        sd[prefix + ".num_batches_tracked"] += 1
    return out


def _conv_bn(sd, conv_prefix, bn_prefix, x, padding, train, rb):
    """conv -> BatchNorm. Operand-matched mode mirrors the CUDA path's storage points: the bias is
    kept out of the bf16-stored conv output when a training-mode BatchNorm follows."""
    w, b = sd[conv_prefix + ".weight"], sd.get(conv_prefix + ".bias")
    if rb and train:
        y = _r(F.conv3d(_r(x, rb), _r(w, rb), None, padding=padding), rb)
        return batch_norm(sd, bn_prefix, y, train, pre_bias=b)
    y = _r(F.conv3d(_r(x, rb), _r(w, rb), b, padding=padding), rb)
    return batch_norm(sd, bn_prefix, y, train)


def net_conv(sd, prefix, x, kernel, slope, train=True, round_bf16=False):
    """NetgConv / NetdConv forward: SpatioTemporalConv -> BatchNorm3d -> LeakyReLU(slope)
    (models/mygannet.py:22-28, 112-116), with SpatioTemporalConv.forward =
    temporal_conv(relu(bn(spatial_conv(x)))) (models/spatiotempconv.py:62-65); stride 1 and
    padding kernel//2 as every hot-path use has."""
    kt, kh, kw = kernel
    rb = round_bf16
    c = prefix + ".conv"
    a = _r(F.relu(_conv_bn(sd, c + ".spatial_conv", c + ".bn", x, (0, kh // 2, kw // 2), train, rb)), rb)
    z = _conv_bn(sd, c + ".temporal_conv", prefix + ".bn", a, (kt // 2, 0, 0), train, rb)
    return _r(F.leaky_relu(z, slope), rb)


def st_conv(sd, prefix, x, kernel, train=True, round_bf16=False):
    """Stand-alone SpatioTemporalConv.forward (models/spatiotempconv.py:62-65)."""
    kt, kh, kw = kernel
    rb = round_bf16
    a = _r(F.relu(_conv_bn(sd, prefix + ".spatial_conv", prefix + ".bn", x, (0, kh // 2, kw // 2), train, rb)), rb)
    return F.conv3d(a, _r(sd[prefix + ".temporal_conv.weight"], rb), sd.get(prefix + ".temporal_conv.bias"),
                    padding=(kt // 2, 0, 0))


# ------------------------------------------------------------------------------------------------
# networks
# ------------------------------------------------------------------------------------------------
def netg_forward(sd, x, train=True, dropout_masks=None, dropout_p=0.25, round_bf16=False, return_latent=False,
                 bottleneck=None):
    """NetG.forward (models/mygannet.py:55-101). ``dropout_masks``: optional list of four
    multiplier tensors (already scaled by 1/(1-p)) applied after uconv5..uconv2, replacing
    nn.Dropout so a test can share masks with the CUDA path; otherwise F.dropout draws from the
    torch RNG in the reference's order."""
    rb = round_bf16
    k = (3, 3, 3)
    skips = []
    h = x
    for i in range(1, 5):
        d = net_conv(sd, f"dconv{i}", h, k, 0.2, train, rb)
        skips.append(d)
        h = _r(F.avg_pool3d(d, 2), rb)
    latent = net_conv(sd, "dconv5", h, k, 0.2, train, rb)
    if bottleneck is not None:      # builder-defined composition (config 3): see netg_lstm_forward
        latent = bottleneck(latent)

    def drop(t, i):
        if dropout_masks is not None:
            return _r(t * dropout_masks[i], rb)
        return F.dropout(t, dropout_p, training=train)

    h = drop(net_conv(sd, "uconv5", latent, k, 0.2, train, rb), 0)
    for n, i in enumerate((4, 3, 2, 1)):
        h = _r(F.interpolate(h, scale_factor=2, mode="trilinear", align_corners=True), rb)
        h = torch.cat([h, skips[i - 1]], dim=1)
        h = net_conv(sd, f"uconv{i}", h, k, 0.2, train, rb)
        if i > 1:
            h = drop(h, n + 1)
    logits = F.conv3d(h, _r(sd["conv_last.weight"], rb), None, padding=1)
    predict = torch.sigmoid(logits)
    return (predict, latent) if return_latent else predict


def sdisc_forward(sd, prefix, x, train=True, round_bf16=False):
    """SDisc.forward (models/mygannet.py:138-162)"""
    rb = round_bf16
    h = x
    for i in range(1, 7):
        h = net_conv(sd, f"{prefix}dconv{i}", h, (1, 3, 3), 0.01, train, rb)
        h = _r(F.avg_pool3d(h, (1, 2, 2)), rb)
    feat = h
    g = F.avg_pool3d(feat, (feat.shape[2], 1, 1), stride=1)
    cls = torch.sigmoid(F.linear(g.reshape(g.shape[0], -1), sd[prefix + "linear.weight"], sd[prefix + "linear.bias"]))
    return cls.squeeze(1), feat


def tdisc_forward(sd, prefix, x, train=True, round_bf16=False):
    """TDisc.forward (models/mygannet.py:180-196)"""
    rb = round_bf16
    h = x
    for i in range(1, 4):
        h = net_conv(sd, f"{prefix}dconv{i}", h, (3, 1, 1), 0.01, train, rb)
        h = _r(F.avg_pool3d(h, (2, 1, 1)), rb)
    feat = h
    g = F.avg_pool3d(feat, (1, feat.shape[3], feat.shape[4]), stride=1)
    cls = torch.sigmoid(F.linear(g.reshape(g.shape[0], -1), sd[prefix + "linear.weight"], sd[prefix + "linear.bias"]))
    return cls.squeeze(1), feat


def netd_forward(sd, x, y, train=True, round_bf16=False):
    """NetD.forward (models/mygannet.py:208-213)"""
    s_cls, s_feat = sdisc_forward(sd, "spatdisc.", x, train, round_bf16)
    t_cls, t_feat = tdisc_forward(sd, "tempdisc.", y, train, round_bf16)
    return s_cls, s_feat, t_cls, t_feat


def convlstm_cell(sd, prefix, x, h_cur, c_cur, round_bf16=False):
    """ConvLSTMCell.forward (models/convlstm.py:42-58): gates split in the order i, f, o, g."""
    w = sd[prefix + "conv.weight"]
    hid = w.shape[0] // 4
    comb = torch.cat([x, h_cur], dim=1)
    cc = F.conv2d(_r(comb, round_bf16), _r(w, round_bf16), sd.get(prefix + "conv.bias"),
                  padding=(w.shape[2] // 2, w.shape[3] // 2))
    cc_i, cc_f, cc_o, cc_g = torch.split(cc, hid, dim=1)
    i, f, o, g = torch.sigmoid(cc_i), torch.sigmoid(cc_f), torch.sigmoid(cc_o), torch.tanh(cc_g)
    c_next = f * c_cur + i * g
    return o * torch.tanh(c_next), c_next


def convlstm_unroll(sd, prefix, x_btchw, round_bf16=False):
    """ConvLSTM.forward for one layer, batch_first, zero initial state (models/convlstm.py:101-151)."""
    w = sd[prefix + "conv.weight"]
    hid = w.shape[0] // 4
    B, T, _, H, W = x_btchw.shape
    h = torch.zeros(B, hid, H, W)
    c = torch.zeros(B, hid, H, W)
    outs = []
    for t in range(T):
        h, c = convlstm_cell(sd, prefix, x_btchw[:, t], h, c, round_bf16)
        outs.append(h)
    return torch.stack(outs, dim=1), (h, c)


# ------------------------------------------------------------------------------------------------
# builder-defined compositions of reference modules (D1, D3, D5)
# ------------------------------------------------------------------------------------------------
def netg_lstm_forward(sd, x, train=True, dropout_masks=None, round_bf16=False, return_latent=False):
    """NetG with a one-layer bias-free ConvLSTM over the latent: ``ConvLSTM(...)(latent.transpose(1, 2))`` and
    back, the wrapping idiom of models/convlstm.py:199-201; weights under ``clstm.cell_list.0.``."""
    def bottleneck(latent):
        out, _ = convlstm_unroll(sd, "clstm.cell_list.0.", latent.transpose(1, 2), round_bf16)
        return _r(out.transpose(1, 2), round_bf16)
    return netg_forward(sd, x, train, dropout_masks, round_bf16=round_bf16, return_latent=return_latent,
                        bottleneck=bottleneck)


def encoder_forward(sd, prefix, x, train=True, round_bf16=False):
    """dconv1..dconv5 + AvgPool3d(2) between them (models/mygannet.py:57-71), weights under ``prefix``."""
    h = x
    for i in range(1, 5):
        h = _r(F.avg_pool3d(net_conv(sd, f"{prefix}dconv{i}", h, (3, 3, 3), 0.2, train, round_bf16), 2), round_bf16)
    return net_conv(sd, f"{prefix}dconv5", h, (3, 3, 3), 0.2, train, round_bf16)


def enc_dec_enc_forward(sd, x, train=True, dropout_masks=None, round_bf16=False):
    """(predict, latent_i, latent_o) of the enc-dec-enc composition (shape of models/ganomaly.py:160-175):
    NetG under ``netg.``, second encoder under ``encoder2.``, fed with gray2rgb(predict)."""
    sd_g = {k[len("netg."):]: v for k, v in sd.items() if k.startswith("netg.")}
    predict, latent_i = netg_forward(sd_g, x, train, dropout_masks, round_bf16=round_bf16, return_latent=True)
    for k, v in sd_g.items():      # running statistics updated in place on the views
        sd["netg." + k] = v
    latent_o = encoder_forward(sd, "encoder2.", gray2rgb(predict), train, round_bf16)
    return predict, latent_i, latent_o


def anomaly_scores(latent_i, latent_o):
    """models/ganomaly.py:372 ``torch.mean(torch.pow(latent_i - latent_o, 2), dim=1)`` for a (B, nz, 1, 1)
    latent, i.e. the mean over every non-batch dim of the 3-D latent."""
    return torch.pow(latent_i - latent_o, 2).flatten(1).mean(dim=1)


def minmax_scale(scores):
    """models/ganomaly.py:396"""
    return (scores - torch.min(scores)) / (torch.max(scores) - torch.min(scores))


def l1_loss(a, b):
    """nn.L1Loss() (models/ganomaly.py:438)"""
    return torch.mean(torch.abs(a - b))


# ------------------------------------------------------------------------------------------------
# STCNN (BASELINE config 4)
# ------------------------------------------------------------------------------------------------
def c2plus1d_block(sd, prefix, x, down_samp, train=True, dropout_mask=None, round_bf16=False):
    """C2plus1d_Block.forward (models/mystcnn.py:26-50). ``dropout_mask``: multiplier tensor replacing
    nn.Dropout on the shortcut input of the up-sampling blocks (None = no dropout)."""
    rb = round_bf16
    inp = x
    y = _r(F.conv3d(_r(x, rb), _r(sd[prefix + "spaceconv.weight"], rb), None, padding=(0, 1, 1)), rb)
    a = _r(F.relu(batch_norm(sd, prefix + "bn1", y, train)), rb)
    y = _r(F.conv3d(a, _r(sd[prefix + "pointwise.weight"], rb), None, padding=(1, 0, 0)), rb)
    h = _r(F.relu(batch_norm(sd, prefix + "bn2", y, train)), rb)
    w1, b1 = _r(sd[prefix + "conv.weight"], rb), sd[prefix + "conv.bias"]
    if down_samp:
        h = _r(F.avg_pool3d(h, 2), rb)
        inp = _r(F.avg_pool3d(_r(F.conv3d(_r(inp, rb), w1, b1), rb), 2), rb)
    else:
        h = _r(F.interpolate(h, scale_factor=2, mode="trilinear", align_corners=True), rb)
        if dropout_mask is not None:
            inp = _r(inp * dropout_mask, rb)
        inp = _r(F.interpolate(inp, scale_factor=2, mode="trilinear", align_corners=True), rb)
        inp = _r(F.conv3d(inp, w1, b1), rb)
    h = torch.cat([h, inp], dim=1)
    return _r(F.conv3d(h, _r(sd[prefix + "conv_last.weight"], rb), None, padding=1), rb)


def autoencoder_forward(sd, x, train=True, dropout_masks=None, round_bf16=False):
    """AutoEncoder.forward (models/mystcnn.py:69-88) -> predict (B,1,D,H,W)."""
    rb = round_bf16
    dm = dropout_masks if dropout_masks is not None else [None] * 4
    d1 = c2plus1d_block(sd, "down_sep1.", x, True, train, None, rb)
    d2 = c2plus1d_block(sd, "down_sep2.", d1, True, train, None, rb)
    d3 = c2plus1d_block(sd, "down_sep3.", d2, True, train, None, rb)
    d4 = c2plus1d_block(sd, "down_sep4.", d3, True, train, None, rb)
    u1 = c2plus1d_block(sd, "up_sep1.", d4, False, train, dm[0], rb)
    u2 = c2plus1d_block(sd, "up_sep2.", torch.cat([u1, d3], dim=1), False, train, dm[1], rb)
    u3 = c2plus1d_block(sd, "up_sep3.", torch.cat([u2, d2], dim=1), False, train, dm[2], rb)
    u4 = c2plus1d_block(sd, "up_sep4.", torch.cat([u3, d1], dim=1), False, train, dm[3], rb)
    return torch.sigmoid(F.conv3d(u4, _r(sd["conv_last.weight"], rb), None, padding=1))


# ------------------------------------------------------------------------------------------------
# host detours of MyGAN.test
# ------------------------------------------------------------------------------------------------
def threshold(data):
    """lib/utils.py:149-152"""
    return (data > 0.5).float()


def morphology_proc(video):
    """lib/utils.py:139-147 restated without cv2: ``cv2.morphologyEx(i, MORPH_OPEN, ones(5,5))`` is applied to
    each clip's (D, H, W) array, which OpenCV reads as a D x H image with W channels, so the 5x5 opening
    (erode then dilate, border pixels never win) runs in the (D, H) plane for every w. Pinned against cv2
    itself by tests/golden/eval_small.pt."""
    B, C, D, H, W = video.shape
    planes = video.permute(0, 1, 4, 2, 3).reshape(B * C * W, 1, D, H)
    eroded = -F.max_pool2d(-planes, 5, stride=1, padding=2)       # max_pool pads with -inf
    opened = F.max_pool2d(eroded, 5, stride=1, padding=2)
    return opened.reshape(B, C, W, D, H).permute(0, 1, 3, 4, 2).contiguous()


def evaluate(labels, scores, metric):
    """lib/evaluate.py:14-91 without the plotting / CSV side effects: the same sklearn calls
    (``roc_curve`` + ``auc``; ``precision_recall_curve`` + ``auc``; ``f1_score`` after the 0.20 binarisation)."""
    from sklearn.metrics import roc_curve, auc, f1_score, precision_recall_curve
    import numpy as np
    labels, scores = np.asarray(labels), np.asarray(scores, dtype=np.float64).copy()
    if metric == "roc":
        fpr, tpr, _ = roc_curve(labels, scores)
        return float(auc(fpr, tpr))
    if metric == "pr":
        precision, recall, _ = precision_recall_curve(labels, scores)
        return float(auc(recall, precision))
    if metric == "f1_score":
        scores[scores >= 0.20] = 1
        scores[scores < 0.20] = 0
        return float(f1_score(labels, scores))
    raise NotImplementedError("Check the evaluation metric.")


# ------------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------------
def l2_loss(a, b):
    """lib/utils.py:59-63"""
    return torch.mean(torch.pow(a - b, 2))


def weighted_bce(p, t, pos_weight=2):
    """lib/utils.py:65-71 (pos_weight multiplies the (1 - target) term; the clamp's upper bound
    1 - 1e-8 is 1.0 in fp32, as in the reference)."""
    p = torch.clamp(p, min=1e-8, max=1 - 1e-8)
    loss = t * torch.log(p) + pos_weight * (1 - t) * torch.log(1 - p)
    return torch.neg(torch.mean(loss))


def gray2rgb(v):
    """lib/utils.py:91-92"""
    return torch.cat([v, v, v], dim=1)


# ------------------------------------------------------------------------------------------------
# the train step
# ------------------------------------------------------------------------------------------------
class OracleTrainer:
    """MyGAN.optimize_params restated on CPU (models/mygannet.py:275-366).

    Holds fp32 copies of both state_dicts; parameters become autograd leaves, buffers are updated
    in place. Optical flow is an input (the reference computes it on the host with cv2 Farneback,
    lib/utils.py:94-129 -- outside the hot path). Adam follows
    ``optim.Adam(lr, betas=(beta1, 0.999))`` (models/mygannet.py:270-273) via torch.optim.Adam on
    the leaves. Dead work is not skipped: ``err_g`` includes the adversarial term and is
    back-propagated with ``retain_graph`` exactly like the reference (its D gradients are wiped by
    ``optimizer_d.zero_grad()``)."""

    def __init__(self, sd_g, sd_d, lr=2e-5, beta1=0.5, w_adv=1, w_con=10, round_bf16=False, netg_fn=None):
        def split(sd):
            params, bufs = {}, {}
            for k, v in sd.items():
                v = v.detach().clone().cpu()
                if k.endswith(("running_mean", "running_var", "num_batches_tracked")):
                    bufs[k] = v
                else:
                    params[k] = v.float().requires_grad_(True)
            return params, bufs

        self.pg, self.bg = split(sd_g)
        self.pd, self.bd = split(sd_d)
        self.opt_g = torch.optim.Adam(list(self.pg.values()), lr=lr, betas=(beta1, 0.999))
        self.opt_d = torch.optim.Adam(list(self.pd.values()), lr=lr, betas=(beta1, 0.999))
        self.w_adv, self.w_con = w_adv, w_con
        self.rb = round_bf16
        self.netg_fn = netg_fn or netg_forward      # netg_lstm_forward for the config-3 composition
        self.bce = torch.nn.BCELoss()

    def sd_g(self):
        return {**self.pg, **self.bg}

    def sd_d(self):
        return {**self.pd, **self.bd}

    def step(self, inp, gt, gt_flow, pre_flow, dropout_masks=None):
        sd_g, sd_d = self.sd_g(), self.sd_d()
        # forward_g (:275-276)
        predict = self.netg_fn(sd_g, inp, True, dropout_masks, round_bf16=self.rb)
        # forward_d (:278-286): every D input is detached
        pre_3ch, gt_3ch = gray2rgb(predict.detach()), gray2rgb(gt)
        s_pr, s_fr, t_pr, t_fr = netd_forward(sd_d, gt_3ch, gt_flow, True, self.rb)
        s_pf, s_ff, t_pf, t_ff = netd_forward(sd_d, pre_3ch, pre_flow, True, self.rb)
        # backward_g (:305-320)
        self.opt_g.zero_grad()
        err_g_adv_s, err_g_adv_t = l2_loss(s_fr, s_ff), l2_loss(t_fr, t_ff)
        err_g_adv = err_g_adv_s + err_g_adv_t
        err_g_con = weighted_bce(predict, gt)
        err_g = err_g_adv * self.w_adv + err_g_con * self.w_con
        err_g.backward(retain_graph=True)
        self.opt_g.step()
        # backward_d (:323-344)
        self.opt_d.zero_grad()
        ones, zeros = torch.ones_like(s_pr), torch.zeros_like(s_pf)
        e_rs, e_rt = self.bce(s_pr, ones), self.bce(t_pr, ones)
        e_fs, e_ft = self.bce(s_pf, zeros), self.bce(t_pf, zeros)
        err_d_real, err_d_fake = (e_rs + e_rt) * 0.5, (e_fs + e_ft) * 0.5
        err_d = (err_d_real + err_d_fake) * 0.5
        err_d.backward()
        self.opt_d.step()
        return {
            "g/err_g": err_g.item(), "g/err_g_adv": err_g_adv.item(), "g/err_g_adv_s": err_g_adv_s.item(),
            "g/err_g_adv_t": err_g_adv_t.item(), "g/err_g_con": err_g_con.item(),
            "d/err_d_real_s": e_rs.item(), "d/err_d_real_t": e_rt.item(), "d/err_d_fake_s": e_fs.item(),
            "d/err_d_fake_t": e_ft.item(), "d/err_d_real": err_d_real.item(), "d/err_d_fake": err_d_fake.item(),
            "d/err_d": err_d.item(),
        }, predict.detach()


def synthetic_batch(batch, nfr, isize, seed=0):
    """Synthetic inputs of SURVEY.md section 8d: clips in [-1,1], sparse binary mask, random flows."""
    g = torch.Generator().manual_seed(seed)
    inp = torch.rand(batch, 3, nfr, isize, isize, generator=g) * 2 - 1
    gt = (torch.rand(batch, 1, nfr, isize, isize, generator=g) > 0.9).float()
    gt_flow = torch.rand(batch, 3, nfr, isize, isize, generator=g) * 2 - 1
    pre_flow = torch.rand(batch, 3, nfr, isize, isize, generator=g) * 2 - 1
    return inp, gt, gt_flow, pre_flow
