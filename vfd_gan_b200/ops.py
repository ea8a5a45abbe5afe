"""torch.library registration and autograd wiring of the libvfd_b200 kernels.

Everything here works on *channels-last bf16* activations: a torch tensor of shape
``[N, D, H, W, C]`` with ``stride(-1) == 1`` and ``C % 8 == 0``; the voxel pitch ``stride(3)`` may
exceed ``C`` (channel slice of a concat buffer). The ``nn.Module`` surface in ``spatiotempconv.py``
/ ``mygannet.py`` / ``convlstm.py`` converts from / to the reference's fp32 NCDHW at its boundary.

Ops are registered in the ``vfd_b200`` torch.library namespace with CUDA implementations that call
the C-ABI through ctypes (``_lib.call``); autograd is provided by the ``*Fn`` classes below.
"""
import os

import torch

from . import _lib

_NS = "vfd_b200"
_deflib = torch.library.Library(_NS, "DEF")
_registered = {}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _define(schema, fn):
    """Define ``vfd_b200::<name>`` and register ``fn`` as its CUDA implementation."""
    name = schema.split("(")[0]
    _deflib.define(schema)
    _deflib.impl(name, fn, "CUDA")
    _registered[name] = fn
    return getattr(getattr(torch.ops, _NS), name)


def round_up(x, m):
    return (x + m - 1) // m * m


def pick_kc(cin_p):
    """Channel block per MMA K-slab (16/32/64): the largest one whose padded K stays within 10 %
    of the tightest padding."""
    import os
    if os.environ.get("VFD_FORCE_KC64") and cin_p > 32:
        return 64
    best = min(round_up(cin_p, k) for k in (16, 32, 64))
    for kc in (64, 32, 16):
        if round_up(cin_p, kc) <= 1.1 * best:
            return kc
    return 16


def _ld(t):
    """Voxel pitch (elements) of a channels-last tensor; singleton voxel dims carry no stride info."""
    if t.is_contiguous():
        return t.shape[-1]
    N, D, H, W, C = t.shape
    for dim, inner in ((3, 1), (2, W), (1, H * W), (0, D * H * W)):
        if t.shape[dim] > 1:
            return t.stride(dim) // inner
    return C


def _is_cl(t):
    if t.dim() != 5 or t.shape[-1] % 8 or (t.numel() and t.stride(-1) != 1):
        return False
    N, D, H, W, C = t.shape
    ld = _ld(t)
    if ld < C or ld % 8 or t.data_ptr() % 16:
        return False
    return not ((W > 1 and t.stride(3) != ld) or (H > 1 and t.stride(2) != W * ld) or
                (D > 1 and t.stride(1) != H * W * ld) or (N > 1 and t.stride(0) != D * H * W * ld))


def _check_cl(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: vfd_gan_b200 has no CPU path; tensor is on {t.device}")
    if t.device.index != torch.cuda.current_device():
        # the C-ABI launches on the current device's stream: refuse tensors of another device instead of
        # launching over peer access without ordering
        raise RuntimeError(f"{what}: tensor is on {t.device} but the current CUDA device is "
                           f"cuda:{torch.cuda.current_device()}; wrap the call in torch.cuda.device(...)")
    if t.dtype != torch.bfloat16 or not _is_cl(t):
        raise RuntimeError(f"{what}: expected a channels-last bf16 [N,D,H,W,C] tensor (C % 8 == 0, dense voxel "
                           f"dims, 16-byte aligned), got {tuple(t.shape)} {t.dtype} strides {t.stride()}")
    N, D, H, W, C = t.shape
    return N, D, H, W, C, _ld(t)


def as_cl_grad(g):
    """Bring an incoming gradient to channels-last bf16 (autograd may hand us fp32 or expanded views)."""
    if g is None:
        return None
    if g.dtype != torch.bfloat16:
        g = g.to(torch.bfloat16)
    return g if _is_cl(g) else g.contiguous()


# ------------------------------------------------------------------------------------------------
# raw ops (mutating "out" style; registered with torch.library)
# ------------------------------------------------------------------------------------------------
def _conv3d_fwd(x, w_packed, bias, out, stats, kd, kh, kw, kc, out_cols, direct):
    N, D, H, W, C, ld = _check_cl(x, "conv3d_fwd input")
    rows, taps, cin_k = w_packed.shape
    out_fp32 = 1 if out.dtype == torch.float32 else 0
    head = [x.data_ptr(), ld, C, w_packed.data_ptr(), rows, cin_k, _ptr(bias), out.data_ptr(), _ld(out), out_cols,
            out_fp32]
    if direct:
        _lib.call_debug("vfd_conv3d_fwd_direct", *head, N, D, H, W, kd, kh, kw, _stream())
        if stats is not None:   # the CUDA-core cross-check path has no fused statistics
            _lib.call("vfd_bn_stats", out.data_ptr(), _ld(out), out.shape[-1], N * D * H * W, stats.data_ptr(),
                      _stream())
    else:
        _lib.call("vfd_conv3d_fwd", *head, _ptr(stats), 0 if stats is None else stats.numel() // 2, N, D, H, W, kd,
                  kh, kw, kc, _stream())


def _conv3d_fwd_narrow(x, w_packed, bias, out):
    N, D, H, W, C, ld = _check_cl(x, "conv3d_fwd_narrow input")
    _lib.call("vfd_conv3d_fwd_narrow", x.data_ptr(), ld, C, w_packed.data_ptr(), w_packed.shape[2], _ptr(bias),
              out.data_ptr(), _ld(out), out.shape[-1], N, D, H, W, _stream())


def _convlstm_step_fwd(comb, w_perm, bias_perm, c_cur, c_next, h_out, act, kh, kw, kc):
    N, D, H, W, C, ld = _check_cl(comb, "convlstm_step_fwd input")
    hid = c_cur.shape[-1]
    _lib.call("vfd_convlstm_step_fwd", comb.data_ptr(), ld, C, w_perm.data_ptr(), w_perm.shape[2], _ptr(bias_perm),
              c_cur.data_ptr(), hid, c_next.data_ptr(), h_out.data_ptr(), _ld(h_out), _ptr(act), N, H, W, kh, kw, kc,
              _stream())


def _conv3d_dgrad_narrow(g, w_dgrad, gx):
    N, D, H, W, _, ld = _check_cl(g, "conv3d_dgrad_narrow gradient")
    _lib.call("vfd_conv3d_dgrad_narrow", g.data_ptr(), ld, w_dgrad.data_ptr(), w_dgrad.shape[0], w_dgrad.shape[2],
              gx.data_ptr(), _ld(gx), N, D, H, W, _stream())


def _conv3d_wgrad_narrow(g, x, acc):
    N, D, H, W, _, g_ld = _check_cl(g, "conv3d_wgrad_narrow gradient")
    _, _, _, _, _, x_ld = _check_cl(x, "conv3d_wgrad_narrow input")
    _lib.call("vfd_conv3d_wgrad_narrow", g.data_ptr(), g_ld, x.data_ptr(), x_ld, acc.data_ptr(), acc.shape[-1], N, D, H, W,
              _stream())


def _conv3d_wgrad_first(dy, cout, x, acc):
    N, D, H, W, _, dy_ld = _check_cl(dy, "conv3d_wgrad_first gradient")
    _, _, _, _, _, x_ld = _check_cl(x, "conv3d_wgrad_first input")
    _lib.call("vfd_conv3d_wgrad_first", dy.data_ptr(), dy_ld, cout, x.data_ptr(), x_ld, acc.data_ptr(), acc.shape[-1],
              N, D, H, W, _stream())


def _conv3d_wgrad(dy, cout, x, cin, acc, kd, kh, kw, direct, layout=0):
    N, D, H, W, _, dy_ld = _check_cl(dy, "conv3d_wgrad dy")
    _, _, _, _, _, x_ld = _check_cl(x, "conv3d_wgrad x")
    if layout == 1:
        taps, co_pad, ci_pad = acc.shape
    else:
        taps, ci_pad, co_pad = acc.shape
    if direct:
        _lib.call_debug("vfd_conv3d_wgrad_direct", dy.data_ptr(), dy_ld, cout, x.data_ptr(), x_ld, cin, acc.data_ptr(),
                  co_pad, ci_pad, N, D, H, W, kd, kh, kw, _stream())
    elif DETERMINISTIC:
        need = int(_lib.lib().vfd_conv3d_wgrad_det_workspace(cout, cin, co_pad, ci_pad, layout, N, D, H, W, kd, kh, kw))
        if need < 0:
            raise RuntimeError("conv3d_wgrad_det: " + _lib.lib().vfd_last_error().decode())
        ws = torch.empty(max(need, 16), dtype=torch.uint8, device=dy.device)
        _lib.call("vfd_conv3d_wgrad_det", dy.data_ptr(), dy_ld, cout, x.data_ptr(), x_ld, cin, acc.data_ptr(), co_pad,
                  ci_pad, layout, N, D, H, W, kd, kh, kw, ws.data_ptr(), ws.numel(), _stream())
    else:
        _lib.call("vfd_conv3d_wgrad", dy.data_ptr(), dy_ld, cout, x.data_ptr(), x_ld, cin, acc.data_ptr(), co_pad,
                  ci_pad, layout, N, D, H, W, kd, kh, kw, _stream())


def _conv3d_wgrad_thin(dy, cout, x, cin, acc, fold=0, cs=0, kd=1, kh=1, kw=1):
    N, D, H, W, _, dy_ld = _check_cl(dy, "conv3d_wgrad_thin dy")
    _, _, _, _, _, x_ld = _check_cl(x, "conv3d_wgrad_thin x")
    _, ci_pad, co_pad = acc.shape
    if DETERMINISTIC:
        need = int(_lib.lib().vfd_conv3d_wgrad_thin_det_workspace(cout, cin, fold, cs, N, D, H, W, kd, kh, kw))
        if need < 0:
            raise RuntimeError("conv3d_wgrad_thin_det: " + _lib.lib().vfd_last_error().decode())
        ws = torch.empty(max(need, 16), dtype=torch.uint8, device=dy.device)
        _lib.call("vfd_conv3d_wgrad_thin_det", dy.data_ptr(), dy_ld, cout, x.data_ptr(), x_ld, cin, acc.data_ptr(),
                  co_pad, ci_pad, fold, cs, N, D, H, W, kd, kh, kw, ws.data_ptr(), ws.numel(), _stream())
        return
    _lib.call("vfd_conv3d_wgrad_thin", dy.data_ptr(), dy_ld, cout, x.data_ptr(), x_ld, cin, acc.data_ptr(), co_pad,
              ci_pad, fold, cs, N, D, H, W, kd, kh, kw, _stream())


def wgrad_layout(cout, cin, kd, kh, kw, H, W):
    """Accumulator layout the tcgen05 wgrad wants: 0 = [tap][ci][co], 1 = [tap][co][ci] (swapped GEMM roles)."""
    return int(_lib.lib().vfd_conv3d_wgrad_layout(cout, cin, kd, kh, kw, H, W))


def _pack_ncdhw(src, dst, C, replicate):
    N, Csrc = src.shape[0], src.shape[1]
    S = src[0, 0].numel() if N > 0 else 0
    _lib.call("vfd_pack_ncdhw", src.data_ptr(), dst.data_ptr(), N, Csrc, S, C, _ld(dst), dst.shape[-1],
              1 if replicate else 0, _stream())


def _unpack_ncdhw(src, dst):
    N, C = dst.shape[0], dst.shape[1]
    S = dst[0, 0].numel() if N > 0 else 0
    _lib.call("vfd_unpack_ncdhw", src.data_ptr(), 1 if src.dtype == torch.float32 else 0, dst.data_ptr(), N, C,
              S, _ld(src), _stream())


def _pack_weight(w, wp, mode):
    cout, cin = w.shape[0], w.shape[1]
    rows, taps, ck = wp.shape
    _lib.call("vfd_pack_weight", w.data_ptr(), wp.data_ptr(), cout, cin, taps, rows, ck, mode, _stream())


def _unpack_wgrad(acc, gw, accumulate=False):
    taps, ci_pad, co_pad = acc.shape
    _lib.call("vfd_unpack_wgrad", acc.data_ptr(), gw.data_ptr(), gw.shape[0], gw.shape[1], taps, co_pad, ci_pad,
              1 if accumulate else 0, _stream())


def _bn_prepare(y, sums, cvalid, pre_bias, gamma, beta, running_mean, running_var, momentum, eps, train, mean,
                invstd, scale, shift, stats_ready):
    N, D, H, W, C, ld = _check_cl(y, "bn input")
    V = N * D * H * W
    if train and not stats_ready:
        _lib.call("vfd_bn_stats", y.data_ptr(), ld, C, V, sums.data_ptr(), _stream())
    _lib.call("vfd_bn_finalize", sums.data_ptr(), C, cvalid, V, _ptr(pre_bias), gamma.data_ptr(), beta.data_ptr(),
              _ptr(running_mean), _ptr(running_var), momentum, eps, 1 if train else 0, mean.data_ptr(),
              invstd.data_ptr(), scale.data_ptr(), shift.data_ptr(), _stream())


def _bn_act_fwd(y, scale, shift, slope, out_full, out_pool, pd, ph, pw, drop_p, seed, seed_dev=None):
    N, D, H, W, C, ld = _check_cl(y, "bn_act_fwd input")
    _lib.call("vfd_bn_act_fwd", y.data_ptr(), ld, N, D, H, W, C, scale.data_ptr(), shift.data_ptr(), slope,
              _ptr(out_full), 0 if out_full is None else _ld(out_full), _ptr(out_pool),
              0 if out_pool is None else _ld(out_pool), pd, ph, pw, drop_p, seed, _ptr(seed_dev), _stream())


def _bn_act_bwd(y, cvalid, mean, invstd, scale, shift, slope, g_full, g_pool, pd, ph, pw, drop_p, seed, train,
                sums, c1, c2, dgamma, dbeta, dy, seed_dev=None, pool_bcast=False, accumulate=False):
    N, D, H, W, C, ld = _check_cl(y, "bn_act_bwd input")
    _lib.call("vfd_bn_act_bwd", y.data_ptr(), ld, N, D, H, W, C, cvalid, mean.data_ptr(), invstd.data_ptr(),
              scale.data_ptr(), shift.data_ptr(), slope, _ptr(g_full), 0 if g_full is None else _ld(g_full),
              _ptr(g_pool), 0 if g_pool is None else _ld(g_pool), pd, ph, pw, drop_p, seed, _ptr(seed_dev),
              (1 if train else 0) | (2 if pool_bcast else 0) | (4 if accumulate else 0), sums.data_ptr(), c1.data_ptr(), c2.data_ptr(), dgamma.data_ptr(),
              dbeta.data_ptr(), dy.data_ptr(), _ld(dy), bn_ticket(y.device).data_ptr() if BN_TICKET else None, _stream())


def _tap_gather(src, cs, dst, kd, kh, kw, sign):
    N, D, H, W, _, ld = _check_cl(src, "tap_gather source")
    _lib.call("vfd_tap_gather", src.data_ptr(), ld, cs, dst.data_ptr(), _ld(dst), dst.shape[-1], N, D, H, W, kd, kh,
              kw, sign, _stream())


def _channel_sum(x, out):
    N, D, H, W, C, ld = _check_cl(x, "channel_sum input")
    if out.dtype != torch.float64:
        raise RuntimeError("channel_sum accumulates in float64")
    _lib.call("vfd_channel_sum", x.data_ptr(), ld, C, N * D * H * W, out.data_ptr(), _stream())


def _upsample2x_fwd(x, out):
    N, D, H, W, C, ld = _check_cl(x, "upsample2x_fwd input")
    _lib.call("vfd_upsample2x_fwd", x.data_ptr(), ld, N, D, H, W, C, out.data_ptr(), _ld(out), _stream())


def _upsample2x_bwd(gout, gx):
    N, D, H, W, C, ld = _check_cl(gx, "upsample2x_bwd output")
    ws = torch.empty(N * 2 * D * H * W * C * 2 * 3, dtype=torch.uint8, device=gx.device)   # two bf16 temporaries
    _lib.call("vfd_upsample2x_bwd", gout.data_ptr(), _ld(gout), N, D, H, W, C, gx.data_ptr(), ld, ws.data_ptr(),
              ws.numel(), _stream())


def _sigmoid_head_fwd(logits, predict):
    _lib.call("vfd_sigmoid_head_fwd", logits.data_ptr(), _ld(logits), predict.numel(), predict.data_ptr(),
              _stream())


def _sigmoid_head_bwd(gpred, predict, dlogit):
    _lib.call("vfd_sigmoid_head_bwd", gpred.data_ptr(), predict.data_ptr(), predict.numel(), dlogit.data_ptr(),
              _stream())


def _weighted_bce(predict, target, pos_weight, grad_scale, loss_sum, gpred):
    _lib.call("vfd_weighted_bce", predict.data_ptr(), target.data_ptr(), predict.numel(), pos_weight, grad_scale,
              loss_sum.data_ptr(), _ptr(gpred), _stream())


def _sqdiff(a, b, out):
    N, D, H, W, C, ld = _check_cl(a, "sqdiff a")
    _check_cl(b, "sqdiff b")
    _lib.call("vfd_sqdiff", a.data_ptr(), ld, b.data_ptr(), _ld(b), C, N * D * H * W, out.data_ptr(),
              _stream())


def _convlstm_cell_fwd(gates, c_cur, h_next, c_next, act):
    hid = c_cur.shape[-1]
    _lib.call("vfd_convlstm_cell_fwd", gates.data_ptr(), _ld(gates), c_cur.data_ptr(), hid,
              c_cur.numel() // hid, h_next.data_ptr(), c_next.data_ptr(), _ptr(act), _stream())


def _convlstm_cell_bwd(act, c_cur, c_next, dh, dc_in, dgates, dc_cur):
    hid = c_cur.shape[-1]
    _lib.call("vfd_convlstm_cell_bwd", act.data_ptr(), c_cur.data_ptr(), c_next.data_ptr(), _ptr(dh), _ptr(dc_in),
              hid, c_cur.numel() // hid, dgates.data_ptr(), _ld(dgates), dc_cur.data_ptr(), _stream())


def _latent_score(a, b, per_clip):
    N, D, H, W, C, ld = _check_cl(a, "latent_score a")
    _check_cl(b, "latent_score b")
    if tuple(b.shape) != tuple(a.shape):
        raise RuntimeError(f"latent_score: shapes differ, {tuple(a.shape)} vs {tuple(b.shape)}")
    _lib.call("vfd_latent_score", a.data_ptr(), ld, b.data_ptr(), _ld(b), C, D * H * W, N, per_clip.data_ptr(),
              _stream())


def _sqdiff_bwd(a, b, gscale, scale, ga, gb):
    N, D, H, W, C, ld = _check_cl(a, "sqdiff_bwd a")
    _check_cl(b, "sqdiff_bwd b")
    _lib.call("vfd_sqdiff_bwd", a.data_ptr(), ld, b.data_ptr(), _ld(b), C, N * D * H * W, _ptr(gscale), scale,
              _ptr(ga), 0 if ga is None else _ld(ga), _ptr(gb), 0 if gb is None else _ld(gb), _stream())


def _l1_loss(a, b, grad_scale, out, ga):
    _lib.call("vfd_l1_loss", a.data_ptr(), b.data_ptr(), a.numel(), grad_scale, out.data_ptr(), _ptr(ga), _stream())


def _bce_loss(p, t, grad_scale, out, gp):
    _lib.call("vfd_bce_loss", p.data_ptr(), t.data_ptr(), p.numel(), grad_scale, out.data_ptr(), _ptr(gp), _stream())


def _score_finalize(per_clip, inv_count, scores, minmax):
    _lib.call("vfd_score_finalize", per_clip.data_ptr(), per_clip.numel(), inv_count, scores.data_ptr(),
              _ptr(minmax), _stream())


def _score_scale(scores, minmax, out):
    _lib.call("vfd_score_scale", scores.data_ptr(), scores.numel(), minmax.data_ptr(), out.data_ptr(), _stream())


def _threshold_open(predict, thr, t_out, m_out):
    N, D, H, W = predict.shape[0], predict.shape[-3], predict.shape[-2], predict.shape[-1]
    _lib.call("vfd_threshold_open", predict.data_ptr(), N, D, H, W, thr, _ptr(t_out), m_out.data_ptr(), _stream())


def _confusion_counts(labels, scores, thr, counts):
    _lib.call("vfd_confusion_counts", labels.data_ptr(), scores.data_ptr(), labels.numel(), thr, counts.data_ptr(),
              _stream())


def _video_to_flow(video, out, raw, ws):
    B, _, D, H, W = video.shape
    _lib.call("vfd_video_to_flow", video.data_ptr(), B, D, H, W, out.data_ptr(), _ptr(raw), ws.data_ptr(), ws.numel(),
              _stream())
    # kernels launched by the call (the counter above added one): 2 fills, frame min/max, grey, 3 per pyramid
    # level for the polynomial expansion, per level 1 + 3 + 2 flow kernels (+ 1 upsample below the coarsest),
    # field min/max, encode
    levels, scale = 0, 1.0
    while levels < 3:
        scale *= 0.5
        if W * scale < 32 or H * scale < 32:
            break
        levels += 1
    _lib.KERNEL_LAUNCHES += 2 + 1 + 1 + 3 * (levels + 1) + 6 * (levels + 1) + levels + 2 - 1


def _roc_auc(scores, labels, out):
    _lib.call("vfd_roc_auc", scores.data_ptr(), labels.data_ptr(), scores.numel(), out.data_ptr(), _stream())


def _roc_auc_large(scores, labels, out, ws):
    _lib.call("vfd_roc_auc_large", scores.data_ptr(), labels.data_ptr(), scores.numel(), out.data_ptr(),
              ws.data_ptr(), ws.numel(), _stream())


def _resize_frames_u8(src, dst, ws):
    n, hin, win, c = src.shape
    _lib.call("vfd_resize_frames_u8", src.data_ptr(), n, hin, win, c, dst.data_ptr(), dst.shape[1], dst.shape[2],
              ws.data_ptr(), ws.numel(), _stream())


def _frames_to_clip(frames, out, pm1):
    b, t, h, w, c = frames.shape
    _lib.call("vfd_frames_to_clip", frames.data_ptr(), b, t, h, w, c, out.shape[1], int(pm1), out.data_ptr(),
              _stream())


conv3d_fwd = _define(
    "conv3d_fwd(Tensor x, Tensor w_packed, Tensor? bias, Tensor(a!) out, Tensor(b!)? stats, int kd, int kh, int kw, "
    "int kc, int out_cols, bool direct) -> ()", _conv3d_fwd)
conv3d_fwd_narrow = _define("conv3d_fwd_narrow(Tensor x, Tensor w_packed, Tensor? bias, Tensor(a!) out) -> ()",
                            _conv3d_fwd_narrow)
convlstm_step_fwd = _define(
    "convlstm_step_fwd(Tensor comb, Tensor w_perm, Tensor? bias_perm, Tensor c_cur, Tensor(a!) c_next, Tensor(b!) h_out, "
    "Tensor(c!)? act, int kh, int kw, int kc) -> ()", _convlstm_step_fwd)
conv3d_dgrad_narrow = _define("conv3d_dgrad_narrow(Tensor g, Tensor w_dgrad, Tensor(a!) gx) -> ()", _conv3d_dgrad_narrow)
conv3d_wgrad_narrow = _define("conv3d_wgrad_narrow(Tensor g, Tensor x, Tensor(a!) acc) -> ()", _conv3d_wgrad_narrow)
conv3d_wgrad_first = _define("conv3d_wgrad_first(Tensor dy, int cout, Tensor x, Tensor(a!) acc) -> ()",
                             _conv3d_wgrad_first)
conv3d_wgrad = _define(
    "conv3d_wgrad(Tensor dy, int cout, Tensor x, int cin, Tensor(a!) acc, int kd, int kh, int kw, bool direct, "
    "int layout=0) -> ()", _conv3d_wgrad)
conv3d_wgrad_thin = _define(
    "conv3d_wgrad_thin(Tensor dy, int cout, Tensor x, int cin, Tensor(a!) acc, int fold=0, int cs=0, int kd=1, "
    "int kh=1, int kw=1) -> ()", _conv3d_wgrad_thin)
pack_ncdhw = _define("pack_ncdhw(Tensor src, Tensor(a!) dst, int C, bool replicate) -> ()", _pack_ncdhw)
unpack_ncdhw = _define("unpack_ncdhw(Tensor src, Tensor(a!) dst) -> ()", _unpack_ncdhw)
pack_weight = _define("pack_weight(Tensor w, Tensor(a!) wp, int mode) -> ()", _pack_weight)
unpack_wgrad = _define("unpack_wgrad(Tensor acc, Tensor(a!) gw, bool accumulate=False) -> ()", _unpack_wgrad)
bn_prepare = _define(
    "bn_prepare(Tensor y, Tensor(a!) sums, int cvalid, Tensor? pre_bias, Tensor gamma, Tensor beta, "
    "Tensor(b!)? running_mean, "
    "Tensor(c!)? running_var, float momentum, float eps, bool train, Tensor(d!) mean, Tensor(e!) invstd, "
    "Tensor(f!) scale, Tensor(g!) shift, bool stats_ready) -> ()", _bn_prepare)
bn_act_fwd = _define(
    "bn_act_fwd(Tensor y, Tensor scale, Tensor shift, float slope, Tensor(a!)? out_full, Tensor(b!)? out_pool, "
    "int pd, int ph, int pw, float drop_p, int seed, Tensor? seed_dev=None) -> ()", _bn_act_fwd)
bn_act_bwd = _define(
    "bn_act_bwd(Tensor y, int cvalid, Tensor mean, Tensor invstd, Tensor scale, Tensor shift, float slope, "
    "Tensor? g_full, Tensor? g_pool, int pd, int ph, int pw, float drop_p, int seed, bool train, "
    "Tensor(a!) sums, Tensor(b!) c1, Tensor(c!) c2, Tensor(d!) dgamma, Tensor(e!) dbeta, Tensor(f!) dy, "
    "Tensor? seed_dev=None, bool pool_bcast=False, bool accumulate=False) -> ()", _bn_act_bwd)
tap_gather = _define("tap_gather(Tensor src, int cs, Tensor(a!) dst, int kd, int kh, int kw, int sign) -> ()",
                     _tap_gather)
channel_sum = _define("channel_sum(Tensor x, Tensor(a!) out) -> ()", _channel_sum)
upsample2x_fwd = _define("upsample2x_fwd(Tensor x, Tensor(a!) out) -> ()", _upsample2x_fwd)
upsample2x_bwd = _define("upsample2x_bwd(Tensor gout, Tensor(a!) gx) -> ()", _upsample2x_bwd)
sigmoid_head_fwd = _define("sigmoid_head_fwd(Tensor logits, Tensor(a!) predict) -> ()", _sigmoid_head_fwd)
sigmoid_head_bwd = _define("sigmoid_head_bwd(Tensor gpred, Tensor predict, Tensor(a!) dlogit) -> ()",
                           _sigmoid_head_bwd)
weighted_bce_op = _define(
    "weighted_bce(Tensor predict, Tensor target, float pos_weight, float grad_scale, Tensor(a!) loss_sum, "
    "Tensor(b!)? gpred) -> ()", _weighted_bce)
sqdiff = _define("sqdiff(Tensor a, Tensor b, Tensor(a!) out) -> ()", _sqdiff)
convlstm_cell_fwd = _define(
    "convlstm_cell_fwd(Tensor gates, Tensor c_cur, Tensor(a!) h_next, Tensor(b!) c_next, Tensor(c!)? act) -> ()",
    _convlstm_cell_fwd)
convlstm_cell_bwd = _define(
    "convlstm_cell_bwd(Tensor act, Tensor c_cur, Tensor c_next, Tensor? dh, Tensor? dc_in, Tensor(a!) dgates, "
    "Tensor(b!) dc_cur) -> ()", _convlstm_cell_bwd)

latent_score = _define("latent_score(Tensor a, Tensor b, Tensor(a!) per_clip) -> ()", _latent_score)
sqdiff_bwd = _define("sqdiff_bwd(Tensor a, Tensor b, Tensor? gscale, float scale, Tensor(a!)? ga, Tensor(b!)? gb) -> ()",
                     _sqdiff_bwd)
l1_loss_op = _define("l1_loss(Tensor a, Tensor b, float grad_scale, Tensor(a!) out, Tensor(b!)? ga) -> ()", _l1_loss)
bce_loss_op = _define("bce_loss(Tensor p, Tensor t, float grad_scale, Tensor(a!) out, Tensor(b!)? gp) -> ()", _bce_loss)
score_finalize = _define("score_finalize(Tensor per_clip, float inv_count, Tensor(a!) scores, Tensor(b!)? minmax) -> ()",
                         _score_finalize)
score_scale = _define("score_scale(Tensor scores, Tensor minmax, Tensor(a!) out) -> ()", _score_scale)
threshold_open_op = _define("threshold_open(Tensor predict, float thr, Tensor(a!)? t_out, Tensor(b!) m_out) -> ()",
                            _threshold_open)
confusion_counts_op = _define("confusion_counts(Tensor labels, Tensor scores, float thr, Tensor(a!) counts) -> ()",
                              _confusion_counts)
video_to_flow_op = _define("video_to_flow(Tensor video, Tensor(a!) out, Tensor(b!)? raw, Tensor(c!) ws) -> ()",
                           _video_to_flow)
roc_auc_op = _define("roc_auc(Tensor scores, Tensor labels, Tensor(a!) out) -> ()", _roc_auc)
resize_frames_u8_op = _define("resize_frames_u8(Tensor src, Tensor(a!) dst, Tensor(b!) ws) -> ()", _resize_frames_u8)
frames_to_clip_op = _define("frames_to_clip(Tensor frames, Tensor(a!) out, bool pm1) -> ()", _frames_to_clip)
roc_auc_large_op = _define("roc_auc_large(Tensor scores, Tensor labels, Tensor(a!) out, Tensor(b!) ws) -> ()",
                           _roc_auc_large)


# ------------------------------------------------------------------------------------------------
# helpers shared by the autograd functions
# ------------------------------------------------------------------------------------------------
# Weight gradients through per-split partial accumulators and an ordered second pass instead of fp32 atomics
# (vfd_conv3d_wgrad_det / vfd_conv3d_wgrad_thin_det): run-to-run identical bits, a few percent slower.
DETERMINISTIC = os.environ.get("VFD_DETERMINISTIC", "0") == "1"


def set_deterministic(on=True):
    global DETERMINISTIC
    DETERMINISTIC = bool(on)


NARROW_CONV = os.environ.get("VFD_NARROW_CONV", "1") != "0"   # conv_last forward through csrc/conv_narrow.cu
LSTM_FUSED = os.environ.get("VFD_LSTM_FUSED", "1") != "0"     # ConvLSTM step with the cell update in the gate conv's epilogue
NARROW_WGRAD = os.environ.get("VFD_NARROW_WGRAD", "1") != "0"
BN_TICKET = os.environ.get("VFD_BN_TICKET", "1") != "0"       # BatchNorm backward: last-block finalize instead of a launch
FIRST_WGRAD = os.environ.get("VFD_FIRST_WGRAD", "1") != "0"
CONV_IMPL_DIRECT = False  # tests flip this to cross-check the tcgen05 path against the CUDA-core convs of libvfd_b200_debug.so
PROFILER = None           # bench.py installs an object with .run(kind, work, thunk) to time kernels


def _timed(kind, work, thunk, nbytes=0.0):
    """Run a kernel thunk; with a profiler installed (bench.py) it is bracketed by CUDA events and recorded
    with its algorithmic work (FLOP for convs, bytes for the HBM-bound kernels) and algorithmic bytes."""
    if PROFILER is None:
        thunk()
    else:
        PROFILER.run(kind, work, thunk, nbytes)

_scratch = {}


_tickets = {}


def bn_ticket(device):
    """A zeroed device counter for vfd_bn_act_bwd's last-block finalize. The kernels leave it zero again; calls rotate
    over 64 of them so that BatchNorm backward passes on different streams never share one."""
    ent = _tickets.get(device)
    if ent is None:
        ent = _tickets[device] = [torch.zeros(64 * 32, dtype=torch.int32, device=device), 0]
    ent[1] = (ent[1] + 1) % 64
    return ent[0][ent[1] * 32:ent[1] * 32 + 1]       # 128 bytes apart


def bn_scratch(device, C):
    """Self-clearing double [2*C] accumulator shared by every BatchNorm of width C on `device`."""
    key = (device, C)
    if key not in _scratch:
        _scratch[key] = torch.zeros(2 * C, dtype=torch.float64, device=device)
    return _scratch[key]


# Bumped after every optimizer step (fused optimizers update parameters without touching
# Tensor._version, so the version counter alone cannot invalidate the packed copies).
_weight_epoch = [0]


def invalidate_packed_weights(*_args, **_kwargs):
    _weight_epoch[0] += 1


from torch.optim.optimizer import register_optimizer_step_post_hook as _reg_post_hook  # noqa: E402

_reg_post_hook(invalidate_packed_weights)


class PackedWeights:
    """bf16 GEMM operands of one conv weight (forward and tap-mirrored dgrad layouts), rebuilt when the fp32
    master changes (Adam step). The two buffers are allocated once and re-packed in place, so a
    ``WeightPacker`` can refresh every weight of a network with one kernel launch."""

    def __init__(self):
        self.key = None
        self.shape_key = None
        self.fwd = None
        self.dgrad = None

    def _allocate(self, weight):
        cout, cin = weight.shape[0], weight.shape[1]
        taps = weight[0, 0].numel()
        cin_p, cout_p = round_up(cin, 8), round_up(cout, 8)
        self.kc_f, self.kc_d = pick_kc(cin_p), pick_kc(cout_p)
        self.fwd = torch.empty(round_up(cout, 16), taps, round_up(cin_p, self.kc_f), dtype=torch.bfloat16,
                               device=weight.device)
        self.dgrad = torch.empty(round_up(cin, 16), taps, round_up(cout_p, self.kc_d), dtype=torch.bfloat16,
                                 device=weight.device)
        self.shape_key = (tuple(weight.shape), weight.device)

    def current_key(self, weight):
        return (weight.data_ptr(), weight._version, weight.device, _weight_epoch[0])

    def get(self, weight):
        key = self.current_key(weight)
        if self.key != key:
            if self.shape_key != (tuple(weight.shape), weight.device):
                self._allocate(weight)
            w = weight.detach()
            pack_weight(w, self.fwd, 0)
            pack_weight(w, self.dgrad, 1)
            self.key = key
        return self


def _packed(weight):
    pw = getattr(weight, "_vfd_packed", None)
    if pw is None:
        pw = PackedWeights()
        weight._vfd_packed = pw
    return pw.get(weight)


class WeightPacker:
    """Re-packs all conv weights of a set of modules with ONE kernel launch (vfd_pack_weights_batched)."""

    def __init__(self, modules):
        import struct
        self.weights = []
        for m in modules:
            for mod in m.modules():
                if isinstance(mod, (torch.nn.Conv3d, torch.nn.Conv2d)):
                    self.weights.append(mod.weight)
        recs, begin = [], 0
        self.entries = []
        for w in self.weights:
            pw = getattr(w, "_vfd_packed", None)
            if pw is None:
                pw = PackedWeights()
                w._vfd_packed = pw
            if pw.shape_key != (tuple(w.shape), w.device):   # keep buffers another packer / a captured graph points at
                pw._allocate(w)
            cout, cin = w.shape[0], w.shape[1]
            taps = w[0, 0].numel()
            for dst, mode in ((pw.fwd, 0), (pw.dgrad, 1)):
                rows, _, ck = dst.shape
                recs.append(struct.pack("<QQiiiiiiq", w.data_ptr(), dst.data_ptr(), cout, cin, taps, rows, ck, mode,
                                        begin))
                begin += dst.numel()
            self.entries.append((w, pw))
        self.total = begin
        self.njobs = len(recs)
        dev = self.weights[0].device
        self.jobs = torch.frombuffer(bytearray(b"".join(recs)), dtype=torch.uint8).to(dev)
        self.ptrs = [w.data_ptr() for w in self.weights]

    def pack_all(self):
        if [w.data_ptr() for w in self.weights] != self.ptrs:
            raise RuntimeError("WeightPacker: a parameter was re-allocated; rebuild the packer")
        _lib.call("vfd_pack_weights_batched", self.jobs.data_ptr(), self.njobs, self.total, _stream())
        for w, pw in self.entries:
            pw.key = pw.current_key(w)


class StepArena:
    """Zero-initialised fp32 scratch for the weight-gradient accumulators of one train step: one fill per step
    instead of one per conv. Sized by the first step that uses it (which still gets ordinary ``torch.zeros``);
    buffers are never freed while the process lives because a captured CUDA graph keeps their addresses."""

    def __init__(self):
        self.buf, self.off, self.want, self.active, self.retired = None, 0, 0, False, []

    def begin(self, device):
        if self.buf is not None and (self.buf.device != device or self.buf.numel() < self.want):
            self.retired.append(self.buf)
            self.buf = None
        if self.buf is None and self.want:
            self.buf = torch.empty(self.want + self.want // 8, dtype=torch.float32, device=device)
        if self.buf is not None:
            self.buf.zero_()
        self.off, self.want, self.active = 0, 0, True

    def end(self):
        self.active = False

    def take(self, shape, device):
        if not self.active:
            return None
        n = 1
        for d in shape:
            n *= d
        n = round_up(n, 64)
        self.want += n
        if self.buf is None or self.buf.device != device or self.off + n > self.buf.numel():
            return None
        t = self.buf[self.off:self.off + n]
        self.off += n
        numel = 1
        for d in shape:
            numel *= d
        return t[:numel].view(*shape)


ARENA = StepArena()
_const_zeros = {}


class StepContext:
    """State of one fused train step (GanTrainStep / StcnnTrainStep), active between begin() and end().

    Inside a step the autograd functions below do not hand parameter gradients back to autograd. They write them
    straight into the parameter's persistent ``.grad`` buffer (a slice of the GradAllReducer's flat bucket): the first
    write of a step overwrites, later ones (NetD runs twice per step) accumulate inside the kernel -- no
    ``AccumulateGrad`` add kernels, no per-parameter temporaries -- and tell the ``sink`` when a parameter has
    received its last contribution (uses are counted in forward), which is what triggers a bucket's all-reduce.

    Weight gradients run on a side stream: a layer's dgrad -> BatchNorm-backward chain does not depend on its
    wgrad, so the many small, latency-bound wgrad launches of the deep layers overlap that chain. Every tensor a
    side-stream kernel reads is kept referenced in ``held`` until ``join()`` has made the main stream wait for the
    side stream, so the caching allocator cannot hand its memory to a main-stream kernel early."""

    def __init__(self):
        self.active, self.sink, self.wstream = False, None, None
        self.uses, self.touched, self.held = {}, set(), []
        self.async_wgrad = os.environ.get("VFD_ASYNC_WGRAD", "1") != "0"
        self.forked = False

    def begin(self, device, sink=None):
        if self.active:
            # autograd runs backward nodes on its own worker threads, so this state cannot be thread-local; the design
            # is one process per GPU and one fused step at a time -- fail loudly instead of mixing two steps' sinks
            raise RuntimeError("a fused train step is already running in this process: GanTrainStep / StcnnTrainStep "
                               "are not re-entrant (one process per GPU)")
        self.active, self.sink = True, sink
        self.uses, self.touched, self.held, self.forked = {}, set(), [], False
        if self.async_wgrad and device.type == "cuda" and self.wstream is None:
            self.wstream = torch.cuda.Stream(device=device)

    def end(self):
        self.join()
        self.active, self.sink = False, None
        self.uses, self.touched, self.held = {}, set(), []

    # ---- gradient sinks
    def count_use(self, param):
        if self.active and param is not None and param.requires_grad:
            k = param.data_ptr()
            self.uses[k] = self.uses.get(k, 0) + 1

    def direct(self, param):
        """The persistent gradient buffer to write into, or None when gradients go through autograd."""
        if not self.active or param is None or param.grad is None or param.data_ptr() not in self.uses:
            return None
        return param.grad

    def first_touch(self, param):
        k = param.data_ptr()
        first = k not in self.touched
        self.touched.add(k)
        return first

    def done(self, param):
        """One backward use of ``param`` has deposited its gradient; tells the sink after the last one."""
        k = param.data_ptr()
        self.uses[k] -= 1
        if self.uses[k] == 0 and self.sink is not None:
            self.sink(param)

    # ---- side stream
    def side(self):
        """Context manager: run the enclosed launches on the wgrad stream, after everything enqueued so far."""
        if not (self.active and self.async_wgrad and self.wstream is not None):
            return _NullCtx()
        self.wstream.wait_stream(torch.cuda.current_stream())
        self.forked = True
        return torch.cuda.stream(self.wstream)

    def hold(self, *tensors):
        if self.forked:
            self.held.extend(t for t in tensors if t is not None)

    def join(self):
        if self.forked and self.wstream is not None:
            torch.cuda.current_stream().wait_stream(self.wstream)
            self.forked = False
        self.held = []


class _NullCtx:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


STEP = StepContext()


def acc_zeros(shape, device):
    """fp32 zeros for a wgrad accumulator (arena slice inside a GanTrainStep, a fresh tensor otherwise)."""
    t = ARENA.take(shape, device)
    return t if t is not None else torch.zeros(*shape, dtype=torch.float32, device=device)


def zero_grad(n, device):
    """Gradient of a parameter whose gradient is identically zero (a conv bias folded into a training-mode
    BatchNorm). Inside a GanTrainStep it is a view of one shared, never-written zero buffer (no fill kernel);
    elsewhere a fresh tensor, because user code may update ``.grad`` in place."""
    if not ARENA.active:
        return torch.zeros(n, dtype=torch.float32, device=device)
    z = _const_zeros.get(device)
    if z is None or z.numel() < n:
        z = torch.zeros(max(n, 4096), dtype=torch.float32, device=device)
        _const_zeros[device] = z
    return z[:n]


def cl_empty(N, D, H, W, C, device, dtype=torch.bfloat16):
    return torch.empty(N, D, H, W, C, dtype=dtype, device=device)


# ------------------------------------------------------------------------------------------------
# autograd functions
# ------------------------------------------------------------------------------------------------
class PackFn(torch.autograd.Function):
    """fp32 NCDHW -> channels-last bf16 (C padded to 8). `replicate_to` > 0 repeats a 1-channel
    source that many times (gray2rgb, lib/utils.py:91-92). Backward unpacks the gradient."""

    @staticmethod
    def forward(ctx, x, replicate_to):
        x = x.contiguous().float()
        N, Csrc = x.shape[0], x.shape[1]
        if replicate_to and Csrc != 1:
            raise RuntimeError("replicate_to needs a single-channel source")
        C = replicate_to if replicate_to else Csrc
        out = cl_empty(N, *x.shape[2:], round_up(C, 8), x.device)
        pack_ncdhw(x, out, C, bool(replicate_to))
        ctx.src_shape = x.shape
        ctx.channels = C
        ctx.replicate = bool(replicate_to)
        return out

    @staticmethod
    def backward(ctx, g):
        g = as_cl_grad(g)
        N = ctx.src_shape[0]
        gx = torch.empty(N, ctx.channels, *ctx.src_shape[2:], dtype=torch.float32, device=g.device)
        unpack_ncdhw(g, gx)
        if ctx.replicate:
            gx = gx.sum(1, keepdim=True)
        return gx, None


class UnpackFn(torch.autograd.Function):
    """channels-last (bf16 / fp32) -> fp32 NCDHW with `channels` valid channels."""

    @staticmethod
    def forward(ctx, x, channels):
        N, D, H, W, Cp = x.shape
        out = torch.empty(N, channels, D, H, W, dtype=torch.float32, device=x.device)
        unpack_ncdhw(x, out)
        ctx.cp = Cp
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().float()
        N, C, D, H, W = g.shape
        gx = cl_empty(N, D, H, W, ctx.cp, g.device)
        pack_ncdhw(g, gx, C, False)
        return gx, None


def _wshape(weight):
    """(cout, cin, kd, kh, kw) of an nn.Conv3d weight, or of an nn.Conv2d weight seen as kd = 1."""
    if weight.dim() == 4:
        return weight.shape[0], weight.shape[1], 1, weight.shape[2], weight.shape[3]
    return tuple(weight.shape)


_FOLD_X = {(1, 3, 3, 3), (1, 3, 3, 1), (1, 3, 3, 2), (3, 1, 1, 1), (3, 1, 1, 2), (3, 1, 1, 3), (3, 1, 1, 4)}
_FOLD_Y = {(3, 3, 3, 1), (1, 3, 3, 1), (1, 3, 3, 2), (1, 3, 3, 3), (3, 1, 1, 1), (3, 1, 1, 2), (3, 1, 1, 3),
           (3, 1, 1, 4)}


def wgrad_fold_mode(cin, cout, kd, kh, kw):
    """Thin convs (taps * channels <= 32 on one side): fold the taps into that side's channel dimension
    (vfd_tap_gather) and take the weight gradient as ONE 1x1x1 wgrad. "x": gather the input, "y": gather dy."""
    if kd * kh * kw == 1 or os.environ.get("VFD_WGRAD_FOLD", "1") == "0":
        return None
    if (kd, kh, kw, cin) in _FOLD_X:
        return "x"
    if (kd, kh, kw, cout) in _FOLD_Y:
        return "y"
    return None


THIN_WGRAD = os.environ.get("VFD_WGRAD_THIN", "1") != "0"


def _wgrad_1x1(dy, cout, x, cin, acc):
    """1x1x1 weight gradient into acc[0][ci][co]: the streaming mma.sync kernel for thin layers, tcgen05 else."""
    if THIN_WGRAD and cout <= 32 and cin <= 32:
        conv3d_wgrad_thin(dy, cout, x, cin, acc)
    else:
        conv3d_wgrad(dy, cout, x, cin, acc, 1, 1, 1, False)


# Gathers conv_thin.cu does in-stream. It also implements ("x", 1, 3, 3, 3) and ("y", 3, 3, 3, 1), but with 9 / 27
# taps the per-element gather is instruction bound and measured slower on B200 than vfd_tap_gather + the plain
# thin kernel (0.41 vs 0.30 ms for dconv1.spatial_conv), so only the 3-tap temporal fold is fused.
_FUSED_FOLDS = {("x", 3, 1, 1, 2)}


def _folded_wgrad(mode, g, x, cin, cout, kd, kh, kw, flops):
    """-> fp32 weight gradient as a strided [cout, cin, taps] view of the accumulator"""
    taps = kd * kh * kw
    N, D, H, W = g.shape[:4]
    cs = cin if mode == "x" else cout
    cols = round_up(taps * cs, 8)
    other = cout if mode == "x" else cin
    fused = THIN_WGRAD and other <= 32 and (mode, kd, kh, kw, cs) in _FUSED_FOLDS
    narrow = (NARROW_CONV and NARROW_WGRAD and mode == "y" and (kd, kh, kw, cout) == (3, 3, 3, 1) and cin == 32
              and x.shape[-1] == 32 and not DETERMINISTIC)
    first = (NARROW_CONV and FIRST_WGRAD and mode == "x" and (kd, kh, kw, cin) == (1, 3, 3, 3) and cout <= 32
             and not DETERMINISTIC)
    folded = None if (fused or narrow or first) else cl_empty(N, D, H, W, cols, g.device)

    def run():
        if mode == "x":     # X'[v][t*cin+ci] = x[v+off(t)][ci];  acc[0][t*cin+ci][co]
            if first:       # 1x3x3 over three channels: gather + GEMM in one kernel (csrc/conv_narrow.cu)
                conv3d_wgrad_first(g, cout, x, acc)
            elif fused:
                conv3d_wgrad_thin(g, cout, x, taps * cin, acc, 1, cs, kd, kh, kw)
            else:
                tap_gather(x, cs, folded, kd, kh, kw, 1)
                _wgrad_1x1(g, cout, folded, taps * cin, acc)
        else:               # Y'[u][t*cout+co] = dy[u-off(t)][co]; acc[0][ci][t*cout+co]
            if narrow:      # conv_last: the gather and the GEMM in one kernel (csrc/conv_narrow.cu)
                conv3d_wgrad_narrow(g, x, acc)
            elif fused:
                conv3d_wgrad_thin(g, taps * cout, x, cin, acc, 2, cs, kd, kh, kw)
            else:
                tap_gather(g, cs, folded, kd, kh, kw, -1)
                _wgrad_1x1(folded, taps * cout, x, cin, acc)

    if mode == "x":
        acc = acc_zeros((1, cols, round_up(cout, 32)), g.device)
        _timed("conv_wgrad", flops, run, 2.0 * (g.numel() + x.numel()))
        return acc[0, :taps * cin, :cout].reshape(taps, cin, cout).permute(2, 1, 0)
    acc = acc_zeros((1, round_up(cin, 8), round_up(taps * cout, 32)), g.device)
    _timed("conv_wgrad", flops, run, 2.0 * (g.numel() + x.numel()))
    return acc[0, :cin, :taps * cout].reshape(cin, taps, cout).permute(2, 0, 1)


MAX_FUSED_STATS_CHANNELS = 1024   # the conv epilogue keeps per-CTA sum / sum-of-squares for at most this many columns


def conv_fuses_stats(cout, out_fp32=False):
    """Whether ``ConvFn(..., fuse_stats=True)`` really accumulates the following BatchNorm's statistics in its
    epilogue (bf16 output, at most MAX_FUSED_STATS_CHANNELS padded channels). Callers derive ``stats_ready`` from
    this, so a wider layer takes the stand-alone ``vfd_bn_stats`` pass instead of reading an empty scratch."""
    return (not out_fp32) and round_up(cout, 8) <= MAX_FUSED_STATS_CHANNELS


def _put(dst3, view3, accumulate):
    if accumulate:
        dst3.add_(view3)
    else:
        dst3.copy_(view3)


def conv_weight_grad_into(g, x, weight, dst, accumulate):
    """Weight gradient of the stride-1 "same" conv (dy = ``g``, input ``x``, both channels-last bf16) written into
    ``dst`` (fp32, ``weight``'s shape): overwritten, or added to when ``accumulate``."""
    cout, cin, kd, kh, kw = _wshape(weight)
    N, D, H, W, _, _ = _check_cl(g, "conv grad")
    taps = kd * kh * kw
    flops = 2.0 * N * D * H * W * cin * cout * taps
    nbytes = 2.0 * (g.numel() + x.numel())
    dst3 = dst.view(cout, cin, taps)
    fold = None if CONV_IMPL_DIRECT else wgrad_fold_mode(cin, cout, kd, kh, kw)
    if fold is not None:
        _put(dst3, _folded_wgrad(fold, g, x, cin, cout, kd, kh, kw, flops), accumulate)
        return
    layout = 0 if CONV_IMPL_DIRECT else wgrad_layout(cout, cin, kd, kh, kw, H, W)
    if taps == 1 and cout <= 32 and cin <= 32 and THIN_WGRAD and not CONV_IMPL_DIRECT:
        acc = acc_zeros((1, round_up(cin, 8), round_up(cout, 32)), g.device)
        _timed("conv_wgrad", flops, lambda: conv3d_wgrad_thin(g, cout, x, cin, acc), nbytes)
        unpack_wgrad(acc, dst, accumulate)
    elif layout == 1:   # swapped GEMM roles: acc[tap][co][ci]
        acc = acc_zeros((taps, round_up(cout, 8), round_up(cin, 32)), g.device)
        _timed("conv_wgrad", flops, lambda: conv3d_wgrad(g, cout, x, cin, acc, kd, kh, kw, False, 1), nbytes)
        _put(dst3, acc[:, :cout, :cin].permute(1, 2, 0), accumulate)
    else:
        acc = acc_zeros((taps, round_up(cin, 8), round_up(cout, 32)), g.device)
        _timed("conv_wgrad", flops, lambda: conv3d_wgrad(g, cout, x, cin, acc, kd, kh, kw, CONV_IMPL_DIRECT), nbytes)
        unpack_wgrad(acc, dst, accumulate)


def conv_backward(x, weight, w_dgrad, kc_d, g, bias, bias_zero, need_x, need_w, need_b):
    """Backward of a stride-1 'same' conv given the gradient of its output: (dx, dW, db), each None when not needed or
    when it was written into the parameter's persistent gradient buffer (inside a fused train step). Shared by ConvFn
    and the fused ConvLSTM step."""
    cout, cin, kd, kh, kw = _wshape(weight)
    g = as_cl_grad(g)
    N, D, H, W, _, _ = _check_cl(g, "conv grad")
    gx = gw = gb = None
    # the weight gradient first: inside a fused step it goes to the side stream, forked here -- after dy is
    # ready, before this layer's dgrad -- so that it overlaps the dgrad -> BatchNorm-backward chain
    if need_w:
        dst = STEP.direct(weight)
        if dst is not None:
            with STEP.side():
                conv_weight_grad_into(g, x, weight, dst, not STEP.first_touch(weight))
            STEP.hold(g, x)
            STEP.done(weight)
        else:
            gw = torch.empty_like(weight, dtype=torch.float32)
            conv_weight_grad_into(g, x, weight, gw, False)
    if need_x:
        gx = cl_empty(N, D, H, W, x.shape[-1], g.device)
        flops = 2.0 * N * D * H * W * cin * cout * kd * kh * kw
        narrow = (NARROW_CONV and cout == 1 and (kd, kh, kw) == (3, 3, 3) and x.shape[-1] == 32
                  and tuple(w_dgrad.shape[:2]) == (32, 27) and not CONV_IMPL_DIRECT)
        if narrow:   # conv_last: one gradient channel in, 32 out -- taps as the GEMM's K dimension (conv_narrow.cu)
            _timed("conv_dgrad", flops, lambda: conv3d_dgrad_narrow(g, w_dgrad, gx), 2.0 * (g.numel() + gx.numel()))
        else:
            _timed("conv_dgrad", flops,
                   lambda: conv3d_fwd(g, w_dgrad, None, gx, None, kd, kh, kw, kc_d, x.shape[-1],
                                      CONV_IMPL_DIRECT), 2.0 * (g.numel() + gx.numel()))
    if bias is not None and need_b:
        dst = STEP.direct(bias)
        if bias_zero:
            if dst is None:
                gb = zero_grad(cout, g.device)
            else:                               # the buffer was zeroed at the start of the step
                STEP.first_touch(bias)
                STEP.done(bias)
        else:
            sums = torch.zeros(g.shape[-1], dtype=torch.float64, device=g.device)
            channel_sum(g, sums)
            sums = sums.float()
            if dst is None:
                gb = sums[:cout].clone()
            else:
                _put(dst, sums[:cout], not STEP.first_touch(bias))
                STEP.done(bias)
    return gx, gw, gb


class ConvFn(torch.autograd.Function):
    """Stride-1 "same" conv3d on channels-last bf16. `bias_grad_exact_zero` marks convs that feed a
    training-mode BatchNorm: there d loss / d bias is identically zero (BN removes the mean).

    Inside a fused train step (``STEP.active``) the parameter gradients are not returned to autograd: they are
    written into the parameters' persistent ``.grad`` buffers (see StepContext), the weight gradient on the side
    stream."""

    @staticmethod
    def forward(ctx, x, weight, bias, out_fp32, bias_grad_exact_zero, fuse_stats=False):
        N, D, H, W, Cin_p, _ = _check_cl(x, "conv input")
        cout, cin, kd, kh, kw = _wshape(weight)
        if round_up(cin, 8) != Cin_p:
            raise RuntimeError(f"conv input has {Cin_p} padded channels, weight expects {cin}")
        pk = _packed(weight)
        cout_p = round_up(cout, 8)
        out = cl_empty(N, D, H, W, cout_p, x.device, torch.float32 if out_fp32 else torch.bfloat16)
        b = None
        if bias is not None:
            b = torch.zeros(pk.fwd.shape[0], dtype=torch.float32, device=x.device)
            b[:cout] = bias.detach()
        flops = 2.0 * N * D * H * W * cin * cout * kd * kh * kw
        # fuse_stats: the epilogue also accumulates the following BatchNorm's sum / sum-of-squares
        # into the shared self-clearing scratch (the caller passes stats_ready=True to BnActFn)
        st = bn_scratch(x.device, cout_p) if (fuse_stats and conv_fuses_stats(cout, out_fp32)) else None
        ctx.stats_fused = st is not None
        # one output channel from 32 inputs through 27 taps (conv_last): taps as the GEMM's N dimension (conv_narrow.cu)
        narrow = (NARROW_CONV and cout == 1 and (kd, kh, kw) == (3, 3, 3) and Cin_p == 32 and out_fp32 and st is None
                  and not CONV_IMPL_DIRECT and pk.fwd.shape[2] == 32)
        if narrow:
            _timed("conv_fwd", flops, lambda: conv3d_fwd_narrow(x, pk.fwd, b, out),
                   2.0 * x.numel() + out.numel() * out.element_size())
        else:
            _timed("conv_fwd", flops,
                   lambda: conv3d_fwd(x, pk.fwd, b, out, st, kd, kh, kw, pk.kc_f, cout_p, CONV_IMPL_DIRECT),
                   2.0 * x.numel() + out.numel() * out.element_size())
        ctx.save_for_backward(x, weight)
        ctx.w_dgrad, ctx.kc_d = pk.dgrad, pk.kc_d     # the master cannot change between forward and backward
        ctx.bias_param = bias
        ctx.bias_zero = bias_grad_exact_zero
        STEP.count_use(weight)
        STEP.count_use(bias)
        return out

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        gx, gw, gb = conv_backward(x, weight, ctx.w_dgrad, ctx.kc_d, g, ctx.bias_param, ctx.bias_zero,
                                   ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        return gx, gw, gb, None, None, None


def lstm_gate_operands(weight, bias):
    """Packed forward weights / bias of a ConvLSTM gate conv with the rows re-ordered for the fused step kernel:
    [n-tile][gate i, f, o, g][64 hidden channels] instead of [gate][hidden channel] (models/convlstm.py:49 splits the
    4*hid output channels gate-major), so that one 256-column accumulator tile holds all four gates of 64 channels."""
    pk = _packed(weight)
    hid = weight.shape[0] // 4
    rows, taps, ck = pk.fwd.shape
    w_perm = pk.fwd[:4 * hid].view(4, hid // 64, 64, taps, ck).permute(1, 0, 2, 3, 4).contiguous().view(4 * hid, taps, ck)
    b_perm = None
    if bias is not None:
        b_perm = bias.detach().float().view(4, hid // 64, 64).permute(1, 0, 2).contiguous().view(-1)
    return w_perm, b_perm, pk.kc_f


def lstm_step_fusable(weight, cin_p):
    """The fused ConvLSTM step serves hidden sizes that are multiples of 64 (one accumulator tile = 4 gates x 64
    channels) with square 1x1 / 3x3 gate kernels."""
    if weight.dim() != 4:
        return False
    rows, cin, kh, kw = weight.shape
    return (LSTM_FUSED and rows % 256 == 0 and kh == kw and kh in (1, 3) and round_up(cin, 8) == cin_p
            and not CONV_IMPL_DIRECT)


class LstmStepFn(torch.autograd.Function):
    """One ConvLSTM step (models/convlstm.py:46-58) as ONE forward kernel: the gate conv on tcgen05 with sigmoid / tanh
    and the cell update in its epilogue -- the gates never go to HBM. (comb bf16 [N,1,H,W,in+hid], c_cur fp32
    [N,H,W,hid]) -> (h_next bf16 [N,1,H,W,hid], c_next fp32). Backward: the cell kernel turns (dh, dc) into the gate
    gradients, then the ordinary conv dgrad / wgrad on the un-permuted weights."""

    @staticmethod
    def forward(ctx, comb, c_cur, weight, bias, w_perm, b_perm, kc):
        N, _, H, W, _, _ = _check_cl(comb, "convlstm step input")
        hid = weight.shape[0] // 4
        c_cur = c_cur.contiguous()
        h = cl_empty(N, 1, H, W, hid, comb.device)
        c_next = torch.empty_like(c_cur)
        act = torch.empty(N, H, W, 4 * hid, dtype=torch.float32, device=comb.device)
        flops = 2.0 * N * H * W * weight.shape[0] * weight.shape[1] * weight.shape[2] * weight.shape[3]
        _timed("conv_fwd", flops,
               lambda: convlstm_step_fwd(comb, w_perm, b_perm, c_cur, c_next, h, act, weight.shape[2], weight.shape[3], kc),
               2.0 * comb.numel() + 2.0 * h.numel() + 4.0 * (2 * c_cur.numel() + act.numel()))
        pk = _packed(weight)
        ctx.save_for_backward(comb, weight, act, c_cur, c_next)
        ctx.w_dgrad, ctx.kc_d = pk.dgrad, pk.kc_d
        ctx.bias_param = bias
        STEP.count_use(weight)
        STEP.count_use(bias)
        return h, c_next

    @staticmethod
    def backward(ctx, dh, dc):
        comb, weight, act, c_cur, c_next = ctx.saved_tensors
        N, _, H, W, _, _ = _check_cl(comb, "convlstm step saved input")
        hid = weight.shape[0] // 4
        dgates = torch.zeros(N, 1, H, W, 4 * hid, dtype=torch.bfloat16, device=act.device)
        dc_cur = torch.empty_like(c_cur)
        dh32 = None if dh is None else dh.reshape(N, H, W, hid).float().contiguous()
        convlstm_cell_bwd(act, c_cur, c_next, dh32, None if dc is None else dc.contiguous(), dgates, dc_cur)
        gx, gw, gb = conv_backward(comb, weight, ctx.w_dgrad, ctx.kc_d, dgates, ctx.bias_param, False,
                                   ctx.needs_input_grad[0], ctx.needs_input_grad[2], ctx.needs_input_grad[3])
        return gx, dc_cur, gw, gb, None, None, None


class BnActFn(torch.autograd.Function):
    """BatchNorm3d (batch statistics in training) + (Leaky)ReLU [+ Dropout] [+ AvgPool3d].

    Returns (full, pooled); either is None when not requested. `full_out` may be a channel slice
    of a concat buffer, in which case the activation is written straight into it."""

    @staticmethod
    def forward(ctx, y, gamma, beta, pre_bias, running_mean, running_var, train, momentum, eps, slope, pool, drop_p,
                seed, want_full, want_pool, full_out_holder, stats_ready=False, seed_dev=None):
        N, D, H, W, C, _ = _check_cl(y, "bn input")
        cvalid = gamma.numel()
        dev = y.device
        stats = torch.empty(4, C, dtype=torch.float32, device=dev)
        mean, invstd, scale, shift = stats[0], stats[1], stats[2], stats[3]
        pb = None if pre_bias is None else pre_bias.detach()
        _timed("bn_stats", 2.0 * y.numel(), lambda: bn_prepare(y, bn_scratch(dev, C), cvalid, pb, gamma.detach(), beta.detach(),
                                                   running_mean, running_var, momentum, eps, train, mean, invstd,
                                                   scale, shift, bool(stats_ready and train)))
        pd, ph, pw = pool
        full = pooled = None
        if want_full:
            full = full_out_holder[0].detach() if full_out_holder is not None else cl_empty(N, D, H, W, C, dev)
        if want_pool:
            pooled = cl_empty(N, D // pd, H // ph, W // pw, C, dev)
        nbytes = 2.0 * y.numel() * (1 + (1 if want_full else 0)) + (2.0 * pooled.numel() if want_pool else 0)
        _timed("bn_act_fwd", nbytes,
               lambda: bn_act_fwd(y, scale, shift, slope, full, pooled, pd, ph, pw, drop_p, seed, seed_dev))
        ctx.save_for_backward(y, stats)
        ctx.cfg = (cvalid, slope, pool, drop_p, seed, train)
        ctx.seed_dev = seed_dev
        ctx.params = (gamma, beta, pre_bias)
        for prm in ctx.params:
            STEP.count_use(prm)
        return full, pooled

    @staticmethod
    def backward(ctx, g_full, g_pool):
        y, stats = ctx.saved_tensors
        cvalid, slope, (pd, ph, pw), drop_p, seed, train = ctx.cfg
        N, D, H, W, C, _ = _check_cl(y, "bn saved input")
        dev = y.device
        # the gradient of a global spatial mean (TDisc head) arrives as a stride-0 expansion over H and W: hand the
        # kernel the [N, D/pd, C] base instead of materialising the broadcast
        bcast = False
        if (g_pool is not None and ph == 1 and pw == 1 and g_pool.dim() == 5 and g_pool.dtype == torch.bfloat16
                and g_pool.shape[2] * g_pool.shape[3] > 1 and g_pool.stride(2) == 0 and g_pool.stride(3) == 0):
            base = g_pool[:, :, 0, 0, :]
            if base.is_contiguous() and base.data_ptr() % 16 == 0:
                g_pool, bcast = base.unsqueeze(2).unsqueeze(3), True
        g_full, g_pool = as_cl_grad(g_full), as_cl_grad(g_pool)
        dy = cl_empty(N, D, H, W, C, dev)
        tmp = torch.empty(2, C, dtype=torch.float32, device=dev)
        gamma, beta, pre_bias = ctx.params
        # inside a fused train step dgamma / dbeta land directly in the parameters' .grad buffers
        dgamma, dbeta = STEP.direct(gamma), STEP.direct(beta)
        direct = dgamma is not None and dbeta is not None and ctx.needs_input_grad[1] and ctx.needs_input_grad[2]
        accumulate = False
        if direct:
            accumulate = not STEP.first_touch(gamma)
            STEP.first_touch(beta)
        else:
            dgamma = torch.empty(cvalid, dtype=torch.float32, device=dev)
            dbeta = torch.empty(cvalid, dtype=torch.float32, device=dev)
        nbytes = 2.0 * y.numel() * 3 + 2.0 * 2 * sum(g.numel() for g in (g_full, g_pool) if g is not None)
        _timed("bn_act_bwd", nbytes,
               lambda: bn_act_bwd(y, cvalid, stats[0], stats[1], stats[2], stats[3], slope, g_full, g_pool, pd, ph, pw,
                                  drop_p, seed, train, bn_scratch(dev, C), tmp[0], tmp[1], dgamma, dbeta, dy,
                                  ctx.seed_dev, bcast, accumulate))
        if direct:
            STEP.done(gamma)
            STEP.done(beta)
            dgamma = dbeta = None
        # a conv bias folded into training-mode BN has an identically zero gradient
        gpb = None
        if pre_bias is not None and ctx.needs_input_grad[3]:
            if STEP.direct(pre_bias) is not None:
                STEP.first_touch(pre_bias)
                STEP.done(pre_bias)
            else:
                gpb = zero_grad(cvalid, dev)
        return (dy, dgamma, dbeta, gpb) + (None,) * 14


class UpCatFn(torch.autograd.Function):
    """x2 trilinear upsample of `low` written into channels [0, C_low) of `buf`; `skip` already
    lives in channels [C_low, C_low + C_skip) of the same buffer (zero-copy torch.cat)."""

    @staticmethod
    def forward(ctx, low, skip, buf_holder):
        buf = buf_holder[0]
        c_low = low.shape[-1]
        upsample2x_fwd(low, buf[..., :c_low])
        ctx.c_low = c_low
        ctx.low_shape = low.shape
        return buf.detach().view(buf.shape)

    @staticmethod
    def backward(ctx, g):
        g = as_cl_grad(g)
        c_low = ctx.c_low
        g_low = torch.empty(ctx.low_shape, dtype=torch.bfloat16, device=g.device)
        upsample2x_bwd(g[..., :c_low], g_low)
        return g_low, g[..., c_low:], None


class SigmoidHeadFn(torch.autograd.Function):
    """predict = sigmoid(conv_last logits) as fp32 [N,1,D,H,W] (NCDHW == channels-last for C = 1)."""

    @staticmethod
    def forward(ctx, logits):
        N, D, H, W, _ = logits.shape
        predict = torch.empty(N, 1, D, H, W, dtype=torch.float32, device=logits.device)
        sigmoid_head_fwd(logits, predict)
        ctx.save_for_backward(predict)
        return predict

    @staticmethod
    def backward(ctx, g):
        (predict,) = ctx.saved_tensors
        N, _, D, H, W = predict.shape
        dlogit = cl_empty(N, D, H, W, 8, predict.device)
        sigmoid_head_bwd(g.contiguous().float(), predict, dlogit)
        return dlogit


class WeightedBceFn(torch.autograd.Function):
    """weighted_bce (lib/utils.py:65-71) with the gradient produced in the same pass."""

    @staticmethod
    def forward(ctx, predict, target, pos_weight):
        p = predict.contiguous().float()
        t = target.contiguous().float()
        V = p.numel()
        loss_sum = torch.zeros((), dtype=torch.float64, device=p.device)
        gpred = torch.empty_like(p) if predict.requires_grad else None
        weighted_bce_op(p, t, float(pos_weight), 1.0 / V, loss_sum, gpred)
        ctx.save_for_backward(gpred)
        return (-(loss_sum / V)).float()

    @staticmethod
    def backward(ctx, g):
        (gpred,) = ctx.saved_tensors
        return gpred * g, None, None


class MeanDimsFn(torch.autograd.Function):
    """fp32 mean of the first `c` channels of a channels-last bf16 tensor over voxel dims `dims` (the global
    AvgPool3d heads of SDisc / TDisc, models/mygannet.py:133,175). Backward writes the (constant per pooled
    element) gradient as one bf16 expand instead of autograd's fp32 expand + slice + cast chain."""

    @staticmethod
    def forward(ctx, feat, dims, c):
        ctx.shape, ctx.dims, ctx.c = tuple(feat.shape), tuple(dims), c
        return feat[..., :c].mean(dim=dims, dtype=torch.float32)

    @staticmethod
    def backward(ctx, g):
        shape, dims, c = ctx.shape, ctx.dims, ctx.c
        count = 1
        for d in dims:
            count *= shape[d]
        kept = [s for i, s in enumerate(shape[:-1]) if i not in dims]
        gb = torch.zeros(*kept, shape[-1], dtype=torch.bfloat16, device=g.device)
        gb[..., :c] = g / count
        for d in sorted(dims):
            gb = gb.unsqueeze(d)
        if dims == (2, 3):      # BnActFn.backward reads the stride-0 expansion directly (no broadcast copy)
            return gb.expand(shape), None, None
        return gb.expand(shape).contiguous(), None, None


class IdentityPoolFn(torch.autograd.Function):
    """AvgPool3d (and / or Dropout) with no BatchNorm in front (the shortcut branch of C2plus1d_Block,
    models/mystcnn.py:37-44): the fused BN+act kernel run with scale 1, shift 0 and slope 1. ``seed_dev``: optional
    device counter added to the dropout seed inside the kernels (fresh masks under CUDA-graph replay)."""

    @staticmethod
    def forward(ctx, x, pool, drop_p, seed, seed_dev=None):
        N, D, H, W, C, _ = _check_cl(x, "pool input")
        pd, ph, pw = pool
        dev = x.device
        k = torch.zeros(4, C, dtype=torch.float32, device=dev)   # mean 0, invstd 1, scale 1, shift 0
        k[1:3] = 1.0
        pooled = pd * ph * pw > 1
        out = cl_empty(N, D // pd, H // ph, W // pw, C, dev)
        bn_act_fwd(x, k[2], k[3], 1.0, None if pooled else out, out if pooled else None, pd, ph, pw, drop_p, seed,
                   seed_dev)
        ctx.save_for_backward(x, k)
        ctx.cfg = (pool, drop_p, seed, pooled)
        ctx.seed_dev = seed_dev
        return out

    @staticmethod
    def backward(ctx, g):
        x, k = ctx.saved_tensors
        (pd, ph, pw), drop_p, seed, pooled = ctx.cfg
        N, D, H, W, C, _ = _check_cl(x, "pool saved input")
        g = as_cl_grad(g)
        dx = cl_empty(N, D, H, W, C, x.device)
        tmp = torch.empty(4, C, dtype=torch.float32, device=x.device)
        bn_act_bwd(x, C, k[0], k[1], k[2], k[3], 1.0, None if pooled else g, g if pooled else None, pd, ph, pw,
                   drop_p, seed, False, bn_scratch(x.device, C), tmp[0], tmp[1], tmp[2], tmp[3], dx, ctx.seed_dev)
        return dx, None, None, None, None


class UpsampleFn(torch.autograd.Function):
    """nn.Upsample(scale_factor=2, 'trilinear', align_corners=True) on channels-last bf16."""

    @staticmethod
    def forward(ctx, x):
        N, D, H, W, C, _ = _check_cl(x, "upsample input")
        out = cl_empty(N, 2 * D, 2 * H, 2 * W, C, x.device)
        upsample2x_fwd(x, out)
        ctx.low_shape = x.shape
        return out

    @staticmethod
    def backward(ctx, g):
        g = as_cl_grad(g)
        gx = torch.empty(ctx.low_shape, dtype=torch.bfloat16, device=g.device)
        upsample2x_bwd(g, gx)
        return gx


class LatentL2Fn(torch.autograd.Function):
    """l2_loss of two channels-last bf16 latents (l_enc, models/ganomaly.py:439,477) with gradients to both,
    returned together with the per-clip means the anomaly score uses (models/ganomaly.py:372)."""

    @staticmethod
    def forward(ctx, a, b, valid_channels):
        N, D, H, W, C, _ = _check_cl(a, "latent a")
        per_clip = torch.zeros(N, dtype=torch.float64, device=a.device)
        latent_score(a, b, per_clip)
        count = D * H * W * valid_channels
        scores = torch.empty(N, dtype=torch.float32, device=a.device)
        score_finalize(per_clip, 1.0 / max(count, 1), scores, None)
        ctx.save_for_backward(a, b)
        ctx.inv = 1.0 / max(N * count, 1)
        ctx.mark_non_differentiable(scores)
        return (per_clip.sum() * ctx.inv).float(), scores

    @staticmethod
    def backward(ctx, g, _gs):
        a, b = ctx.saved_tensors
        ga = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        gb = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        sqdiff_bwd(a, b, g.contiguous().float().reshape(1), ctx.inv, ga, gb)
        return ga, gb, None


class L1LossFn(torch.autograd.Function):
    """nn.L1Loss() (mean) on fp32 tensors (l_con of the enc-dec-enc composition, models/ganomaly.py:438,476)."""

    @staticmethod
    def forward(ctx, a, b):
        a, b = a.contiguous().float(), b.contiguous().float()
        if not a.is_cuda:
            raise RuntimeError("l1_loss: vfd_gan_b200 has no CPU path")
        out = torch.zeros((), dtype=torch.float64, device=a.device)
        ga = torch.empty_like(a) if ctx.needs_input_grad[0] or ctx.needs_input_grad[1] else None
        l1_loss_op(a, b, 1.0 / max(a.numel(), 1), out, ga)
        ctx.save_for_backward(ga)
        return (out / max(a.numel(), 1)).float()

    @staticmethod
    def backward(ctx, g):
        (ga,) = ctx.saved_tensors
        return (ga * g if ctx.needs_input_grad[0] else None), (-ga * g if ctx.needs_input_grad[1] else None)


class BceLossFn(torch.autograd.Function):
    """nn.BCELoss() (mean, log clamped at -100) on fp32 tensors, gradient to the prediction only."""

    @staticmethod
    def forward(ctx, p, t):
        p, t = p.contiguous().float(), t.contiguous().float()
        if not p.is_cuda:
            raise RuntimeError("bce_loss: vfd_gan_b200 has no CPU path")
        out = torch.zeros((), dtype=torch.float64, device=p.device)
        gp = torch.empty_like(p) if ctx.needs_input_grad[0] else None
        bce_loss_op(p, t, 1.0 / max(p.numel(), 1), out, gp)
        ctx.save_for_backward(gp)
        return (out / max(p.numel(), 1)).float()

    @staticmethod
    def backward(ctx, g):
        (gp,) = ctx.saved_tensors
        return gp * g, None


def mse_cl(a, b, valid_channels):
    """l2_loss (lib/utils.py:59-63) of two channels-last bf16 feature maps (no gradient)."""
    out = torch.zeros((), dtype=torch.float64, device=a.device)
    sqdiff(a, b, out)
    N, D, H, W, _ = a.shape
    return (out / (N * D * H * W * valid_channels)).float()
