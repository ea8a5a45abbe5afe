"""Generator / discriminator of the vfd_gan ``mygan`` model on the B200 kernels.

Same module tree, constructor signatures and ``state_dict`` as the reference
(``models/mygannet.py:13-213``): ``NetgConv``, ``NetG``, ``NetdConv``, ``SDisc``, ``TDisc``, ``NetD``.
``forward`` keeps the reference contract (fp32 NCDHW tensors); ``forward_cl`` is the internal
channels-last bf16 path the fused train step uses to avoid layout round trips.

Fusions relative to the reference graph (all inside libvfd_b200):
  * BatchNorm3d + LeakyReLU + AvgPool3d (+ Dropout) are one kernel;
  * ``torch.cat([upsampled, skip])`` is zero-copy: the encoder writes its skip activation straight
    into the channel slice of the decoder's input buffer and the x2 trilinear upsample writes the
    other slice.
"""
import torch
import torch.nn as nn

from . import ops
from .spatiotempconv import SpatioTemporalConv, bn_apply


def _draw_seed():
    # one 62-bit draw from torch's CPU generator: dropout masks follow torch.manual_seed
    return int(torch.randint(0, 2 ** 62, (1,)).item())


class NetgConv(nn.Module):
    """SpatioTemporalConv(k, same) -> BatchNorm3d -> LeakyReLU(0.2) (models/mygannet.py:13-28)."""

    def __init__(self, in_fi, out_fi, kernel_size=3):
        super().__init__()
        self.conv = SpatioTemporalConv(in_fi, out_fi, kernel_size, padding=kernel_size // 2)
        self.bn = nn.BatchNorm3d(out_fi)
        self.lrelu = nn.LeakyReLU(0.2, inplace=True)
        self.out_fi = out_fi

    def forward_cl(self, xc, **kw):
        if self.bn.training or self.bn.running_mean is None:
            y, tb = self.conv.forward_cl(xc, fold_bias=True)
            return bn_apply(self.bn, y, self.lrelu.negative_slope, pre_bias=tb,
                            stats_ready=ops.conv_fuses_stats(self.out_fi), **kw)
        return bn_apply(self.bn, self.conv.forward_cl(xc), self.lrelu.negative_slope, **kw)

    def forward(self, x):
        full, _ = self.forward_cl(ops.PackFn.apply(x, 0))
        return ops.UnpackFn.apply(full, self.out_fi)


class NetG(nn.Module):
    """3-D U-Net mask generator (models/mygannet.py:31-101); returns ``predict`` in (0, 1)."""

    def __init__(self, nc=3, ngf=32):
        super().__init__()
        if ngf % 8:
            raise ValueError("NetG on B200 needs ngf % 8 == 0 (channel slices of the concat buffers are 16-byte aligned)")
        self.dconv1 = NetgConv(nc, ngf)
        self.dconv2 = NetgConv(ngf, ngf * 2)
        self.dconv3 = NetgConv(ngf * 2, ngf * 4)
        self.dconv4 = NetgConv(ngf * 4, ngf * 8)
        self.dconv5 = NetgConv(ngf * 8, ngf * 16)

        self.avgpool = nn.AvgPool3d(2)

        self.uconv5 = NetgConv(ngf * 16, ngf * 8)
        self.uconv4 = NetgConv(ngf * 8 + ngf * 8, ngf * 8)
        self.uconv3 = NetgConv(ngf * 8 + ngf * 4, ngf * 4)
        self.uconv2 = NetgConv(ngf * 4 + ngf * 2, ngf * 2)
        self.uconv1 = NetgConv(ngf * 2 + ngf, ngf)

        self.dropout = nn.Dropout(p=0.25)
        self.upsamp = nn.Upsample(scale_factor=2, mode='trilinear', align_corners=True)

        self.conv_last = nn.Conv3d(ngf, 1, 3, stride=1, padding=1, bias=False)
        self.sigmoid = nn.Sigmoid()
        self.ngf = ngf
        self.last_dropout_seeds = None

    def forward_cl(self, xc, dropout_seeds=None, seed_dev=None):
        """channels-last bf16 clip -> (fp32 conv_last logits channels-last, latent_i channels-last).
        ``seed_dev``: optional device int64 scalar added to every dropout seed inside the kernels (the
        per-step counter of a CUDA-graph-captured train step)."""
        N, D, H, W, _ = xc.shape
        if D % 16 or H % 16 or W % 16:
            raise RuntimeError(f"NetG needs nfr and isize divisible by 16, got D={D} H={H} W={W}")
        g = self.ngf
        dev = xc.device
        enc = [self.dconv1, self.dconv2, self.dconv3, self.dconv4]
        up_c = [2 * g, 4 * g, 8 * g, 8 * g]       # channels arriving from the decoder at each level
        skip_c = [g, 2 * g, 4 * g, 8 * g]
        bufs, skips = [], []
        x = xc
        for lvl, blk in enumerate(enc):
            d, h, w = D >> lvl, H >> lvl, W >> lvl
            buf = ops.cl_empty(N, d, h, w, up_c[lvl] + skip_c[lvl], dev)
            full, x = blk.forward_cl(x, pool=(2, 2, 2), want_full=True, want_pool=True,
                                     full_out=buf[..., up_c[lvl]:])
            bufs.append(buf)
            skips.append(full)
        latent, _ = self.dconv5.forward_cl(x)
        latent = self.bottleneck_cl(latent)

        p = self.dropout.p if self.dropout.training else 0.0
        if p > 0.0:
            seeds = list(dropout_seeds) if dropout_seeds is not None else [_draw_seed() for _ in range(4)]
        else:
            seeds = [0, 0, 0, 0]
        self.last_dropout_seeds = seeds
        x, _ = self.uconv5.forward_cl(latent, drop_p=p, seed=seeds[0], seed_dev=seed_dev)
        dec = [self.uconv4, self.uconv3, self.uconv2, self.uconv1]
        for i, blk in enumerate(dec):
            lvl = 3 - i
            cat = ops.UpCatFn.apply(x, skips[lvl], [bufs[lvl]])
            if i < 3:
                x, _ = blk.forward_cl(cat, drop_p=p, seed=seeds[i + 1], seed_dev=seed_dev)
            else:
                x, _ = blk.forward_cl(cat)
        logits = ops.ConvFn.apply(x, self.conv_last.weight, None, True, False)
        return logits, latent

    def bottleneck_cl(self, latent):
        """Hook between dconv5 and uconv5 (identity in the reference's NetG; composed.NetGLstm overrides it)."""
        return latent

    def encode_cl(self, xc):
        """dconv1..dconv5 with the 2x2x2 average pools only (the encoder half, models/mygannet.py:57-71)."""
        x = xc
        for blk in (self.dconv1, self.dconv2, self.dconv3, self.dconv4):
            _, x = blk.forward_cl(x, pool=(2, 2, 2), want_full=False, want_pool=True)
        latent, _ = self.dconv5.forward_cl(x)
        return latent

    def forward(self, x):
        logits, _ = self.forward_cl(ops.PackFn.apply(x, 0))
        return ops.SigmoidHeadFn.apply(logits)


class NetdConv(nn.Module):
    """SpatioTemporalConv -> BatchNorm3d -> LeakyReLU() (slope 0.01) (models/mygannet.py:104-116)."""

    def __init__(self, in_fi, out_fi, kernel_size=None, padding=None):
        super().__init__()
        self.conv = SpatioTemporalConv(in_fi, out_fi, kernel_size, padding=padding)
        self.bn = nn.BatchNorm3d(out_fi)
        self.lrelu = nn.LeakyReLU()
        self.out_fi = out_fi

    def forward_cl(self, xc, **kw):
        if self.bn.training or self.bn.running_mean is None:
            y, tb = self.conv.forward_cl(xc, fold_bias=True)
            return bn_apply(self.bn, y, self.lrelu.negative_slope, pre_bias=tb,
                            stats_ready=ops.conv_fuses_stats(self.out_fi), **kw)
        return bn_apply(self.bn, self.conv.forward_cl(xc), self.lrelu.negative_slope, **kw)

    def forward(self, x):
        full, _ = self.forward_cl(ops.PackFn.apply(x, 0))
        return ops.UnpackFn.apply(full, self.out_fi)


class SDisc(nn.Module):
    """Spatial discriminator: six (1,3,3) blocks with (1,2,2) pools (models/mygannet.py:119-162).

    ``isize`` generalises the reference's hard-coded 2x2 final map (``Linear(ndf*32*2*2, 1)`` only
    fits isize=128); at the default 128 the layer is identical to the reference's."""

    def __init__(self, nc, nfr, ndf=32, kernel=None, padding=None, isize=128):
        super().__init__()
        netdconv = lambda in_fi, out_fi: NetdConv(in_fi, out_fi, kernel_size=kernel, padding=padding)
        self.dconv1 = netdconv(nc, ndf)
        self.dconv2 = netdconv(ndf, ndf * 2)
        self.dconv3 = netdconv(ndf * 2, ndf * 4)
        self.dconv4 = netdconv(ndf * 4, ndf * 8)
        self.dconv5 = netdconv(ndf * 8, ndf * 16)
        self.dconv6 = netdconv(ndf * 16, ndf * 32)

        self.avgpool = nn.AvgPool3d((1, 2, 2))
        self.gpool = nn.AvgPool3d((nfr, 1, 1), stride=1)
        self.linear = nn.Linear(ndf * 32 * (isize // 64) ** 2, 1)
        self.sigmoid = nn.Sigmoid()
        self.feat_channels = ndf * 32

    def forward_cl(self, xc):
        """-> (classifier [N], features channels-last bf16 [N, nfr, S/64, S/64, ndf*32])."""
        x = xc
        for blk in (self.dconv1, self.dconv2, self.dconv3, self.dconv4, self.dconv5, self.dconv6):
            _, x = blk.forward_cl(x, pool=(1, 2, 2), want_full=False, want_pool=True)
        feat = x
        c = self.feat_channels
        kd = self.gpool.kernel_size[0]
        if feat.shape[1] != kd:
            raise RuntimeError(f"SDisc built for nfr={kd}, got {feat.shape[1]} frames")
        pooled = ops.MeanDimsFn.apply(feat, (1,), c)                         # [N, h, w, C]
        flat = pooled.permute(0, 3, 1, 2).reshape(pooled.shape[0], -1)      # (C, h, w) order like .view()
        cls = self.sigmoid(self.linear(flat))
        return cls.squeeze(1), feat

    def forward(self, x):
        cls, feat = self.forward_cl(ops.PackFn.apply(x, 0))
        return cls, ops.UnpackFn.apply(feat, self.feat_channels)


class TDisc(nn.Module):
    """Temporal (optical-flow) discriminator: three (3,1,1) blocks with (2,1,1) pools
    (models/mygannet.py:164-196). ``nfr`` generalises ``Linear(ndf*4*2, 1)`` (nfr=16 only)."""

    def __init__(self, nc, isize, ndf=32, kernel=None, padding=None, nfr=16):
        super().__init__()
        netdconv = lambda in_fi, out_fi: NetdConv(in_fi, out_fi, kernel_size=kernel, padding=padding)
        self.dconv1 = netdconv(nc, ndf)
        self.dconv2 = netdconv(ndf, ndf * 2)
        self.dconv3 = netdconv(ndf * 2, ndf * 4)

        self.avgpool = nn.AvgPool3d((2, 1, 1))
        self.gpool = nn.AvgPool3d((1, isize, isize), stride=1)
        self.linear = nn.Linear(ndf * 4 * (nfr // 8), 1)
        self.sigmoid = nn.Sigmoid()
        self.feat_channels = ndf * 4

    def forward_cl(self, xc):
        """-> (classifier [N], features channels-last bf16 [N, nfr/8, S, S, ndf*4])."""
        x = xc
        for blk in (self.dconv1, self.dconv2, self.dconv3):
            _, x = blk.forward_cl(x, pool=(2, 1, 1), want_full=False, want_pool=True)
        feat = x
        c = self.feat_channels
        ks = self.gpool.kernel_size
        if feat.shape[2] != ks[1] or feat.shape[3] != ks[2]:
            raise RuntimeError(f"TDisc built for isize={ks[1]}, got {feat.shape[2]}x{feat.shape[3]}")
        pooled = ops.MeanDimsFn.apply(feat, (2, 3), c)                       # [N, d, C]
        flat = pooled.permute(0, 2, 1).reshape(pooled.shape[0], -1)         # (C, d) order like .view()
        cls = self.sigmoid(self.linear(flat))
        return cls.squeeze(1), feat

    def forward(self, x):
        cls, feat = self.forward_cl(ops.PackFn.apply(x, 0))
        return cls, ops.UnpackFn.apply(feat, self.feat_channels)


class NetD(nn.Module):
    """Two-branch discriminator (models/mygannet.py:200-213): ``NetD(args)`` reads ``args.nfr`` and
    ``args.isize``; ``forward(x, y) -> (s_cls, s_feat, t_cls, t_feat)``."""

    def __init__(self, args):
        super().__init__()
        self.spatdisc = SDisc(3, args.nfr, kernel=(1, 3, 3), padding=(0, 1, 1), isize=args.isize)
        self.tempdisc = TDisc(3, args.isize, kernel=(3, 1, 1), padding=(1, 0, 0), nfr=args.nfr)

    def forward_cl(self, xc, yc):
        s_cls, s_feat = self.spatdisc.forward_cl(xc)
        t_cls, t_feat = self.tempdisc.forward_cl(yc)
        return s_cls, s_feat, t_cls, t_feat

    def forward(self, x, y):
        s_cls, s_feat = self.spatdisc(x)
        t_cls, t_feat = self.tempdisc(y)
        return s_cls, s_feat, t_cls, t_feat
