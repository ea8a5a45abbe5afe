"""R(2+1)D factored convolution block on the B200 kernels.

Drop-in for the reference's ``models/spatiotempconv.py:7-65``: same constructor arguments, same
children (``spatial_conv``, ``bn``, ``relu``, ``temporal_conv`` -- real ``nn.Conv3d`` /
``nn.BatchNorm3d`` parameter holders so ``weights_init`` and ``state_dict`` behave identically),
same forward contract (fp32 NCDHW in, fp32 NCDHW out). The arithmetic runs in the tcgen05
implicit-GEMM conv and the fused BatchNorm+ReLU kernels (``ops.ConvFn`` / ``ops.BnActFn``).
"""
import math

import torch
import torch.nn as nn
from torch.nn.modules.utils import _triple

from . import ops


def intermed_channels(in_channels, out_channels, kernel_size):
    """Parameter-matching bottleneck width M of R(2+1)D (models/spatiotempconv.py:44-45)."""
    kt, kh, kw = kernel_size
    return int(math.floor((kt * kh * kw * in_channels * out_channels) /
                          (kh * kw * in_channels + kt * out_channels)))


DEFER_BN_COUNTERS = None   # list collecting num_batches_tracked tensors while GanTrainStep runs


def bn_apply(bn, y, slope, pool=(1, 1, 1), drop_p=0.0, seed=0, want_full=True, want_pool=False, full_out=None,
             pre_bias=None, stats_ready=False, seed_dev=None):
    """nn.BatchNorm3d ``bn`` + (Leaky)ReLU(slope) [+ dropout] [+ average pool] on channels-last bf16.
    ``pre_bias``: bias of the conv that produced ``y`` when it was left out of ``y`` (see
    ``SpatioTemporalConv.forward_cl``)."""
    train = bn.training or bn.running_mean is None
    if bn.momentum is None:
        raise NotImplementedError("BatchNorm3d(momentum=None) (cumulative average) is not supported")
    if not bn.affine:
        raise NotImplementedError("BatchNorm3d(affine=False) is not supported")
    full, pooled = ops.BnActFn.apply(y, bn.weight, bn.bias, pre_bias, bn.running_mean, bn.running_var, train,
                                     float(bn.momentum), float(bn.eps), float(slope), tuple(pool), float(drop_p),
                                     int(seed), want_full, want_pool, None if full_out is None else [full_out],
                                     stats_ready, seed_dev)
    if train and bn.track_running_stats and bn.num_batches_tracked is not None:
        if DEFER_BN_COUNTERS is not None:
            DEFER_BN_COUNTERS.append(bn.num_batches_tracked)   # the fused train step adds them in one launch
        else:
            bn.num_batches_tracked += 1
    return full, pooled


class SpatioTemporalConv(nn.Module):
    """``Conv3d(1 x kh x kw)`` -> ``BatchNorm3d`` -> ``ReLU`` -> ``Conv3d(kt x 1 x 1)``.

    Only what the vfd_gan hot path uses is implemented in CUDA: stride 1, kernel extents 1 or 3 and
    "same" padding (``padding == kernel // 2`` per axis); anything else raises NotImplementedError.
    """

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True):
        super().__init__()
        k, s, p = _triple(kernel_size), _triple(stride), _triple(padding)
        mid = intermed_channels(in_channels, out_channels, k)
        self.spatial_conv = nn.Conv3d(in_channels, mid, (1, k[1], k[2]), stride=(1, s[1], s[2]),
                                      padding=(0, p[1], p[2]), bias=bias)
        self.bn = nn.BatchNorm3d(mid)
        self.relu = nn.ReLU()
        self.temporal_conv = nn.Conv3d(mid, out_channels, (k[0], 1, 1), stride=(s[0], 1, 1),
                                       padding=(p[0], 0, 0), bias=bias)
        self.out_channels = out_channels
        if any(v != 1 for v in s) or any(kk not in (1, 3) for kk in k) or any(pp != kk // 2 for pp, kk in zip(p, k)):
            self._unsupported = f"kernel={k} stride={s} padding={p}"
        else:
            self._unsupported = None

    def forward_cl(self, xc, out_fp32=False, fold_bias=False):
        """channels-last bf16 in -> channels-last out (bf16, or fp32 when ``out_fp32``).

        A conv bias that feeds a training-mode BatchNorm cancels in the normalisation, so it is not
        added to the stored bf16 tensor (which keeps that tensor centred and its rounding error
        small); it is handed to the BatchNorm kernel instead, where it only shifts ``running_mean``.
        The inner ``spatial_conv -> bn`` pair always does this; with ``fold_bias`` the caller promises
        the same for ``temporal_conv`` and receives ``(y, temporal_bias)``."""
        if self._unsupported:
            raise NotImplementedError("SpatioTemporalConv on B200 supports stride 1 / kernel 1|3 / same padding "
                                      "only, got " + self._unsupported)
        bn_train = self.bn.training or self.bn.running_mean is None
        sb = self.spatial_conv.bias
        if bn_train:
            # bias folded into the BN (if any) and BN statistics accumulated by the conv epilogue
            y1 = ops.ConvFn.apply(xc, self.spatial_conv.weight, None, False, False, True)
            a1, _ = bn_apply(self.bn, y1, 0.0, pre_bias=sb,
                             stats_ready=ops.conv_fuses_stats(self.spatial_conv.out_channels))
        else:
            y1 = ops.ConvFn.apply(xc, self.spatial_conv.weight, sb, False, False)
            a1, _ = bn_apply(self.bn, y1, 0.0)
        tb = self.temporal_conv.bias
        if fold_bias:   # the caller's BatchNorm is in training mode: same treatment for temporal_conv
            return ops.ConvFn.apply(a1, self.temporal_conv.weight, None, out_fp32, False, True), tb
        return ops.ConvFn.apply(a1, self.temporal_conv.weight, tb, out_fp32, False)

    def forward(self, x):
        xc = ops.PackFn.apply(x, 0)
        return ops.UnpackFn.apply(self.forward_cl(xc, out_fp32=True), self.out_channels)
