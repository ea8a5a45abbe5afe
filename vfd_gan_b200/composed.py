"""Builder-defined compositions of the reference's modules that BASELINE.json's configs 3 and 5 name but the
reference does not ship. Each is made only of reference modules, so the
oracle (and the golden fixtures) are the same compositions written with the reference's own classes:

* ``NetGLstm``    -- ``NetG`` with a ``ConvLSTM`` over the latent (time = the latent's depth axis), the wrapping
                     idiom of models/convlstm.py:199-201;
* ``Encoder``     -- a second copy of NetG's ``dconv1..dconv5`` (models/mygannet.py:35-39,57-71);
* ``EncDecEncG``  -- NetG -> gray2rgb(predict) -> Encoder, returning ``(predict, latent_i, latent_o)`` like the
                     2-D GANomaly generator (models/ganomaly.py:160-175);
* ``AnomalyScorer`` -- the test-time score of models/ganomaly.py:372,396: per-clip mean squared latent
                     difference (over every non-batch dim), min-max scaled over the sweep.
"""
import torch
import torch.distributed as dist
import torch.nn as nn

from . import ops
from .convlstm import ConvLSTM
from .mygannet import NetG, NetgConv


class NetGLstm(NetG):
    """``NetG`` + ``ConvLSTM(input_size=(isize/16, isize/16), input_dim=hidden_dim=16*ngf, kernel (3,3), 1 layer,
    batch_first, bias=False)`` applied to ``latent_i.transpose(1, 2)`` and transposed back."""

    def __init__(self, nc=3, ngf=32, isize=128):
        super().__init__(nc, ngf)
        s = isize // 16
        self.clstm = ConvLSTM(input_size=(s, s), input_dim=ngf * 16, hidden_dim=ngf * 16, kernel_size=(3, 3),
                              num_layers=1, batch_first=True, bias=False)

    def bottleneck_cl(self, latent):
        return self.clstm.forward_cl(latent)


class Encoder(nn.Module):
    """dconv1..dconv5 of NetG with 2x2x2 average pools between them; clip -> latent (B, 16*ngf, D/16, H/16, W/16)."""

    def __init__(self, nc=3, ngf=32):
        super().__init__()
        self.dconv1 = NetgConv(nc, ngf)
        self.dconv2 = NetgConv(ngf, ngf * 2)
        self.dconv3 = NetgConv(ngf * 2, ngf * 4)
        self.dconv4 = NetgConv(ngf * 4, ngf * 8)
        self.dconv5 = NetgConv(ngf * 8, ngf * 16)
        self.avgpool = nn.AvgPool3d(2)
        self.out_channels = ngf * 16

    encode_cl = NetG.encode_cl

    def forward_cl(self, xc):
        return self.encode_cl(xc)

    def forward(self, x):
        return ops.UnpackFn.apply(self.encode_cl(ops.PackFn.apply(x, 0)), self.out_channels)


class EncDecEncG(nn.Module):
    """enc-dec-enc generator: ``forward(x) -> (predict, latent_i, latent_o)``."""

    def __init__(self, nc=3, ngf=32, netg=None):
        super().__init__()
        self.netg = netg if netg is not None else NetG(nc, ngf)
        self.encoder2 = Encoder(3, self.netg.ngf)
        self.latent_channels = self.netg.ngf * 16

    def forward_cl(self, xc, dropout_seeds=None):
        logits, latent_i = self.netg.forward_cl(xc, dropout_seeds)
        predict = ops.SigmoidHeadFn.apply(logits)
        latent_o = self.encoder2.encode_cl(ops.PackFn.apply(predict, 3))      # gray2rgb folded into the pack
        return predict, latent_i, latent_o

    def forward(self, x):
        predict, li, lo = self.forward_cl(ops.PackFn.apply(x, 0))
        c = self.latent_channels
        return predict, ops.UnpackFn.apply(li, c), ops.UnpackFn.apply(lo, c)


def latent_l2_and_scores(latent_i, latent_o, valid_channels):
    """-> (l_enc = l2_loss(latent_o, latent_i) with gradients, per-clip scores fp32 [B]) from channels-last
    bf16 latents, one fused reduction."""
    return ops.LatentL2Fn.apply(latent_o, latent_i, valid_channels)


def anomaly_scores(latent_i, latent_o, valid_channels):
    """Per-clip anomaly score, fp32 [B] (models/ganomaly.py:372 over every non-batch dim)."""
    with torch.no_grad():
        return ops.LatentL2Fn.apply(latent_o, latent_i, valid_channels)[1]


def gather_scores(local, group=None):
    """Concatenate the per-rank score vectors in rank order (equal shard sizes). Works on whatever backend the
    process group uses (NCCL on the GPUs, gloo in the CPU tests)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    out = torch.empty(world * local.numel(), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


class AnomalyScorer:
    """Inference sweep of BASELINE config 5: ``score_batch`` per batch, ``finish`` gathers every rank's scores
    and applies the global min-max scaling of models/ganomaly.py:396."""

    def __init__(self, model, group=None):
        self.model = model
        self.group = group
        self.chunks = []

    @torch.no_grad()
    def score_batch(self, clips):
        """clips fp32 (B,3,D,H,W) on the GPU -> raw scores fp32 [B] (device, no sync)."""
        _, li, lo = self.model.forward_cl(ops.PackFn.apply(clips, 0))
        s = anomaly_scores(li, lo, self.model.latent_channels)
        self.chunks.append(s)
        return s

    @torch.no_grad()
    def finish(self):
        """-> (scaled scores over all ranks in rank-major order, raw scores)."""
        if not self.chunks:
            raise RuntimeError("AnomalyScorer.finish(): no batch was scored")
        local = torch.cat(self.chunks)
        self.chunks = []
        raw = gather_scores(local, self.group)
        mm = torch.stack([raw.min(), raw.max()])
        out = torch.empty_like(raw)
        ops.score_scale(raw, mm, out)
        return out, raw
