"""Shared builders for the tests."""
import os
import types

import torch
import torch.nn as nn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


def build_cfg1_nets():
    """Full-size NetG / NetD initialised exactly like tests/golden/make_golden.py (BASELINE config 1):
    same RNG call sequence, NetD Linears re-created for isize=64 after construction."""
    import vfd_gan_b200 as V
    torch.manual_seed(0)
    netg = V.NetG()
    netd = V.NetD(types.SimpleNamespace(nfr=16, isize=128))       # RNG consumption of the reference's NetD
    netd.tempdisc.gpool = nn.AvgPool3d((1, 64, 64), stride=1)
    netd.spatdisc.linear = nn.Linear(32 * 32 * 1, 1)
    netd.tempdisc.linear = nn.Linear(32 * 4 * 2, 1)
    netg.apply(V.weights_init)
    netd.apply(V.weights_init)
    netg.dropout.p = 0.0
    return netg, netd


def build_small_nets():
    """NetG(3,8), SDisc/TDisc(ndf=8) + inputs with the RNG sequence of make_golden.py (seed 14)."""
    import vfd_gan_b200 as V
    torch.manual_seed(14)
    g = V.NetG(3, 8)
    g.apply(V.weights_init)
    g.dropout.p = 0.0
    xg = torch.rand(2, 3, 16, 32, 32) * 2 - 1
    sdisc = V.SDisc(3, 16, ndf=8, kernel=(1, 3, 3), padding=(0, 1, 1))
    tdisc = V.TDisc(3, 32, ndf=8, kernel=(3, 1, 1), padding=(1, 0, 0))
    sdisc.apply(V.weights_init)
    tdisc.apply(V.weights_init)
    xs = torch.rand(1, 3, 16, 128, 128)
    xt = torch.rand(2, 3, 16, 32, 32) * 2 - 1
    return g, xg, sdisc, xs, tdisc, xt


# per-clip contrast of the 16 scoring clips of composed_small.pt (tests/golden/make_golden.py SCORE_AMPS)
SCORE_AMPS = [0.1 + 0.9 * k / 15 for k in (7, 0, 12, 3, 15, 9, 1, 5, 10, 14, 2, 6, 4, 13, 8, 11)]


def score_batch(b):
    gen = torch.Generator().manual_seed(500 + b)
    xb = torch.rand(4, 3, 16, 32, 32, generator=gen) * 2 - 1
    return xb * torch.tensor([SCORE_AMPS[4 * b + j] for j in range(4)]).view(4, 1, 1, 1, 1)


def build_lstm_net():
    """NetGLstm(3, 8, isize=32) + input with the RNG sequence of make_golden.composed_fixture (seed 21)."""
    import vfd_gan_b200 as V
    torch.manual_seed(21)
    g = V.NetGLstm(3, 8, isize=32)
    g.apply(V.weights_init)
    g.dropout.p = 0.0
    x = torch.rand(2, 3, 32, 32, 32) * 2 - 1
    return g, x


def build_enc_dec_enc():
    """EncDecEncG(3, 8) with the RNG sequence of make_golden.composed_fixture (seed 22)."""
    import vfd_gan_b200 as V
    torch.manual_seed(22)
    m = V.EncDecEncG(3, 8)
    m.apply(V.weights_init)
    m.netg.dropout.p = 0.0
    return m


def build_stcnn():
    """AutoEncoder + (clip, mask) with the RNG sequence of make_golden.stcnn_fixture (seed 15 / 600)."""
    import vfd_gan_b200 as V
    torch.manual_seed(15)
    m = V.AutoEncoder()
    m.apply(V.weights_init)
    for blk in m.children():
        if hasattr(blk, "dropout"):
            blk.dropout.p = 0.0
    gen = torch.Generator().manual_seed(600)
    x = torch.rand(2, 3, 16, 32, 32, generator=gen) * 2 - 1
    gt = (torch.rand(2, 1, 16, 32, 32, generator=gen) > 0.9).float()
    return m, x, gt


def flow_clip(B, D, S, seed):
    """Seeded smooth clip in [-1, 1] -- same generator as tests/golden/make_golden.py flow_clip."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(B, 3, D, S + 16, S + 16, generator=g)
    vid = torch.nn.functional.avg_pool3d(base, (1, 9, 9), stride=1, padding=(0, 4, 4))[:, :, :, 8:8 + S, 8:8 + S]
    vid = (vid - vid.min()) / (vid.max() - vid.min())
    return (vid * 0.8 + 0.2 * torch.rand(B, 3, D, 1, 1, generator=g)) * 2 - 1


def flow_level_agreement(levels, want):
    """(fraction of bytes equal, fraction within one grey level modulo 256) between two uint8 level tensors."""
    d = (levels.int() - want.int()).abs()
    d = torch.minimum(d, 256 - d)
    return float((d == 0).float().mean()), float((d <= 1).float().mean())


def build_headline_nets(dropout=0.0):
    """Full-size NetG / NetD for the headline geometry (16 x 112 x 112) with the RNG call sequence of
    tests/golden/make_golden_headline.py::reference_nets_112."""
    import vfd_gan_b200 as V
    torch.manual_seed(0)
    netg = V.NetG()
    netd = V.NetD(types.SimpleNamespace(nfr=16, isize=128))       # RNG consumption of the reference's NetD
    netd.tempdisc.gpool = nn.AvgPool3d((1, 112, 112), stride=1)
    netd.spatdisc.linear = nn.Linear(32 * 32 * 1, 1)
    netd.tempdisc.linear = nn.Linear(32 * 4 * 2, 1)
    netg.apply(V.weights_init)
    netd.apply(V.weights_init)
    netg.dropout.p = dropout
    return netg, netd


def dropout_masks_from_seeds(seeds, B, D, S, device, p=0.25):
    """The Philox keep-masks (scaled by 1/(1-p)) NetG's four decoder dropouts drew for ``seeds``, as fp32 NCDHW CPU
    tensors the oracle accepts as ``dropout_masks`` (the fused BN+act kernel run on zeros with shift 1, slope 1)."""
    from vfd_gan_b200 import ops
    shapes = [(B, 256, D // 16, S // 16, S // 16), (B, 256, D // 8, S // 8, S // 8), (B, 128, D // 4, S // 4, S // 4),
              (B, 64, D // 2, S // 2, S // 2)]
    masks = []
    for seed, (N, C, d, h, w) in zip(seeds, shapes):
        o = torch.empty(N, d, h, w, C, dtype=torch.bfloat16, device=device)
        ops.bn_act_fwd(torch.zeros_like(o), torch.zeros(C, device=device), torch.ones(C, device=device), 1.0, o, None,
                       1, 1, 1, p, seed)
        masks.append(o.float().permute(0, 4, 1, 2, 3).cpu().contiguous())
    return masks
