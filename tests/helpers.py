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
