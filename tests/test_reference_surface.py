"""CPU: the nn.Module surface is a drop-in for the reference's (names, shapes, init, state_dict)."""
import types

import pytest
import torch

import vfd_gan_b200 as V
from oracle import vfd_oracle as O

G_TABLE = {"dconv1": (3, 21, 32), "dconv2": (32, 115, 64), "dconv3": (64, 230, 128), "dconv4": (128, 460, 256),
           "dconv5": (256, 921, 512), "uconv5": (512, 658, 256), "uconv4": (512, 658, 256), "uconv3": (384, 345, 128),
           "uconv2": (192, 172, 64), "uconv1": (96, 86, 32)}
S_TABLE = {1: (3, 14, 32), 2: (32, 52, 64), 3: (64, 104, 128), 4: (128, 209, 256), 5: (256, 418, 512),
           6: (512, 837, 1024)}
T_TABLE = {1: (3, 2, 32), 2: (32, 27, 64), 3: (64, 54, 128)}


def test_state_dict_layout_matches_survey_appendix_a():
    g = V.NetG()
    sd = g.state_dict()
    assert sum(p.numel() for p in g.parameters()) == 13_527_885
    for b, (cin, m, cout) in G_TABLE.items():
        assert tuple(sd[f"{b}.conv.spatial_conv.weight"].shape) == (m, cin, 1, 3, 3)
        assert tuple(sd[f"{b}.conv.spatial_conv.bias"].shape) == (m,)
        assert tuple(sd[f"{b}.conv.temporal_conv.weight"].shape) == (cout, m, 3, 1, 1)
        for s in ("weight", "bias", "running_mean", "running_var"):
            assert tuple(sd[f"{b}.conv.bn.{s}"].shape) == (m,) and tuple(sd[f"{b}.bn.{s}"].shape) == (cout,)
        assert sd[f"{b}.bn.num_batches_tracked"].dtype == torch.int64
    assert tuple(sd["conv_last.weight"].shape) == (1, 32, 3, 3, 3)
    d = V.NetD(types.SimpleNamespace(nfr=16, isize=128))
    sdd = d.state_dict()
    assert sum(p.numel() for p in d.parameters()) == 6_324_353
    for i, (cin, m, cout) in S_TABLE.items():
        assert tuple(sdd[f"spatdisc.dconv{i}.conv.spatial_conv.weight"].shape) == (m, cin, 1, 3, 3)
        assert tuple(sdd[f"spatdisc.dconv{i}.conv.temporal_conv.weight"].shape) == (cout, m, 1, 1, 1)
    for i, (cin, m, cout) in T_TABLE.items():
        assert tuple(sdd[f"tempdisc.dconv{i}.conv.spatial_conv.weight"].shape) == (m, cin, 1, 1, 1)
        assert tuple(sdd[f"tempdisc.dconv{i}.conv.temporal_conv.weight"].shape) == (cout, m, 3, 1, 1)
    assert tuple(sdd["spatdisc.linear.weight"].shape) == (1, 4096)
    assert tuple(sdd["tempdisc.linear.weight"].shape) == (1, 256)
    assert O.intermed_channels(96, 32, (3, 3, 3)) == 86


def test_generalised_discriminator_heads():
    d = V.NetD(types.SimpleNamespace(nfr=16, isize=112))
    assert d.spatdisc.linear.in_features == 1024 and d.tempdisc.linear.in_features == 256
    d = V.NetD(types.SimpleNamespace(nfr=32, isize=128))
    assert d.spatdisc.linear.in_features == 4096 and d.tempdisc.linear.in_features == 512


def test_convlstm_surface():
    cell = V.ConvLSTMCell((8, 8), 16, 32, (3, 3), True)
    assert set(cell.state_dict()) == {"conv.weight", "conv.bias"}
    assert tuple(cell.conv.weight.shape) == (128, 48, 3, 3)
    m = V.ConvLSTM((8, 8), 512, 512, (3, 3), 1, batch_first=True, bias=False)
    assert list(m.state_dict()) == ["cell_list.0.conv.weight"]
    with pytest.raises(ValueError):
        V.ConvLSTM((8, 8), 4, 4, 3, 1)
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 1, 512, 8, 8), hidden_state=[None])


def test_unsupported_configs_fail_loudly():
    m = V.SpatioTemporalConv(4, 8, 3, stride=2, padding=1)
    with pytest.raises(NotImplementedError):
        m.forward_cl(None)
    with pytest.raises(ValueError):
        V.NetG(3, 4)
    with pytest.raises((RuntimeError, NotImplementedError)):     # CPU tensors: there is no CPU path
        V.NetG()(torch.zeros(1, 3, 16, 16, 16))


def test_same_seed_same_state_dict_as_reference(reference_modules):
    r = reference_modules
    torch.manual_seed(0)
    rg = r.mygannet.NetG()
    rg.apply(r.utils.weights_init)
    torch.manual_seed(0)
    mg = V.NetG()
    mg.apply(V.weights_init)
    a, b = rg.state_dict(), mg.state_dict()
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
    args = types.SimpleNamespace(nfr=16, isize=128)
    torch.manual_seed(1)
    rd = r.mygannet.NetD(args)
    rd.apply(r.utils.weights_init)
    torch.manual_seed(1)
    md = V.NetD(args)
    md.apply(V.weights_init)
    a, b = rd.state_dict(), md.state_dict()
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
    mg.load_state_dict(rg.state_dict())            # both directions
    rg.load_state_dict(mg.state_dict())


def test_oracle_is_bit_exact_against_reference_modules(reference_modules):
    r = reference_modules
    torch.manual_seed(0)
    rg = r.mygannet.NetG()
    rg.apply(r.utils.weights_init)
    rg.train()
    sd = {k: v.clone() for k, v in rg.state_dict().items()}
    x = torch.rand(1, 3, 16, 32, 32) * 2 - 1
    torch.manual_seed(5)
    want = rg(x)                                    # dropout active: F.dropout draws in the same order
    torch.manual_seed(5)
    got = O.netg_forward(sd, x, True)
    assert torch.equal(want, got)
    assert all(torch.allclose(rg.state_dict()[k].float(), sd[k].float(), atol=1e-6) for k in sd)
    args = types.SimpleNamespace(nfr=16, isize=128)
    torch.manual_seed(1)
    rd = r.mygannet.NetD(args)
    rd.apply(r.utils.weights_init)
    rd.train()
    sdd = {k: v.clone() for k, v in rd.state_dict().items()}
    xx, yy = torch.rand(1, 3, 16, 128, 128), torch.rand(1, 3, 16, 128, 128) * 2 - 1
    for a, b in zip(rd(xx, yy), O.netd_forward(sdd, xx, yy, True)):
        assert torch.equal(a, b)
    cell = r.convlstm.ConvLSTMCell((8, 8), 16, 32, (3, 3), True)
    xi, h, c = torch.randn(2, 16, 8, 8), torch.randn(2, 32, 8, 8), torch.randn(2, 32, 8, 8)
    for a, b in zip(cell(xi, (h, c)), O.convlstm_cell(cell.state_dict(), "", xi, h, c)):
        assert torch.allclose(a, b, atol=1e-6)
    p, t = torch.rand(2, 1, 4, 8, 8), (torch.rand(2, 1, 4, 8, 8) > 0.5).float()
    assert torch.equal(r.utils.weighted_bce(p, t), O.weighted_bce(p, t))
    assert torch.equal(r.utils.l2_loss(p, t), O.l2_loss(p, t))


def test_compat_install_rebinds_the_reference_names(reference_modules):
    """vfd_gan_b200.compat.install() is the 'two import lines' of INTEGRATION.md done programmatically: the
    reference's own modules then build the B200 nets, and uninstall() restores them."""
    import vfd_gan_b200 as V
    from vfd_gan_b200 import compat
    mg, lu = reference_modules.mygannet, reference_modules.utils
    ref_netg, ref_flow = mg.NetG, lu.video_to_flow
    compat.install()
    try:
        assert mg.NetG is V.NetG and mg.NetD is V.NetD and mg.SpatioTemporalConv is V.SpatioTemporalConv
        assert reference_modules.spatiotempconv.SpatioTemporalConv is V.SpatioTemporalConv
        assert reference_modules.convlstm.ConvLSTMCell is V.ConvLSTMCell
        assert mg.video_to_flow is V.video_to_flow and lu.morphology_proc is V.evaluate.morphology_proc
        assert mg.weighted_bce is V.weighted_bce
        net = mg.NetG()                               # what MyGAN.__init__ does (models/mygannet.py:232)
        assert isinstance(net, V.NetG)
        net.apply(lu.weights_init)                    # the reference's own initialiser works on our modules
    finally:
        compat.uninstall()
    assert mg.NetG is ref_netg and lu.video_to_flow is ref_flow


def test_checkpoint_files_load_both_ways(reference_modules, tmp_path):
    """Checkpoints are ``{'epoch', 'state_dict'}`` files (lib/train_gan.py:52-57), possibly with DataParallel's
    ``module.`` prefix (lib/utils.py:15-22; test.py:117-120 loads them into NetG()). A file written from the
    reference's nets loads into ours and back, bit for bit."""
    import types
    import vfd_gan_b200 as V
    mg = reference_modules.mygannet
    torch.manual_seed(4)
    rg, rd = mg.NetG(), mg.NetD(types.SimpleNamespace(nfr=16, isize=128))
    path_g, path_d = tmp_path / "roc_ep0003_netG.pth", tmp_path / "roc_ep0003_netD.pth"
    torch.save({"epoch": 4, "state_dict": {"module." + k: v for k, v in rg.state_dict().items()}}, path_g)
    torch.save({"epoch": 4, "state_dict": rd.state_dict()}, path_d)
    og, od = V.NetG(), V.NetD(types.SimpleNamespace(nfr=16, isize=128))
    og.load_state_dict(V.strip_module_prefix(torch.load(path_g)["state_dict"]))
    od.load_state_dict(V.strip_module_prefix(torch.load(path_d)["state_dict"]))
    for k, v in rg.state_dict().items():
        assert torch.equal(og.state_dict()[k], v), k
    for k, v in rd.state_dict().items():
        assert torch.equal(od.state_dict()[k], v), k
    torch.save({"epoch": 5, "state_dict": og.state_dict()}, path_g)          # and back into the reference's NetG
    rg2 = mg.NetG()
    rg2.load_state_dict(torch.load(path_g)["state_dict"])
    assert all(torch.equal(rg2.state_dict()[k], v) for k, v in rg.state_dict().items())
