"""CPU: the oracle (oracle/vfd_oracle.py) against the golden fixtures produced by the reference's own
modules (tests/golden/make_golden.py). fp32 on both sides -> tight tolerances."""
import torch

from oracle import vfd_oracle as O
from helpers import golden, rel, build_cfg1_nets, build_small_nets


def test_st_conv_forward_backward_matches_reference_fixture():
    f = golden("st_conv_small.pt")
    sd = {("m." + k): v.clone() for k, v in f["sd"].items()}
    for k in sd:
        if sd[k].is_floating_point() and "running" not in k:
            sd[k].requires_grad_(True)
    x = f["x"].clone().requires_grad_(True)
    y = O.st_conv(sd, "m", x, (3, 3, 3), train=True)
    assert torch.allclose(y, f["y"], atol=1e-5, rtol=1e-5)
    y.backward(f["gy"])
    assert rel(x.grad, f["gx"]) < 1e-5
    assert rel(sd["m.spatial_conv.weight"].grad, f["gw_spatial"]) < 1e-5
    assert rel(sd["m.temporal_conv.weight"].grad, f["gw_temporal"]) < 1e-5
    assert rel(sd["m.temporal_conv.bias"].grad, f["gb_temporal"]) < 1e-5
    assert rel(sd["m.bn.weight"].grad, f["g_bn_w"]) < 1e-5
    for k in ("bn.running_mean", "bn.running_var", "bn.num_batches_tracked"):
        assert torch.allclose(sd["m." + k].float(), f["sd_after"][k].float(), atol=1e-6)


def test_convlstm_cell_matches_reference_fixture():
    f = golden("convlstm_cell.pt")
    h, c = O.convlstm_cell(f["sd"], "", f["x"], f["h"], f["c"])
    assert torch.allclose(h, f["h_next"], atol=1e-6) and torch.allclose(c, f["c_next"], atol=1e-6)


def test_losses_match_reference_fixture():
    f = golden("losses.pt")
    assert torch.allclose(O.l2_loss(f["a"], f["b"]), f["l2"])
    assert torch.allclose(O.weighted_bce(f["p"], f["t"]), f["wbce"])
    assert torch.allclose(O.weighted_bce(f["p"], f["t"], 3), f["wbce_pw3"])


def test_small_nets_match_reference_fixture():
    f = golden("netg_netd_small.pt")
    g, xg, sdisc, xs, tdisc, xt = build_small_nets()
    ic = f["init_check"]
    assert torch.equal(g.dconv1.conv.spatial_conv.weight.detach().flatten()[:8], ic["g"])
    assert torch.equal(sdisc.dconv1.conv.spatial_conv.weight.detach().flatten()[:8], ic["s"])
    assert torch.equal(tdisc.linear.weight.detach().flatten()[:8], ic["t"])
    assert torch.equal(xg.flatten()[:8], ic["xg"]) and torch.equal(xs.flatten()[:8], ic["xs"])
    sdg = {k: v.clone() for k, v in g.state_dict().items()}
    pred = O.netg_forward(sdg, xg, True, [1.0] * 4)
    assert torch.allclose(pred, f["predict"], atol=1e-5)
    for k, v in f["g_running"].items():
        assert torch.allclose(sdg[k], v, atol=1e-5)
    s_cls, s_feat = O.sdisc_forward({k: v.clone() for k, v in sdisc.state_dict().items()}, "", xs, True)
    t_cls, t_feat = O.tdisc_forward({k: v.clone() for k, v in tdisc.state_dict().items()}, "", xt, True)
    assert torch.allclose(s_cls, f["s_cls"], atol=1e-5) and rel(s_feat, f["s_feat"]) < 1e-4
    assert torch.allclose(t_cls, f["t_cls"], atol=1e-5) and rel(t_feat, f["t_feat"]) < 1e-4


def test_train_trajectory_matches_reference_fixture():
    """10 optimize_params steps at BASELINE config 1 (B=4, 16x3x64x64): all 12 logged losses."""
    f = golden("step_traj_cfg1.pt")
    netg, netd = build_cfg1_nets()
    assert torch.equal(netg.dconv1.conv.spatial_conv.weight.detach().flatten()[:8], f["init_check"]["g_first"])
    assert torch.equal(netd.spatdisc.linear.weight.detach().flatten()[:8], f["init_check"]["d_lin"])
    tr = O.OracleTrainer(netg.state_dict(), netd.state_dict())
    cfg = f["config"]
    for it, want in enumerate(f["traj"]):
        got, _ = tr.step(*O.synthetic_batch(cfg["B"], cfg["D"], cfg["S"], seed=cfg["data_seed0"] + it),
                         dropout_masks=[1.0] * 4)
        for k in want:
            assert abs(got[k] - want[k]) <= 2e-4 * abs(want[k]) + 1e-6, (it, k, got[k], want[k])


def test_operand_matched_mode_stays_close_to_fp32():
    g, xg, *_ = build_small_nets()
    sd = g.state_dict()
    a = O.netg_forward({k: v.clone() for k, v in sd.items()}, xg, True, [1.0] * 4, round_bf16=True)
    b = O.netg_forward({k: v.clone() for k, v in sd.items()}, xg, True, [1.0] * 4, round_bf16=False)
    assert rel(a, b) < 2e-2
