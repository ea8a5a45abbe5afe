"""GPU: the reference's OWN training loop (``MyGAN`` + ``GANBaseModel.train``, models/mygannet.py:216-366,
lib/train_gan.py:59-85) runs unchanged over ``vfd_gan_b200.compat.install()``.

The unmodified reference sources come from ``oracle/_ref`` (staged by ``oracle/make_ref.py`` in the build
container; git-ignored, shipped to the GPU box). Nothing here reads /root/reference.
"""
import copy
import os
import types

import pytest
import torch

import vfd_gan_b200 as V
from oracle import make_ref
from oracle import vfd_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def ref():
    if not os.path.isdir(os.path.join(make_ref.REF_DST, "models")):
        pytest.skip("oracle/_ref is not staged (run `python oracle/make_ref.py` where /root/reference exists)")
    return make_ref.import_ref(make_ref.REF_DST)


def make_args(tmp_path, batch, nfr, isize, **over):
    """What lib/args.py:8-40 parses, with the reference's defaults for the hyper-parameters."""
    a = types.SimpleNamespace(gpu=[0], ep=1, result_root=str(tmp_path), isize=isize, ich=3, nfr=nfr, batchsize=batch,
                              workers=0, model="mygan", lr=2e-5, beta1=0.5, w_adv=1, w_con=10, pos_weight=2,
                              freq=10 ** 9, resume="", ae=False)
    a.__dict__.update(over)
    return a


class Loader:
    """A stand-in for the reference's DataLoader dict entry: yields (input, real, gt, lb) CPU batches
    (lib/data.py:119-131 order) and lets the test look at the model between batches."""

    def __init__(self, batches, between=None):
        self.batches, self.between = batches, between

    def __len__(self):
        return len(self.batches)

    def __iter__(self):
        for i, b in enumerate(self.batches):
            yield b
            if self.between is not None:
                self.between(i)


def test_reference_mygan_train_loop_over_compat(ref, tmp_path):
    """trainer.py's ``MyGAN(args, dataloader).train()`` with only ``compat.install()`` added: the reference's
    forward_g / forward_d / backward_g (incl. the dead adversarial backward with retain_graph) / backward_d /
    two optim.Adam steps drive the B200 modules. Its logged errors must equal GanTrainStep's on the same
    weights, batches and dropout masks, and the CPU oracle's (same masks injected)."""
    from vfd_gan_b200 import compat, ops
    B, D, S, steps = 2, 16, 64, 3
    batches = []
    for it in range(steps):
        inp, gt, _gf, _pf = O.synthetic_batch(B, D, S, seed=40 + it)
        batches.append((inp, inp.clone(), gt, torch.zeros(B, D)))
    compat.install()
    try:
        mg = ref.mygannet
        torch.manual_seed(0)
        logged, seeds, flows = [], [], []
        model = mg.MyGAN(make_args(tmp_path, B, D, S), {"train": None})
        assert isinstance(model.netg, V.NetG) and isinstance(model.netd, V.NetD)
        assert type(model).optimize_params is mg.MyGAN.optimize_params          # the reference's own step
        netg0, netd0 = copy.deepcopy(model.netg), copy.deepcopy(model.netd)      # weights before training

        def between(i):
            logged.append(dict(model.errors_dict))
            seeds.append(list(model.netg.last_dropout_seeds))
            flows.append(tuple(model.color_video_dict["train/input-real-inflow-genflow"]
                               .split(S, dim=3)[2:4]))     # gt_flow, pre_flow as forward_d computed them
        model.dataloader = {"train": Loader(batches, between)}
        model.train()                                                             # lib/train_gan.py:59-85
        assert model.global_step == steps and len(logged) == steps
    finally:
        compat.uninstall()
    assert int(model.netd.spatdisc.dconv1.bn.num_batches_tracked) == 2 * steps  # NetD runs twice per step
    assert int(model.netg.dconv1.bn.num_batches_tracked) == steps

    # the fused step on the same starting weights, same batches, same dropout masks (seeds), flows in-step
    fused = V.GanTrainStep(netg0, netd0, graph=False)
    sd_g = {k: v.detach().cpu().clone() for k, v in netg0.state_dict().items()}
    sd_d = {k: v.detach().cpu().clone() for k, v in netd0.state_dict().items()}
    oracle = O.OracleTrainer(sd_g, sd_d)
    shapes = [(B, 256, D // 16, S // 16, S // 16), (B, 256, D // 8, S // 8, S // 8), (B, 128, D // 4, S // 4, S // 4),
              (B, 64, D // 2, S // 2, S // 2)]
    for it in range(steps):
        inp, _real, gt, _lb = batches[it]
        fused.step(inp.to(DEV), gt.to(DEV), dropout_seeds=seeds[it])
        got = fused.losses_dict()
        masks = []
        for seed, (N, C, d, h, w) in zip(seeds[it], shapes):       # the Philox masks the kernels drew
            o = torch.empty(N, d, h, w, C, dtype=torch.bfloat16, device=DEV)
            ops.bn_act_fwd(torch.zeros_like(o), torch.zeros(C, device=DEV), torch.ones(C, device=DEV), 1.0, o, None,
                           1, 1, 1, 0.25, seed)
            masks.append(o.float().permute(0, 4, 1, 2, 3).cpu().contiguous())
        gf, pf = (t.detach().cpu().float() for t in flows[it])
        want, _ = oracle.step(inp, gt, gf, pf, dropout_masks=masks)
        for k, w_ in want.items():
            ref_logged = logged[it][k + "/train"]
            tol = 2e-2 if "adv" in k or k == "g/err_g" else 1e-2
            assert abs(ref_logged - got[k]) <= tol * abs(got[k]) + 1e-6, ("compat vs fused", it, k, ref_logged, got[k])
            assert abs(ref_logged - w_) <= tol * abs(w_) + 1e-6, ("compat vs oracle", it, k, ref_logged, w_)


def test_reference_mygan_test_loop_over_compat(ref, tmp_path):
    """``MyGAN.test`` (models/mygannet.py:369-475) over compat: runs on the device path (flow, threshold + opening,
    discriminator) and hands lib/evaluate.py's sklearn metrics finite values."""
    from vfd_gan_b200 import compat
    B, D, S = 2, 16, 64
    batches = []
    for it in range(2):
        inp, gt, _gf, _pf = O.synthetic_batch(B, D, S, seed=60 + it)
        batches.append((inp, inp.clone(), gt, torch.zeros(B, D)))
    compat.install()
    try:
        torch.manual_seed(1)
        model = ref.mygannet.MyGAN(make_args(tmp_path, B, D, S), {"train": Loader(batches[:1]), "test": Loader(batches)})
        model.epoch = 0
        model.test()
    finally:
        compat.uninstall()
    assert all(v == v for v in model.score_dict.values()) and len(model.score_dict) >= 3
    assert any(k.endswith("/test") for k in model.errors_dict)
