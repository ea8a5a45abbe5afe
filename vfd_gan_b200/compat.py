"""Drop the B200 path into an unmodified checkout of the reference: ``install()`` rebinds the names the reference's
own modules resolve at call time, so ``trainer.py`` / ``lib/train_gan.py`` / ``test.py`` run unchanged.

    import sys; sys.path.insert(0, "/path/to/vfd_gan")
    import vfd_gan_b200.compat as compat
    compat.install()                  # before trainer.py builds its model
    # ... trainer.py as it is: MyGAN(args, dataloader).train()

What is rebound (reference file:line of the definition that gets shadowed):
  models.spatiotempconv.SpatioTemporalConv                        models/spatiotempconv.py:7
  models.mygannet.{SpatioTemporalConv, NetgConv, NetG, NetdConv, SDisc, TDisc, NetD}   models/mygannet.py:10,13-213
  models.mygannet.{weighted_bce, weights_init, video_to_flow, threshold, morphology_proc, fix_model_state_dict} and the same in
  lib.utils (fix_model_state_dict raises NameError in the reference)   lib/utils.py:15-22,65-71,94-129,139-152
  models.convlstm.{ConvLSTMCell, ConvLSTM}                        models/convlstm.py:6-169
  models.mystcnn.{C2plus1d_Block, AutoEncoder}                    models/mystcnn.py:6-88
``l2_loss`` and ``gray2rgb`` stay the reference's (plain torch on tensors our modules return). ``uninstall()``
restores the originals.
"""
import importlib

_saved = []


def _rebind(module, name, value):
    _saved.append((module, name, getattr(module, name, None), hasattr(module, name)))
    setattr(module, name, value)


def install(device_flow=True, device_morphology=True):
    """Rebind the reference's hot-path names to the B200 implementations. The reference must be importable
    (its root on ``sys.path``). ``device_flow`` / ``device_morphology`` = False keep the host cv2 versions."""
    import vfd_gan_b200 as V
    if _saved:
        return
    stc = importlib.import_module("models.spatiotempconv")
    mg = importlib.import_module("models.mygannet")
    cl = importlib.import_module("models.convlstm")
    ms = importlib.import_module("models.mystcnn")
    lu = importlib.import_module("lib.utils")
    _rebind(stc, "SpatioTemporalConv", V.SpatioTemporalConv)
    for name in ("SpatioTemporalConv", "NetgConv", "NetG", "NetdConv", "SDisc", "TDisc", "NetD"):
        _rebind(mg, name, getattr(V, name))
    for name in ("ConvLSTMCell", "ConvLSTM"):
        _rebind(cl, name, getattr(V, name))
    for name in ("C2plus1d_Block", "AutoEncoder"):
        _rebind(ms, name, getattr(V, name))
    repl = {"weighted_bce": V.weighted_bce, "fix_model_state_dict": V.strip_module_prefix,
            "weights_init": V.weights_init}   # same initialiser + invalidation of the packed bf16 weight copies
    if device_flow:
        repl["video_to_flow"] = V.video_to_flow
    if device_morphology:
        repl["threshold"] = V.evaluate.threshold
        repl["morphology_proc"] = V.evaluate.morphology_proc
    for name, fn in repl.items():
        _rebind(lu, name, fn)
        _rebind(mg, name, fn)            # ``from lib.utils import *`` copied the names into models.mygannet


def uninstall():
    while _saved:
        module, name, old, had = _saved.pop()
        if had:
            setattr(module, name, old)
        else:
            delattr(module, name)
