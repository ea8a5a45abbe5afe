"""Checkpoint save / resume of the training loop.

Reference: ``GANBaseModel.save_weights`` (lib/train_gan.py:52-57) writes ``{'epoch', 'state_dict'}`` files
``<weights>/<name_head>_ep%04d_netG.pth`` / ``..._netD.pth``; ``MyGAN.__init__`` (models/mygannet.py:245-256) resumes
from ``args.resume`` = the generator file and derives the discriminator's name as
``resume.rsplit("_", 1)[0] + "netD.pth"`` -- which drops the underscore (``..._ep0003netD.pth``), so the reference's
resume never finds the file it saved, and ``fix_model_state_dict`` raises ``NameError`` before that (lib/utils.py:15-22).
It also restores the weights only: Adam's moments and step counts restart from zero.

``save_weights`` / ``load_weights`` keep the reference's file format and names (files are interchangeable both ways)
and make the resume work; ``save_training_state`` / ``load_training_state`` add what a faithful continuation needs:
both Adam states, the epoch and the device dropout counter of the fused step."""
import os

import torch

from . import ops
from .losses import strip_module_prefix


def weight_paths(weight_dir, name_head, epoch):
    """File names of lib/train_gan.py:54-57."""
    return ('%s/%s_ep%04d_netG.pth' % (weight_dir, name_head, epoch),
            '%s/%s_ep%04d_netD.pth' % (weight_dir, name_head, epoch))


def save_weights(weight_dir, name_head, epoch, netg, netd):
    """lib/train_gan.py:52-57: ``{'epoch': epoch + 1, 'state_dict': ...}`` for the generator and the discriminator."""
    os.makedirs(weight_dir, exist_ok=True)
    g_path, d_path = weight_paths(weight_dir, name_head, epoch)
    torch.save({'epoch': epoch + 1, 'state_dict': netg.state_dict()}, g_path)
    torch.save({'epoch': epoch + 1, 'state_dict': netd.state_dict()}, d_path)
    return g_path, d_path


def discriminator_path(g_resume):
    """The discriminator file that belongs to a generator checkpoint: the name ``save_weights`` wrote
    (``..._netD.pth``), else the name models/mygannet.py:249 derives (``...netD.pth``, underscore lost)."""
    stem = g_resume.rsplit("_", 1)[0]
    for cand in (stem + "_netD.pth", stem + "netD.pth"):
        if os.path.exists(cand):
            return cand
    raise IOError(f"Model weights not found: neither {stem}_netD.pth nor {stem}netD.pth exists")


def load_weights(g_resume, netg, netd=None, map_location=None):
    """models/mygannet.py:245-256 made to work: load ``state_dict`` (with or without DataParallel's ``module.``
    prefix) into ``netg`` and, when given, the matching discriminator file into ``netd``. Returns the stored epoch."""
    if not os.path.exists(g_resume):
        raise IOError(f"Model weights not found: {g_resume}")
    map_location = map_location or next(netg.parameters()).device
    ck = torch.load(g_resume, map_location=map_location)
    netg.load_state_dict(strip_module_prefix(ck['state_dict']))
    if netd is not None:
        dk = torch.load(discriminator_path(g_resume), map_location=map_location)
        netd.load_state_dict(strip_module_prefix(dk['state_dict']))
    ops.invalidate_packed_weights()      # the bf16 GEMM operands are stale copies of the old masters
    return int(ck.get('epoch', 0))


def save_training_state(path, trainer, epoch, extra=None):
    """Everything ``GanTrainStep`` needs to continue as if it had not stopped: both nets (incl. BatchNorm running
    statistics), both Adam states (moments + step counts), the epoch, and the device-side dropout counter."""
    state = {
        'epoch': int(epoch),
        'netg': trainer.netg.state_dict(), 'netd': trainer.netd.state_dict(),
        'optimizer_g': trainer.optimizer_g.state_dict(), 'optimizer_d': trainer.optimizer_d.state_dict(),
        'step_counter': None if trainer._step_counter is None else int(trainer._step_counter.item()),
        'extra': extra,
    }
    tmp = path + ".tmp"
    torch.save(state, tmp)
    os.replace(tmp, path)                # a crash mid-write never leaves a truncated checkpoint behind
    return path


def load_training_state(path, trainer, map_location=None):
    """Inverse of ``save_training_state``. The optimizer states are replaced by new tensors, so a CUDA graph captured
    before the call would keep updating the old ones: the captured step is dropped and re-recorded on the next calls."""
    map_location = map_location or next(trainer.netg.parameters()).device
    state = torch.load(path, map_location=map_location)
    trainer.netg.load_state_dict(strip_module_prefix(state['netg']))
    trainer.netd.load_state_dict(strip_module_prefix(state['netd']))
    trainer.optimizer_g.load_state_dict(state['optimizer_g'])
    trainer.optimizer_d.load_state_dict(state['optimizer_d'])
    if trainer._step_counter is not None and state.get('step_counter') is not None:
        trainer._step_counter.fill_(state['step_counter'])
    trainer._graph, trainer._static_in, trainer._eager_steps = None, None, 0
    ops.invalidate_packed_weights()
    return int(state['epoch']), state.get('extra')
