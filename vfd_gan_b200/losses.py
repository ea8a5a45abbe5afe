"""Loss / init helpers of the hot path (reference: lib/utils.py:51-71,91-92)."""
import torch
import torch.nn as nn

from . import ops


def weights_init(m):
    """Conv3d ~ N(0, 0.02); BatchNorm3d weight ~ N(1, 0.02), bias 0 (lib/utils.py:51-56)."""
    if isinstance(m, nn.Conv3d):
        m.weight.data.normal_(0.0, 0.02)
    elif isinstance(m, nn.BatchNorm3d):
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.fill_(0)
    # writes through .data do not bump Tensor._version: drop the packed bf16 copies of the conv weights explicitly
    ops.invalidate_packed_weights()


def l2_loss(input, target, size_average=True):
    """mean((input - target)^2) (lib/utils.py:59-63); plain torch -- autograd-capable."""
    d = torch.pow(input - target, 2)
    return torch.mean(d) if size_average else d


def weighted_bce(input, target, pos_weight=2):
    """Class-weighted BCE of lib/utils.py:65-71 through the fused CUDA reduction (loss and gradient
    in one pass). ``pos_weight`` multiplies the (1 - target) term exactly like the reference."""
    if pos_weight is None:
        pos_weight = 1
    return ops.WeightedBceFn.apply(input, target, float(pos_weight))


def strip_module_prefix(state_dict):
    """What ``fix_model_state_dict`` (lib/utils.py:15-22) is meant to do -- drop the ``module.`` prefix that
    ``nn.DataParallel`` puts on every key -- without its NameError (it uses an ``OrderedDict`` the module never
    imports). Checkpoints are ``{'epoch': int, 'state_dict': ...}`` files (lib/train_gan.py:52-57)."""
    from collections import OrderedDict
    out = OrderedDict()
    for k, v in state_dict.items():
        out[k[7:] if k.startswith("module.") else k] = v
    return out


def gray2rgb(video):
    return torch.cat([video, video, video], dim=1)
