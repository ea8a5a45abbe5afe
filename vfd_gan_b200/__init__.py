"""vfd_gan_b200 -- B200-native (sm_100a) implementation of the vfd_gan ``mygan`` training hot path.

Public surface mirrors the reference's modules (models/spatiotempconv.py, models/mygannet.py,
models/convlstm.py, lib/utils.py losses) so it drops into trainer.py / lib/train_gan.py unchanged;
see INTEGRATION.md. Compute happens only in libvfd_b200.so (include/vfd_b200.h) -- there is no CPU
or eager-PyTorch fallback for the kernels.
"""
from . import _lib, ops  # noqa: F401
from .spatiotempconv import SpatioTemporalConv  # noqa: F401
from .mygannet import NetgConv, NetG, NetdConv, SDisc, TDisc, NetD  # noqa: F401
from .convlstm import ConvLSTMCell, ConvLSTM  # noqa: F401
from .losses import weights_init, l2_loss, weighted_bce, gray2rgb, strip_module_prefix  # noqa: F401
from .composed import NetGLstm, Encoder, EncDecEncG, AnomalyScorer, anomaly_scores, latent_l2_and_scores  # noqa: F401
from .stcnn import C2plus1d_Block, AutoEncoder, StcnnTrainStep  # noqa: F401
from . import evaluate, data, checkpoint  # noqa: F401
from .data import ClipPrefetcher, DeviceTestTransform  # noqa: F401
from .flow import video_to_flow  # noqa: F401
from .step import GanTrainStep, HostBatchStep, GradAllReducer, LOSS_KEYS  # noqa: F401

__all__ = ["SpatioTemporalConv", "NetgConv", "NetG", "NetdConv", "SDisc", "TDisc", "NetD", "ConvLSTMCell",
           "ConvLSTM", "weights_init", "l2_loss", "weighted_bce", "gray2rgb", "strip_module_prefix", "GanTrainStep", "HostBatchStep",
           "GradAllReducer", "LOSS_KEYS", "NetGLstm", "Encoder", "EncDecEncG", "AnomalyScorer", "anomaly_scores",
           "latent_l2_and_scores", "C2plus1d_Block", "AutoEncoder", "StcnnTrainStep", "evaluate", "video_to_flow", "data", "checkpoint", "ClipPrefetcher", "DeviceTestTransform"]
