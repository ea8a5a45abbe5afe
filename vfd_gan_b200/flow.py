"""``video_to_flow`` on the device (reference: lib/utils.py:94-129, host cv2 Farneback + HSV encoding).

The reference calls it twice per training step on tensors it first moves to the host (models/mygannet.py:281-282)
-- once on the generator's fresh output, which serialises the whole step behind one CPU core. Here it is a chain
of small CUDA kernels on the current stream (no host synchronisation, CUDA-graph capturable); the arithmetic
follows OpenCV's closely enough that > 98 % of the output bytes equal the reference's and the rest differ by one
grey level (tests/test_flow_gpu.py, fixture produced by the reference function itself).
"""
import torch

from . import _lib, ops

_ws_cache = {}


def _workspace(B, D, H, W, device):
    need = int(_lib.lib().vfd_video_to_flow_workspace(B, D, H, W))
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def video_to_flow(video, return_raw=False):
    """video fp32 (B, 3, D, H, W) in [-1, 1] on the GPU -> flow video fp32 (B, 3, D, H, W) in [-1, 1] (on the GPU:
    the reference returns a CPU tensor that every caller immediately moves with ``.to('cuda')``).
    ``return_raw`` also returns the Farneback fields (B, D-1, H, W, 2)."""
    if not video.is_cuda:
        raise RuntimeError("video_to_flow: vfd_gan_b200 has no CPU path (the reference's host version is lib/utils.py:94)")
    if video.dim() != 5 or video.shape[1] != 3:
        raise RuntimeError(f"video_to_flow expects (B, 3, D, H, W), got {tuple(video.shape)}")
    B, _, D, H, W = video.shape
    if D < 2:
        raise RuntimeError("video_to_flow needs at least two frames")   # the reference fails too (rgb undefined)
    v = video.detach().contiguous().float()
    out = torch.empty_like(v)
    raw = torch.empty(B, D - 1, H, W, 2, dtype=torch.float32, device=v.device) if return_raw else None
    ops.video_to_flow_op(v, out, raw, _workspace(B, D, H, W, v.device))
    return (out, raw) if return_raw else out
