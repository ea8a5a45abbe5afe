"""Test infrastructure (CPU oracle, never imported by vfd_gan_b200/): the deterministic part of the reference's clip
pipeline restated over numpy arrays, each function citing the reference lines it follows. Pinned against the
reference's own ``videotransforms`` classes in tests/test_clip_oracle.py (bit-exact). The arithmetic lives in Pillow
(``Image.resize``: src/libImaging/Resample.c; requirements.txt pins Pillow 6.2.1, this image has 12.2 -- the bilinear
8-bit resample has not changed) and in torch (float32 division)."""
import numpy as np
import PIL.Image
import torch


def resize_frames(frames, size):
    """videotransforms/video_transforms.py:91-110 -> functional.py:43-58 on PIL images: ``Resize(size)`` keeps its
    default interpolation 'nearest', which the PIL branch maps to ``PIL.Image.BILINEAR`` (:54-57, the two names are
    swapped there), and calls ``img.resize((size[1], size[0]), pil_inter)`` per frame.
    frames uint8 (n, H, W, C) with C in {1, 3} -> uint8 (n, size[0], size[1], C)."""
    out = []
    for f in np.asarray(frames):
        img = PIL.Image.fromarray(f[..., 0] if f.shape[-1] == 1 else f)       # lib/data.py:108 (np.uint8 frame)
        r = np.array(img.resize((int(size[1]), int(size[0])), PIL.Image.BILINEAR))
        out.append(r[..., None] if r.ndim == 2 else r)
    return np.stack(out)


def clip_to_tensor(frames, channel_nb=3):
    """videotransforms/volume_transforms.py:17-58 (``ClipToTensor``, div_255): frames uint8 (T, H, W, C) -> float32
    (channel_nb, T, H, W): the frames are written into a float64 array (an (H, W) 'L' frame is broadcast over the
    channels, :44-46 with utils/images.py:4-12), converted to float32 and divided by 255 in float32."""
    frames = np.asarray(frames)
    t, h, w, c = frames.shape
    np_clip = np.zeros([channel_nb, t, h, w])
    for i in range(t):
        img = frames[i]
        img = img.transpose(2, 0, 1) if c > 1 else img[..., 0][None]
        np_clip[:, i] = img
    return torch.from_numpy(np_clip).float().div(255)


def mdf_item(data_u8, mask_u8=None):
    """What ``MdfDataLoader.__getitem__`` returns for already decoded and resized frames (lib/data.py:56-78):
    data / real ``*2-1``; the mask goes through the 3-channel ClipToTensor together with the RGB frames (:62-66), keeps
    channel 0 and is NOT rescaled (:78)."""
    data = clip_to_tensor(data_u8) * 2 - 1
    if mask_u8 is None:
        return data, torch.zeros((1,) + tuple(data.shape[1:]))               # "Original" branch, :71
    mask = clip_to_tensor(mask_u8, channel_nb=3)
    return data, torch.unsqueeze(mask[0], dim=0)


# ------------------------------------------------------------------------------------------------
# Pillow's 8-bit bilinear resample restated in numpy (src/libImaging/Resample.c: precompute_coeffs,
# normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc / Vertical_8bpc). This is the arithmetic csrc/clip_io.cu
# implements; tests/test_clip_oracle.py pins it bit-exact against Image.resize itself.
# ------------------------------------------------------------------------------------------------
PRECISION_BITS = 32 - 8 - 2


def resample_coeffs(in_size, out_size):
    """-> (bounds int[out][2] = (first source index, count), kk int[out][ksize] 22-bit fixed-point weights)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale                       # bilinear filter support = 1
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)    # C's (int) truncates toward zero, like Python's int()
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = np.zeros(ksize)
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            w[x] = 1.0 - a if a < 1.0 else 0.0
        ww = 0.0
        for x in range(xmax):
            ww += w[x]
        if ww != 0.0:
            w[:xmax] = w[:xmax] / ww
        for x in range(ksize):
            v = w[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if w[x] < 0 else int(0.5 + v)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_axis(img, out_size, axis):
    img = np.moveaxis(np.asarray(img, dtype=np.int64), axis, 0)
    bounds, kk = resample_coeffs(img.shape[0], out_size)
    out = np.empty((out_size,) + img.shape[1:], dtype=np.int64)
    for xx in range(out_size):
        xmin, xmax = bounds[xx]
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(xmax):
            acc += img[xmin + x] * kk[xx, x]
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return np.moveaxis(out, 0, axis).astype(np.uint8)


def resample_u8_restated(frames, size):
    """frames uint8 (n, H, W, C) -> (n, size[0], size[1], C): horizontal pass (skipped when the width is unchanged),
    then vertical pass on the uint8 intermediate (skipped when the height is unchanged)."""
    frames = np.asarray(frames)
    if frames.shape[2] != size[1]:
        frames = _resample_axis(frames, int(size[1]), 2)
    if frames.shape[1] != size[0]:
        frames = _resample_axis(frames, int(size[0]), 1)
    return frames
