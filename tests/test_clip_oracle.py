"""CPU: the clip-pipeline oracle (oracle/clip_oracle.py) against the reference's own videotransforms classes, and the
checkpoint helpers (reference file format, the resume that lib/train_gan.py / models/mygannet.py intend)."""
import os
import types

import numpy as np
import pytest
import torch

import vfd_gan_b200 as V
from oracle import clip_oracle as C
from oracle import make_ref


def _ref_transforms():
    if make_ref.ref_root() is None:
        pytest.skip("the reference is neither mounted nor staged under oracle/_ref")
    make_ref.import_ref()
    from videotransforms import video_transforms, volume_transforms
    return video_transforms, volume_transforms


def test_clip_oracle_is_bit_exact_against_reference_transforms():
    import PIL.Image
    vt, vol = _ref_transforms()
    rng = np.random.default_rng(0)
    for (h, w, size) in [(120, 160, (112, 112)), (50, 70, (64, 64)), (112, 200, (112, 112)), (97, 97, (33, 41))]:
        frames = rng.integers(0, 256, size=(4, h, w, 3), dtype=np.uint8)
        masks = (rng.integers(0, 2, size=(4, h, w, 1), dtype=np.uint8) * 255)
        clip = [PIL.Image.fromarray(f) for f in frames] + [PIL.Image.fromarray(m[..., 0]) for m in masks]
        want = vt.Compose([vt.Resize(size), vol.ClipToTensor()])(clip)             # test.py:150-153
        got_rgb = C.clip_to_tensor(C.resize_frames(frames, size))
        got_mask = C.clip_to_tensor(C.resize_frames(masks, size), channel_nb=3)
        assert torch.equal(want[:, :4], got_rgb)
        assert torch.equal(want[:, 4:], got_mask)
        data, mask = C.mdf_item(C.resize_frames(frames, size), C.resize_frames(masks, size))
        assert torch.equal(data, want[:, :4] * 2 - 1) and torch.equal(mask, want[:1, 4:])
        want_mask = vt.Compose([vt.Resize(size), vol.ClipToTensor(channel_nb=1)])(clip[4:])   # lib/data.py:21-24
        assert torch.equal(want_mask, C.clip_to_tensor(C.resize_frames(masks, size), channel_nb=1))


def test_numpy_restatement_of_the_pillow_resample_is_bit_exact():
    rng = np.random.default_rng(1)
    for (h, w, size) in [(120, 160, (112, 112)), (50, 70, (64, 64)), (112, 200, (112, 112)), (97, 97, (33, 41)),
                         (360, 640, (112, 112)), (30, 40, (112, 112)), (64, 64, (64, 64))]:
        for c in (3, 1):
            f = rng.integers(0, 256, size=(2, h, w, c), dtype=np.uint8)
            assert np.array_equal(C.resize_frames(f, size), C.resample_u8_restated(f, size)), (h, w, size, c)


def test_weight_files_follow_the_reference_format_and_resume_works(tmp_path):
    torch.manual_seed(0)
    args = types.SimpleNamespace(nfr=16, isize=64)
    netg, netd = V.NetG(), V.NetD(args)
    g_path, d_path = V.checkpoint.save_weights(str(tmp_path / "weights"), "mygan", 3, netg, netd)
    assert os.path.basename(g_path) == "mygan_ep0003_netG.pth" and os.path.basename(d_path) == "mygan_ep0003_netD.pth"
    ck = torch.load(g_path)
    assert set(ck) == {"epoch", "state_dict"} and ck["epoch"] == 4                 # lib/train_gan.py:54
    assert list(ck["state_dict"]) == list(netg.state_dict())
    g2, d2 = V.NetG(), V.NetD(args)
    assert V.checkpoint.load_weights(g_path, g2, d2, map_location="cpu") == 4
    for a, b in ((netg, g2), (netd, d2)):
        for (k, x), (_, y) in zip(a.state_dict().items(), b.state_dict().items()):
            assert torch.equal(x, y), k
    # a file written under the name models/mygannet.py:249 derives (underscore lost) is found as well, and so is a
    # DataParallel checkpoint ('module.' prefix)
    os.replace(d_path, str(tmp_path / "weights" / "mygan_ep0003netD.pth"))
    torch.save({"epoch": 4, "state_dict": {"module." + k: v for k, v in netg.state_dict().items()}}, g_path)
    g3, d3 = V.NetG(), V.NetD(args)
    V.checkpoint.load_weights(g_path, g3, d3, map_location="cpu")
    assert torch.equal(g3.state_dict()["conv_last.weight"], netg.state_dict()["conv_last.weight"])
    os.remove(str(tmp_path / "weights" / "mygan_ep0003netD.pth"))
    with pytest.raises(IOError):
        V.checkpoint.load_weights(g_path, g3, d3, map_location="cpu")
