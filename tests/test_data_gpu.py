"""GPU (-m gpu): the clip pipeline from decoded uint8 frames -- the device resize / ClipToTensor
kernels against the oracle (Pillow itself + the reference's float32 arithmetic; bit-exact), the pinned-ring
prefetcher, and the training-state checkpoint."""
import types

import numpy as np
import pytest
import torch

import vfd_gan_b200 as V
from oracle import clip_oracle as C
from oracle import vfd_oracle as O
from oracle import make_ref

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("h,w,size", [(120, 160, (112, 112)), (50, 70, (64, 64)), (112, 200, (112, 112)),
                                      (97, 97, (33, 41)), (360, 640, (112, 112)), (30, 40, (112, 112)),
                                      (64, 64, (64, 64)), (200, 112, (112, 112))])
def test_resize_frames_is_bit_exact_against_pillow(h, w, size):
    rng = np.random.default_rng(h * 1000 + w)
    for c in (3, 1):
        frames = rng.integers(0, 256, size=(5, h, w, c), dtype=np.uint8)
        got = V.data.resize_frames(torch.from_numpy(frames).to(DEV), size).cpu().numpy()
        assert np.array_equal(got, C.resize_frames(frames, size)), c


def test_frames_to_clip_is_bit_exact():
    rng = np.random.default_rng(3)
    frames = rng.integers(0, 256, size=(2, 4, 9, 11, 3), dtype=np.uint8)
    frames[0, 0, 0, :, 0] = np.arange(11) * 25                      # every residue class of x / 255
    masks = rng.integers(0, 2, size=(2, 4, 9, 11, 1), dtype=np.uint8) * 255
    got = V.data.frames_to_clip(torch.from_numpy(frames).to(DEV)).cpu()
    got01 = V.data.frames_to_clip(torch.from_numpy(frames).to(DEV), pm1=False).cpu()
    gotm = V.data.frames_to_clip(torch.from_numpy(masks).to(DEV), channels=3, pm1=False).cpu()
    for b in range(2):
        want = C.clip_to_tensor(frames[b])
        assert torch.equal(got01[b], want) and torch.equal(got[b], want * 2 - 1)
        assert torch.equal(gotm[b], C.clip_to_tensor(masks[b], channel_nb=3))
    every = torch.arange(256, dtype=torch.uint8).view(1, 1, 1, 256, 1)
    assert torch.equal(V.data.frames_to_clip(every.to(DEV)).cpu().view(-1), torch.arange(256).float().div(255) * 2 - 1)
    assert V.data.frames_to_clip(torch.empty((0, 4, 9, 11, 3), dtype=torch.uint8, device=DEV)).shape == (0, 3, 4, 9, 11)
    with pytest.raises(RuntimeError):
        V.data.frames_to_clip(torch.from_numpy(frames))            # CPU tensor: no fallback


def test_device_test_transform_against_the_reference_compose():
    """test.py:150-153 / lib/data.py:78 with the reference's own videotransforms (staged under oracle/_ref)."""
    if make_ref.ref_root() is None:
        pytest.skip("oracle/_ref not staged")
    import PIL.Image
    make_ref.import_ref()
    from videotransforms import video_transforms as vt, volume_transforms as vol
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, size=(2, 16, 90, 120, 3), dtype=np.uint8)
    masks = rng.integers(0, 2, size=(2, 16, 90, 120, 1), dtype=np.uint8) * 255
    tf = V.DeviceTestTransform(64)
    got = tf(torch.from_numpy(frames).to(DEV)).cpu()
    gotm = tf(torch.from_numpy(masks).to(DEV), mask=True).cpu()
    compose = vt.Compose([vt.Resize((64, 64)), vol.ClipToTensor()])
    for b in range(2):
        clip = [PIL.Image.fromarray(f) for f in frames[b]] + [PIL.Image.fromarray(m[..., 0]) for m in masks[b]]
        data, mask = torch.split(compose(clip), 16, dim=1)          # lib/data.py:64-66
        assert torch.equal(got[b], data * 2 - 1)                    # :78
        assert torch.equal(gotm[b], torch.unsqueeze(mask[0], dim=0))


@pytest.mark.parametrize("depth,pin", [(1, False), (2, True), (3, False)])
def test_clip_prefetcher_yields_the_loader_batches_in_order(depth, pin):
    g = torch.Generator().manual_seed(depth)
    batches = []
    for i in range(7):
        b = 4 if i < 6 else 3                                       # ragged last batch (no drop_last)
        t = (torch.rand(b, 3, 4, 16, 16, generator=g), torch.rand(b, 3, 4, 16, 16, generator=g),
             (torch.rand(b, 1, 4, 16, 16, generator=g) > 0.9).float(), torch.zeros(b, 4))
        batches.append(tuple(x.pin_memory() for x in t) if pin else t)
    pf = V.ClipPrefetcher(batches, DEV, depth=depth)
    big = torch.randn(2048, 2048, device=DEV)
    sums = []
    for dev_batch in pf:
        assert all(t.is_cuda for t in dev_batch)
        for _ in range(6):                                          # keep the consumer stream busy while copies run ahead
            big = torch.tanh(big @ big * 1e-3)
        sums.append([float(t.double().sum()) for t in dev_batch])
    assert len(sums) == 7
    for got, host in zip(sums, batches):
        for a, t in zip(got, host):
            assert abs(a - float(t.double().sum())) < 1e-6 * max(1.0, abs(a))
    assert pf.bytes_copied == sum(t.numel() * 4 for b in batches for t in b)
    # uint8 frames + device transform: 4x fewer bytes over the link, same clips as the host pipeline
    frames = [torch.randint(0, 256, (2, 4, 40, 50, 3), dtype=torch.uint8, generator=g) for _ in range(3)]
    pf8 = V.ClipPrefetcher(frames, DEV, depth=depth, transform=V.DeviceTestTransform(32))
    for clip, host in zip(pf8, frames):
        want = torch.stack([C.clip_to_tensor(C.resize_frames(h.numpy(), (32, 32))) * 2 - 1 for h in host])
        assert torch.equal(clip.cpu(), want)
    assert pf8.bytes_copied == sum(f.numel() for f in frames)


def test_training_state_resume_continues_the_trajectory(tmp_path):
    """3 steps + save + 2 steps against load + 2 steps, in deterministic mode: the same losses to 1e-6, with and without
    the captured graph (the load must drop a graph captured earlier). A weights-only resume (what the reference saves)
    restarts Adam's moments and must NOT give the same second step."""
    B, D, S = 2, 16, 64                                             # SDisc pools 64 -> 1
    args = types.SimpleNamespace(nfr=D, isize=S)

    def build(seed):
        torch.manual_seed(seed)
        netg, netd = V.NetG(), V.NetD(args)
        netg.apply(V.weights_init)
        netd.apply(V.weights_init)
        netg.dropout.p = 0.0
        return netg.to(DEV), netd.to(DEV)

    batches = [tuple(t.to(DEV) for t in O.synthetic_batch(B, D, S, seed=i)) for i in range(5)]
    from vfd_gan_b200 import ops
    ops.set_deterministic(True)      # bit-reproducible steps: the comparison below is exact, not a noise envelope
    try:
        _resume_cases(tmp_path, build, batches)
    finally:
        ops.set_deterministic(False)


def _resume_cases(tmp_path, build, batches):
    for graph in (False, True):
        netg, netd = build(0)
        tr = V.GanTrainStep(netg, netd, graph=graph, lr=2e-3)       # a large lr makes the optimizer state matter
        for i in range(3):
            tr.step(*batches[i])
        path = V.checkpoint.save_training_state(str(tmp_path / f"state_{graph}.pt"), tr, epoch=7, extra={"it": 3})
        V.checkpoint.save_weights(str(tmp_path / "w"), "mygan", 7, netg, netd)
        want = []
        for i in (3, 4):
            tr.step(*batches[i])
            want.append(tr.losses_dict())
        netg2, netd2 = build(1)                                     # different init: everything must come from the file
        tr2 = V.GanTrainStep(netg2, netd2, graph=graph, lr=2e-3)
        if graph:
            for i in range(3):                                      # capture a graph first: the load must drop it
                tr2.step(*batches[i])
        epoch, extra = V.checkpoint.load_training_state(path, tr2)
        assert epoch == 7 and extra == {"it": 3}
        for i, w in zip((3, 4), want):
            tr2.step(*batches[i])
            got = tr2.losses_dict()
            for k in w:
                assert abs(got[k] - w[k]) <= 1e-6 * abs(w[k]) + 1e-9, (graph, i, k, got[k], w[k])
        netg3, netd3 = build(2)
        tr3 = V.GanTrainStep(netg3, netd3, graph=False, lr=2e-3)
        V.checkpoint.load_weights(str(tmp_path / "w" / "mygan_ep0007_netG.pth"), netg3, netd3)
        tr3.step(*batches[3])
        first = tr3.losses_dict()
        assert abs(first["g/err_g_con"] - want[0]["g/err_g_con"]) <= 2e-3 * abs(want[0]["g/err_g_con"])  # same weights
        tr3.step(*batches[4])
        second = tr3.losses_dict()
        assert abs(second["g/err_g_con"] - want[1]["g/err_g_con"]) > 1e-4 * abs(want[1]["g/err_g_con"])  # fresh Adam
