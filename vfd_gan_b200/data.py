"""Clip pipeline between the reference's DataLoader and the train / evaluation step (SURVEY.md section 8(f) row 4;
lib/data.py:14-161, lib/train_gan.py:66-70).

The reference decodes each clip with cv2 inside DataLoader worker processes, runs PIL transforms on the frames,
converts to float32 (``ClipToTensor``), collates, and the training loop moves the four float tensors to the GPU with a
blocking ``d.to('cuda')`` per batch. What is rebuilt here is everything after the decode:

* ``resize_frames`` / ``frames_to_clip`` / ``DeviceTestTransform``: the deterministic test transform
  (``Resize((isize, isize))`` + ``ClipToTensor``, test.py:150-153, lib/data.py:143-146) and the ``*2-1`` of
  ``MdfDataLoader.__getitem__`` (lib/data.py:78) on the device, from uint8 frames -- bit-exact against Pillow and
  against the reference's float32 arithmetic. Frames cross PCIe / NVLink-C2C as bytes (4x less than float32).
* ``ClipPrefetcher``: wraps any iterable of host batches (the reference's DataLoader as it is, or a loader of uint8
  frames) and yields device batches. Batches are staged through a ring of pinned host buffers and copied on a side
  stream while the previous step computes; the consumer stream waits on an event, never on the host.

The random training augmentations (rotation, crop, flip: lib/data.py:133-141) and the video decode stay where the
reference has them (host worker processes); there is no NVDEC binding in this image.
"""
import torch

from . import _lib, ops


def _need_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: vfd_gan_b200 has no CPU path")


def resize_frames(frames, size):
    """uint8 frames (n, H, W, C) on the GPU, C in {1, 3} -> uint8 (n, size[0], size[1], C): what
    ``video_transforms.Resize(size)`` does to a list of PIL images (``img.resize((w, h), PIL.Image.BILINEAR)``,
    videotransforms/functional.py:43-58), bit-exact."""
    _need_cuda(frames, "resize_frames")
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] not in (1, 3):
        raise RuntimeError(f"resize_frames expects uint8 (n, H, W, 1|3), got {frames.dtype} {tuple(frames.shape)}")
    frames = frames.contiguous()
    n, hin, win, c = frames.shape
    hout, wout = int(size[0]), int(size[1])
    out = torch.empty((n, hout, wout, c), dtype=torch.uint8, device=frames.device)
    need = int(_lib.lib().vfd_resize_frames_u8_workspace(n, hin, win, c, hout, wout))
    if need < 0:
        raise RuntimeError("resize_frames: bad geometry")
    ws = torch.empty(max(need, 256), dtype=torch.uint8, device=frames.device)
    ops.resize_frames_u8_op(frames, out, ws)
    return out


def frames_to_clip(frames, channels=None, pm1=True):
    """uint8 frames (B, T, H, W, C) on the GPU -> float32 (B, channels, T, H, W): ``ClipToTensor`` (x / 255 in
    float32, videotransforms/volume_transforms.py:17-58) followed, with ``pm1``, by the ``*2-1`` of lib/data.py:78.
    A single-channel clip is broadcast over ``channels`` (the mask frames that go through the 3-channel
    ``ClipToTensor`` together with the RGB frames, lib/data.py:62-66)."""
    _need_cuda(frames, "frames_to_clip")
    if frames.dtype != torch.uint8 or frames.dim() != 5:
        raise RuntimeError(f"frames_to_clip expects uint8 (B, T, H, W, C), got {frames.dtype} {tuple(frames.shape)}")
    frames = frames.contiguous()
    b, t, h, w, c = frames.shape
    cout = c if channels is None else int(channels)
    if cout != c and c != 1:
        raise RuntimeError(f"frames_to_clip: cannot map {c} channels onto {cout}")
    out = torch.empty((b, cout, t, h, w), dtype=torch.float32, device=frames.device)
    ops.frames_to_clip_op(frames, out, bool(pm1))
    return out


class DeviceTestTransform:
    """``Compose([Resize((isize, isize)), ClipToTensor()])`` + ``*2-1`` (test.py:150-153, lib/data.py:78,143-146) for a
    batch of decoded clips: uint8 (B, T, H, W, 3) -> float32 (B, 3, T, isize, isize) in [-1, 1]; a mask batch uint8
    (B, T, H, W, 1) -> float32 (B, 1, T, isize, isize) in [0, 1] (``mask_transforms`` of lib/data.py:21-24 and the
    missing ``*2-1`` on the mask at :78)."""

    def __init__(self, isize):
        self.isize = int(isize)

    def __call__(self, frames, mask=False):
        b, t = frames.shape[:2]
        small = resize_frames(frames.reshape((b * t,) + tuple(frames.shape[2:])), (self.isize, self.isize))
        return frames_to_clip(small.view((b, t) + tuple(small.shape[1:])), pm1=not mask)


class ClipPrefetcher:
    """Iterate device batches over an iterable of host batches.

        for input, real, gt, lb in ClipPrefetcher(dataloader['train'], 'cuda'):   # lib/train_gan.py:66-70
            ...                                                                  # tensors are already on the GPU

    Every host batch (a tensor or a tuple / list of tensors, e.g. the ``(input, real, gt, lb)`` of lib/data.py:78)
    is copied into one of ``depth + 1`` pinned staging slots (skipped when the tensors are already pinned, e.g.
    ``DataLoader(pin_memory=True)``) and from there to the device on a private copy stream, ``depth`` batches ahead of
    the consumer. The consumer's current stream waits on the copy's event; a slot's device buffers are reused only
    after the consumer stream has passed the point where the following batch was requested, so kernels still reading
    a batch never race with the next copy. ``transform`` (optional) maps the tuple of device tensors to what the step
    consumes (e.g. ``DeviceTestTransform`` on uint8 frames) and runs on the consumer stream. Without a transform the
    yielded tensors are the slot's own device buffers: they are overwritten ``depth + 1`` batches later (clone what
    must live longer).

    ``bytes_copied`` counts the host->device bytes issued so far."""

    def __init__(self, loader, device="cuda", depth=2, transform=None):
        self.loader, self.device, self.depth, self.transform = loader, torch.device(device), int(depth), transform
        if self.device.type != "cuda":
            raise RuntimeError("ClipPrefetcher: vfd_gan_b200 has no CPU path")
        if self.depth < 1:
            raise ValueError("ClipPrefetcher: depth must be >= 1")
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.bytes_copied = 0
        self._slots = [None] * (self.depth + 1)

    def __len__(self):
        return len(self.loader)

    def _slot_buffers(self, idx, tensors):
        slot = self._slots[idx]
        sig = [(tuple(t.shape), t.dtype) for t in tensors]
        if slot is None or slot["sig"] != sig:
            slot = {"sig": sig,
                    "pinned": [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in tensors],
                    "dev": [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in tensors],
                    "copied": torch.cuda.Event(), "released": None, "used": False}
            self._slots[idx] = slot
        return slot

    def _issue(self, idx, batch):
        single = torch.is_tensor(batch)
        tensors = [batch] if single else list(batch)
        slot = self._slot_buffers(idx, tensors)
        if slot["released"] is not None:
            self.copy_stream.wait_event(slot["released"])      # the consumer is done with this slot's device buffers
        with torch.cuda.stream(self.copy_stream):
            for t, pin, dev in zip(tensors, slot["pinned"], slot["dev"]):
                src = t
                if not t.is_pinned():
                    if slot["used"]:
                        slot["copied"].synchronize()   # the pinned buffer was the source of this slot's previous copy
                    pin.copy_(t)
                    src = pin
                dev.copy_(src, non_blocking=True)
                self.bytes_copied += t.numel() * t.element_size()
            slot["copied"].record(self.copy_stream)
        slot["used"] = True
        slot["sources"] = tensors          # already-pinned sources must outlive their asynchronous copy
        return slot, single

    def __iter__(self):
        it = iter(self.loader)
        queue = []
        n = 0
        exhausted = False
        prev = None
        while True:
            while not exhausted and len(queue) < self.depth:
                try:
                    batch = next(it)
                except StopIteration:
                    exhausted = True
                    break
                queue.append(self._issue(n % (self.depth + 1), batch))
                n += 1
            if prev is not None:
                # the consumer asked for the next batch: everything it enqueued on its stream so far used `prev`
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self.device))
                prev["released"] = ev
                prev = None
            if not queue:
                return
            slot, single = queue.pop(0)
            torch.cuda.current_stream(self.device).wait_event(slot["copied"])
            out = slot["dev"]
            if self.transform is not None:
                out = self.transform(*out)
                single = torch.is_tensor(out)
                out = [out] if single else out
            prev = slot
            yield out[0] if single else tuple(out)
