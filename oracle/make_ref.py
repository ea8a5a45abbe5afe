#!/usr/bin/env python
"""Test infrastructure: stage the UNMODIFIED reference (pure Python) under ``oracle/_ref`` and import it.

The reference has no build step: its hot path is five small ``.py`` files (``models/spatiotempconv.py``,
``models/mygannet.py``, ``models/convlstm.py``, ``lib/train_gan.py``, ``lib/utils.py``) plus what they import
(``lib/evaluate.py``, ``videotransforms/``, ``models/mystcnn.py`` ...). ``stage()`` copies those packages from
``/root/reference`` into ``oracle/_ref/`` byte for byte. ``oracle/_ref/`` is git-ignored (reference sources never
enter this repository's history) but NOT gpurun-ignored, so the copy travels to the GPU box, where

  * ``tests/test_dropin_gpu.py`` drives the reference's own ``MyGAN`` / ``GANBaseModel.train`` loop over
    ``vfd_gan_b200.compat.install()``, and
  * ``bench.py --impl reference`` times the reference's own modules on the host cores (``kind: "reference"``).

Only ``tests/``, ``__graft_entry__`` and ``bench.py``'s CPU arms may import this file; nothing under
``vfd_gan_b200/`` does.

    python oracle/make_ref.py            # stage (idempotent)
"""
import hashlib
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DST = os.path.join(HERE, "_ref")
PACKAGES = ("models", "lib", "videotransforms")
FILES = ("trainer.py", "test.py")


def _digest(root):
    h = hashlib.sha256()
    for pkg in PACKAGES:
        for dirpath, _dirs, names in sorted(os.walk(os.path.join(root, pkg))):
            for n in sorted(names):
                if n.endswith(".py"):
                    p = os.path.join(dirpath, n)
                    h.update(os.path.relpath(p, root).encode())
                    with open(p, "rb") as f:
                        h.update(f.read())
    return h.hexdigest()


def stage(force=False):
    """Copy the reference's Python packages into oracle/_ref (no-op when /root/reference is absent, e.g. on the
    GPU box, or when the staged copy is already identical). Returns the staged path or None."""
    if not os.path.isdir(REF_SRC):
        return REF_DST if os.path.isdir(os.path.join(REF_DST, "models")) else None
    if not force and os.path.isdir(os.path.join(REF_DST, "models")) and _digest(REF_DST) == _digest(REF_SRC):
        return REF_DST
    if os.path.isdir(REF_DST):
        shutil.rmtree(REF_DST)
    os.makedirs(REF_DST)
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc")
    for pkg in PACKAGES:
        shutil.copytree(os.path.join(REF_SRC, pkg), os.path.join(REF_DST, pkg), ignore=ignore)
    for f in FILES:
        shutil.copy2(os.path.join(REF_SRC, f), os.path.join(REF_DST, f))
    with open(os.path.join(REF_DST, "STAGED_FROM"), "w") as f:
        f.write(f"{REF_SRC} sha256(py sources) {_digest(REF_DST)}\n")
    return REF_DST


def ref_root():
    """Where the reference can be imported from: the read-only mount in the build container, else oracle/_ref."""
    if os.path.isdir(os.path.join(REF_SRC, "models")):
        return REF_SRC
    if os.path.isdir(os.path.join(REF_DST, "models")):
        return REF_DST
    return None


def import_ref(root=None):
    """Import the reference's modules (SURVEY.md appendix C recipe: matplotlib / skimage are absent from this
    image and only used by plotting / data augmentation, so they are stubbed). -> namespace of modules."""
    root = root or ref_root()
    if root is None:
        raise ImportError("the reference is neither at /root/reference nor staged under oracle/_ref "
                          "(run `python oracle/make_ref.py` in the build container)")
    class _Stub(types.ModuleType):
        """Plotting / augmentation packages absent from this image: every attribute is a no-op callable."""

        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return lambda *a, **k: None

    for n in ("matplotlib", "matplotlib.pyplot", "skimage", "skimage.transform"):
        try:
            __import__(n)
        except ImportError:
            sys.modules[n] = _Stub(n)                                # lib/evaluate.py:8-12,40-56 (plots only),
    mpl, ski = sys.modules["matplotlib"], sys.modules["skimage"]     # videotransforms/functional.py:6
    if isinstance(mpl, _Stub):
        mpl.pyplot = sys.modules["matplotlib.pyplot"]
    if isinstance(ski, _Stub):
        ski.transform = sys.modules["skimage.transform"]
    sys.dont_write_bytecode = True                                   # /root/reference is read-only
    if root not in sys.path:
        sys.path.insert(0, root)
    import lib.evaluate as ev
    import lib.train_gan as tg
    import lib.utils as lu
    import models.convlstm as cl
    import models.mygannet as mg
    import models.mystcnn as ms
    import models.spatiotempconv as stc
    return types.SimpleNamespace(root=root, mygannet=mg, spatiotempconv=stc, convlstm=cl, mystcnn=ms, utils=lu,
                                 train_gan=tg, evaluate=ev)


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
