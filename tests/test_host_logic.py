"""CPU: host-side logic -- the C-ABI library loads and exports what include/vfd_b200.h declares, the
tiling / padding helpers, and the multi-rank gradient all-reduce (gloo, world_size 2)."""
import ctypes
import os
import re

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vfd_gan_b200 import _lib, ops
from vfd_gan_b200.spatiotempconv import intermed_channels

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    hdr = open(os.path.join(ROOT, "include", "vfd_b200.h")).read()
    declared = set(re.findall(r"VFD_API\s+[\w\s\*]+?\b(vfd_\w+)\s*\(", hdr))
    assert len(declared) >= 23
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert declared - {"vfd_last_error", "vfd_abi_version"} == set(_lib.SIGNATURES)
    lib.vfd_abi_version.restype = ctypes.c_int
    assert lib.vfd_abi_version() == 3
    # test-only kernels and debug switches live in the debug library, not in the product ABI
    for name in _lib.DEBUG_SIGNATURES:
        assert not hasattr(lib, name), name
    dbg = ctypes.CDLL(_lib.DEBUG_LIB_PATH)
    dhdr = open(os.path.join(ROOT, "include", "vfd_b200_debug.h")).read()
    assert set(re.findall(r"VFD_API\s+[\w\s\*]+?\b(vfd_\w+)\s*\(", dhdr)) == set(_lib.DEBUG_SIGNATURES)
    for name in list(_lib.DEBUG_SIGNATURES) + list(_lib.SIGNATURES):
        assert hasattr(dbg, name), name


def test_intermediate_channels_follow_the_reference_formula():
    assert [intermed_channels(i, o, (3, 3, 3)) for i, o in ((3, 32), (32, 64), (64, 128), (128, 256), (256, 512))] == \
        [21, 115, 230, 460, 921]
    assert [intermed_channels(i, o, (1, 3, 3)) for i, o in ((3, 32), (32, 64), (512, 1024))] == [14, 52, 837]
    assert [intermed_channels(i, o, (3, 1, 1)) for i, o in ((3, 32), (32, 64), (64, 128))] == [2, 27, 54]


def test_channel_block_choice():
    assert ops.pick_kc(8) == 16 and ops.pick_kc(24) == 32 and ops.pick_kc(96) == 32
    assert ops.pick_kc(64) == 64 and ops.pick_kc(512) == 64 and ops.pick_kc(664) == 64
    for c in range(8, 1100, 8):
        kc = ops.pick_kc(c)
        assert ops.round_up(c, kc) <= 1.1 * min(ops.round_up(c, k) for k in (16, 32, 64))


def test_channels_last_view_checks():
    t = torch.zeros(2, 3, 4, 5, 16, dtype=torch.bfloat16)
    assert ops._is_cl(t) and ops._ld(t) == 16
    assert ops._is_cl(t[..., 8:]) and ops._ld(t[..., 8:]) == 16
    assert not ops._is_cl(torch.zeros(2, 1, 1, 1, 16, dtype=torch.bfloat16).expand(2, 4, 1, 1, 16))
    assert not ops._is_cl(t[..., 4:12])            # slice start not 16-byte aligned
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops._check_cl(t, "x")


def test_ops_are_registered_in_torch_library():
    for name in ("conv3d_fwd", "conv3d_wgrad", "bn_act_fwd", "bn_act_bwd", "upsample2x_fwd", "weighted_bce",
                 "convlstm_cell_fwd"):
        assert hasattr(torch.ops.vfd_b200, name)
    with pytest.raises((NotImplementedError, RuntimeError)):   # no CPU kernel registered
        torch.ops.vfd_b200.sqdiff(torch.zeros(1, 1, 1, 1, 8, dtype=torch.bfloat16),
                                  torch.zeros(1, 1, 1, 1, 8, dtype=torch.bfloat16), torch.zeros((), dtype=torch.float64))


def _allreduce_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vfd_gan_b200.step import GradAllReducer
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(n)) for n in (5, 300, 7, 1000)]
    red = GradAllReducer(params, bucket_mb=0.001)     # tiny buckets -> several collectives
    assert len(red.buckets) > 1
    red.zero()
    red.begin()
    loss = sum(((p * (rank + 1)) ** 2).sum() for p in params[:3])   # params[3] unused: finish() must cover it
    loss.backward()
    red.finish()
    want = [2 * p.detach() * sum((r + 1) ** 2 for r in range(world)) / world for p in params[:3]]
    ok = all(torch.allclose(p.grad, w, rtol=1e-5) for p, w in zip(params[:3], want))
    ok = ok and float(params[3].grad.abs().max()) == 0.0
    out[rank] = ok
    dist.destroy_process_group()


def test_gradient_allreduce_two_ranks_gloo():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_allreduce_worker, args=(world, 29731, out), nprocs=world, join=True)
        assert all(out[r] for r in range(world))


def _gather_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vfd_gan_b200.composed import gather_scores
    local = torch.arange(3, dtype=torch.float32) + 10 * rank          # rank r scores clips [3r, 3r+3)
    got = gather_scores(local)
    out[rank] = got.tolist()
    dist.destroy_process_group()


def test_score_gather_two_ranks_gloo():
    """Config 5 shards the clips over ranks with no data-path collective; only the per-clip scores are
    all-gathered (rank-major) before the global min-max scaling."""
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_gather_worker, args=(world, 29741, out), nprocs=world, join=True)
        for r in range(world):
            assert out[r] == [0.0, 1.0, 2.0, 10.0, 11.0, 12.0]


def test_new_modules_fail_loudly_on_cpu():
    import vfd_gan_b200 as V
    with pytest.raises(RuntimeError):
        V.evaluate.threshold_open(torch.rand(1, 1, 4, 8, 8))
    with pytest.raises(RuntimeError):
        V.evaluate.roc_auc(torch.zeros(4), torch.rand(4))
    with pytest.raises(RuntimeError):
        V.EncDecEncG(3, 8)(torch.rand(1, 3, 16, 16, 16))
    with pytest.raises(RuntimeError):
        V.AutoEncoder()(torch.rand(1, 3, 16, 16, 16))
    with pytest.raises(RuntimeError):
        V.video_to_flow(torch.rand(1, 3, 4, 32, 32))


def test_step_arena_sizing_reuse_and_growth():
    """ops.StepArena (weight-gradient accumulators of one train step): the first step only measures, later steps
    hand out zeroed, 256-byte aligned slices of one buffer, and a grown arena keeps the old buffer alive (a captured
    CUDA graph may still point into it)."""
    from vfd_gan_b200 import ops
    a = ops.StepArena()
    dev = torch.device("cpu")
    assert a.take((2, 3), dev) is None                      # inactive outside a step
    a.begin(dev)
    assert a.take((3, 8, 32), dev) is None and a.take((1, 8, 32), dev) is None   # first step: measure only
    a.end()
    a.begin(dev)
    x, y = a.take((3, 8, 32), dev), a.take((1, 8, 32), dev)
    assert x.shape == (3, 8, 32) and y.shape == (1, 8, 32) and x.dtype == torch.float32
    assert x.data_ptr() % 256 == a.buf.data_ptr() % 256 and (y.data_ptr() - x.data_ptr()) % 256 == 0
    x.fill_(1.0)
    y.fill_(2.0)
    a.end()
    first = a.buf
    a.begin(dev)                                            # same demand: same buffer, zeroed again
    x2 = a.take((3, 8, 32), dev)
    assert a.buf is first and x2.data_ptr() == x.data_ptr() and float(x2.abs().max()) == 0.0
    a.take((1, 8, 32), dev)
    assert a.take((64, 64, 64), dev) is None                # over the capacity: falls back, remembers the demand
    a.end()
    a.begin(dev)
    assert a.buf is not first and first in a.retired and a.take((64, 64, 64), dev) is not None
    a.end()


def test_zero_grad_views_are_shared_only_inside_a_step():
    from vfd_gan_b200 import ops
    dev = torch.device("cpu")
    g1, g2 = ops.zero_grad(5, dev), ops.zero_grad(5, dev)
    assert g1.data_ptr() != g2.data_ptr()                   # outside a fused step: private tensors
    ops.ARENA.begin(dev)
    try:
        h1, h2 = ops.zero_grad(5, dev), ops.zero_grad(7, dev)
        assert h1.data_ptr() == h2.data_ptr() and float(h2.abs().max()) == 0.0
    finally:
        ops.ARENA.end()


def test_public_header_is_plain_c(tmp_path):
    """include/vfd_b200.h is the drop-in boundary: it must compile as C99 and as C++ with no torch / CUDA types."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "use_header.c"
    src.write_text('#include "vfd_b200.h"\nint main(void) { return vfd_abi_version() == 0; }\n')
    for cc, std in (("gcc", "-std=c99"), ("g++", "-std=c++17")):
        if shutil.which(cc) is None:
            pytest.skip(f"{cc} not installed")
        args = [cc, std, "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(root, "include")]
        if cc == "g++":
            args += ["-x", "c++"]
        r = subprocess.run(args + [str(src)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    text = open(os.path.join(root, "include", "vfd_b200.h")).read()
    assert "at::" not in text and "c10::" not in text and "#include <torch" not in text


def test_step_context_refuses_a_second_concurrent_step():
    """The per-step state (gradient sinks, arena) is process-wide because autograd runs backward nodes on its own
    threads; starting a second fused step while one is active must fail loudly, not interleave."""
    import torch
    from vfd_gan_b200 import ops
    ops.STEP.begin(torch.device("cpu"))
    try:
        with pytest.raises(RuntimeError, match="not re-entrant"):
            ops.STEP.begin(torch.device("cpu"))
    finally:
        ops.STEP.end()
    ops.STEP.begin(torch.device("cpu"))
    ops.STEP.end()
