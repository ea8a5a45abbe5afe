"""The supervised STCNN mask segmenter (BASELINE config 4) on the B200 kernels.

Drop-in for the reference's ``models/mystcnn.py:6-88``: ``C2plus1d_Block(in_ch, out_ch, k=5)`` with children
``conv``, ``spaceconv``, ``pointwise``, ``bn1``, ``bn2``, ``avgpool``, ``dropout``, ``upsamp``, ``relu``,
``conv_last`` and ``AutoEncoder()`` with ``down_sep1-4``, ``up_sep1-4``, ``conv_last``, ``sigmoid`` -- same
constructor arguments, ``state_dict`` and fp32 NCDHW forward contract. It reuses the conv / BatchNorm / pool /
upsample kernels of the GAN path (kernel shapes 1x3x3, 3x1x1, 1x1x1 and 3x3x3); the train step is
``StcnnTrainStep`` = ``VFD_STCNN.train``'s inner loop (lib/train_stcnn.py:100-109).
"""
import torch
import torch.nn as nn

from . import _lib, ops
from .mygannet import _draw_seed
from .spatiotempconv import bn_apply


class C2plus1d_Block(nn.Module):
    def __init__(self, in_ch, out_ch, k=5):
        super().__init__()
        if in_ch % 8 and in_ch > 8:
            raise NotImplementedError("C2plus1d_Block on B200 needs in_ch <= 8 or divisible by 8")
        if out_ch % 8:
            raise NotImplementedError("C2plus1d_Block on B200 needs out_ch divisible by 8 (channel concat)")
        self.conv = nn.Conv3d(in_ch, out_ch, 1, stride=1)
        self.spaceconv = nn.Conv3d(in_ch, in_ch, (1, 3, 3), stride=1, padding=(0, 1, 1), dilation=1, bias=False)
        self.pointwise = nn.Conv3d(in_ch, out_ch, (3, 1, 1), stride=1, padding=(1, 0, 0), dilation=1, bias=False)
        self.bn1 = nn.BatchNorm3d(in_ch)
        self.bn2 = nn.BatchNorm3d(out_ch)
        self.avgpool = nn.AvgPool3d(2)
        self.dropout = nn.Dropout(p=0.25)
        self.upsamp = nn.Upsample(scale_factor=2, mode='trilinear', align_corners=True)
        self.relu = nn.ReLU(inplace=True)
        self.conv_last = nn.Conv3d(out_ch + out_ch, out_ch, 3, stride=1, padding=1, dilation=1, bias=False)
        self.out_ch = out_ch
        self.last_dropout_seed = None

    def forward_cl(self, xc, down_samp=False, dropout_seed=None, seed_dev=None):
        """channels-last bf16 -> channels-last bf16 (models/mystcnn.py:26-50)."""
        inp = xc
        tr1 = self.bn1.training or self.bn1.running_mean is None
        tr2 = self.bn2.training or self.bn2.running_mean is None
        y = ops.ConvFn.apply(xc, self.spaceconv.weight, None, False, False, tr1)
        a, _ = bn_apply(self.bn1, y, 0.0, stats_ready=tr1 and ops.conv_fuses_stats(self.spaceconv.out_channels))
        y = ops.ConvFn.apply(a, self.pointwise.weight, None, False, False, tr2)
        tr2 = tr2 and ops.conv_fuses_stats(self.pointwise.out_channels)
        if down_samp:
            _, x = bn_apply(self.bn2, y, 0.0, pool=(2, 2, 2), want_full=False, want_pool=True, stats_ready=tr2)
            inp = ops.ConvFn.apply(inp, self.conv.weight, self.conv.bias, False, False)
            inp = ops.IdentityPoolFn.apply(inp, (2, 2, 2), 0.0, 0)
        else:
            x, _ = bn_apply(self.bn2, y, 0.0, stats_ready=tr2)
            x = ops.UpsampleFn.apply(x)
            p = self.dropout.p if self.dropout.training else 0.0
            if p > 0.0:
                seed = int(dropout_seed) if dropout_seed is not None else _draw_seed()
                self.last_dropout_seed = seed
                inp = ops.IdentityPoolFn.apply(inp, (1, 1, 1), float(p), seed, seed_dev)
            inp = ops.UpsampleFn.apply(inp)
            inp = ops.ConvFn.apply(inp, self.conv.weight, self.conv.bias, False, False)
        x = torch.cat([x, inp], dim=-1)
        return ops.ConvFn.apply(x, self.conv_last.weight, None, False, False)

    def forward(self, x, down_samp=False):
        return ops.UnpackFn.apply(self.forward_cl(ops.PackFn.apply(x, 0), down_samp), self.out_ch)


class AutoEncoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.down_sep1 = C2plus1d_Block(3, 64)
        self.down_sep2 = C2plus1d_Block(64, 128)
        self.down_sep3 = C2plus1d_Block(128, 256)
        self.down_sep4 = C2plus1d_Block(256, 512)

        self.up_sep1 = C2plus1d_Block(512, 256)
        self.up_sep2 = C2plus1d_Block(256 + 256, 256)
        self.up_sep3 = C2plus1d_Block(256 + 128, 128)
        self.up_sep4 = C2plus1d_Block(128 + 64, 64)

        self.conv_last = nn.Conv3d(64, 1, 3, stride=1, padding=1, bias=False)
        self.sigmoid = nn.Sigmoid()

    def forward_cl(self, xc, dropout_seeds=None, seed_dev=None):
        """channels-last bf16 clip -> fp32 conv_last logits (models/mystcnn.py:69-88). ``seed_dev``: device counter
        added to every dropout seed inside the kernels (a CUDA-graph-captured step draws fresh masks per replay)."""
        N, D, H, W, _ = xc.shape
        if D % 16 or H % 16 or W % 16:
            raise RuntimeError(f"AutoEncoder needs nfr and isize divisible by 16, got D={D} H={H} W={W}")
        sd = list(dropout_seeds) if dropout_seeds is not None else [None] * 4
        d1 = self.down_sep1.forward_cl(xc, True)
        d2 = self.down_sep2.forward_cl(d1, True)
        d3 = self.down_sep3.forward_cl(d2, True)
        d4 = self.down_sep4.forward_cl(d3, True)
        u1 = self.up_sep1.forward_cl(d4, False, sd[0], seed_dev)
        u2 = self.up_sep2.forward_cl(torch.cat([u1, d3], dim=-1), False, sd[1], seed_dev)
        u3 = self.up_sep3.forward_cl(torch.cat([u2, d2], dim=-1), False, sd[2], seed_dev)
        u4 = self.up_sep4.forward_cl(torch.cat([u3, d1], dim=-1), False, sd[3], seed_dev)
        return ops.ConvFn.apply(u4, self.conv_last.weight, None, True, False)

    def forward(self, x):
        return ops.SigmoidHeadFn.apply(self.forward_cl(ops.PackFn.apply(x, 0)))


class StcnnTrainStep:
    """``opt.zero_grad(); predict = model(input); err = BCELoss(predict, gt); err.backward(); opt.step()``
    (lib/train_stcnn.py:104-109) with the fused BCE reduction and Adam(lr, (beta1, 0.999)) of :91.

    Like ``GanTrainStep``: gradients are written by the kernels into persistent flat buckets, averaged over the ranks
    of the default process group (one process per GPU, batch sharded on dim 0, per-rank BatchNorm statistics) with
    NCCL overlapped with backward, and after two eager steps the whole step is captured in a CUDA graph and
    replayed (``graph=False`` or ``VFD_CUDA_GRAPH=0`` keeps it eager)."""

    GRAPH_WARMUP_STEPS = 2

    def __init__(self, model, lr=2e-5, beta1=0.5, bucket_mb=16.0, graph=None):
        import os
        from .step import GradAllReducer
        self.model = model
        dev = next(model.parameters()).device
        fused = dev.type == "cuda"
        if graph is None:
            graph = fused and os.environ.get("VFD_CUDA_GRAPH", "1") != "0"
        self.use_graph = bool(graph) and fused
        self.red = GradAllReducer(list(model.parameters()), bucket_mb)
        self.opt = torch.optim.Adam(model.parameters(), lr=lr, betas=(beta1, 0.999), fused=fused,
                                    capturable=self.use_graph)
        self.packer = ops.WeightPacker([model]) if fused else None
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self.predict = None
        self._graph, self._static_in, self._eager_steps = None, None, 0
        self._step_counter = torch.zeros((), dtype=torch.int64, device=dev) if fused else None

    def step(self, inp, gt, dropout_seeds=None):
        """inp (B,3,D,H,W), gt (B,1,D,H,W) on the device -> the BCE loss (device scalar, no synchronisation)."""
        if not self.use_graph or dropout_seeds is not None:
            return self._step_impl(inp, gt, dropout_seeds)
        if self._graph is None:
            if self._eager_steps < self.GRAPH_WARMUP_STEPS:
                self._eager_steps += 1
                return self._step_impl(inp, gt, None, self._step_counter)
            self._static_in = [inp.detach().clone().contiguous().float(), gt.detach().clone().contiguous().float()]
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            c0, k0 = _lib.LAUNCHES, _lib.KERNEL_LAUNCHES
            with torch.cuda.graph(graph):
                self._step_impl(*self._static_in, None, self._step_counter)
            self._graph_calls, self._graph_kernels = _lib.LAUNCHES - c0, _lib.KERNEL_LAUNCHES - k0
            _lib.LAUNCHES, _lib.KERNEL_LAUNCHES = c0, k0        # capture records, it does not launch
            self._graph = graph
        elif inp.shape != self._static_in[0].shape or gt.shape != self._static_in[1].shape:
            return self._step_impl(inp, gt, None, self._step_counter)
        self._static_in[0].copy_(inp, non_blocking=True)
        self._static_in[1].copy_(gt, non_blocking=True)
        self._graph.replay()
        _lib.LAUNCHES += self._graph_calls
        _lib.KERNEL_LAUNCHES += self._graph_kernels         # kernels of ours the replay just launched
        ops.invalidate_packed_weights()
        return self.loss

    def _step_impl(self, inp, gt, dropout_seeds=None, seed_dev=None):
        from . import spatiotempconv
        self.model.train()
        if seed_dev is not None:
            seed_dev += 1
        if self.packer is not None:
            self.packer.pack_all()
        spatiotempconv.DEFER_BN_COUNTERS = counters = []
        ops.ARENA.begin(inp.device)
        ops.STEP.begin(inp.device, self.red.ready_tensor)
        try:
            self.red.zero()
            self.red.begin()
            logits = self.model.forward_cl(ops.PackFn.apply(inp, 0), dropout_seeds, seed_dev)
            predict = ops.SigmoidHeadFn.apply(logits)
            err = ops.BceLossFn.apply(predict, gt)
            err.backward()
            ops.STEP.join()
            self.red.finish()
            self.opt.step()
            self.predict = predict.detach()
            self.loss.copy_(err.detach())
        finally:
            ops.STEP.end()
            ops.ARENA.end()
            spatiotempconv.DEFER_BN_COUNTERS = None
            if counters:
                torch._foreach_add_(list({c.data_ptr(): c for c in counters}.values()), 1)
        return self.loss
