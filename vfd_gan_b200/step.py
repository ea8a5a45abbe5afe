"""The GAN training step (``MyGAN.optimize_params``, models/mygannet.py:275-366) on the B200 kernels.

``GanTrainStep.step`` performs exactly the parameter updates of the reference step:

    predict = netg(input)                                     forward_g   :275-276
    netd(gray2rgb(gt), gt_flow), netd(gray2rgb(predict.detach()), pre_flow)   forward_d :278-286
    err_g = w_adv * (l2(s_feat) + l2(t_feat)) + w_con * weighted_bce(predict, gt)
    optimizer_g.zero_grad(); err_g.backward(); optimizer_g.step()             :359-361
    err_d = ((bce(s_r,1)+bce(t_r,1))/2 + (bce(s_f,0)+bce(t_f,0))/2)/2
    optimizer_d.zero_grad(); err_d.backward(); optimizer_d.step()             :364-366

with two deliberate differences that do not change any result:
  * every discriminator input is detached in the reference, so the adversarial term has no path
    to NetG and its backward through NetD only deposits gradients that ``optimizer_d.zero_grad()``
    wipes; the term is evaluated for logging (fused squared-difference reduction) but not
    back-propagated;
  * the 12 logged scalars are written into one device tensor and read back once, instead of 12
    ``.item()`` synchronisations.

Optical flow (``video_to_flow``, host cv2 Farneback, lib/utils.py:94-129) is an *input* here, as in
SURVEY.md section 8d.

CUDA graph: after two eager steps the whole step (both forward passes, both backward passes, the weight
re-packing, the gradient all-reduces and the two fused-Adam updates -- about 830 kernel launches) is
captured once into a CUDA graph and replayed, which removes the Python / dispatcher launch overhead
that otherwise bounds the step. Dropout masks stay fresh under replay because the kernels add a device
resident step counter to their Philox seed. ``VFD_CUDA_GRAPH=0`` (or ``graph=False``) keeps the step eager.

Data parallelism (one process per GPU): ``GradAllReducer`` averages gradients over ranks with NCCL
in buckets, launched from post-accumulate-grad hooks on a side stream so the collectives overlap
the rest of backward. BatchNorm statistics stay per rank (DataParallel semantics).
"""
import os

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import _lib, ops, spatiotempconv
from .flow import video_to_flow

LOSS_KEYS = ("g/err_g", "g/err_g_adv", "g/err_g_adv_s", "g/err_g_adv_t", "g/err_g_con",
             "d/err_d_real_s", "d/err_d_real_t", "d/err_d_fake_s", "d/err_d_fake_t",
             "d/err_d_real", "d/err_d_fake", "d/err_d")


class GradAllReducer:
    """Persistent flat gradient buckets + bucketed, backward-overlapped averaging over the default process group.

    Every parameter's ``.grad`` is a view into a flat fp32 bucket buffer for the life of the trainer (static
    addresses: what a captured CUDA graph and the fused Adam need). Inside a fused step the kernels write the
    gradients straight into those views (``ops.StepContext``) and call ``ready(p)`` after a parameter's last
    contribution; parameters whose gradient still comes from autograd (the two Linear heads) reach ``ready`` through a
    post-accumulate-grad hook. With more than one rank a bucket is all-reduced (average) on ``comm_stream`` as soon
    as its last parameter is ready, overlapping the rest of backward; ``finish()`` joins the side stream. With one
    rank the same code runs without the collectives."""

    def __init__(self, params, bucket_mb=16.0, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        self.buckets, self.flat, self.pending, self.bucket_of = [], [], [], {}
        self.comm_stream, self.active = None, False
        cap = int(bucket_mb * 1024 * 1024 / 4)
        cur, cur_n = [], 0
        for p in reversed(self.params):  # backward produces gradients roughly in reverse order
            cur.append(p)
            cur_n += p.numel()
            if cur_n >= cap:
                self.buckets.append(cur)
                cur, cur_n = [], 0
        if cur:
            self.buckets.append(cur)
        for bi, bucket in enumerate(self.buckets):
            n = sum(p.numel() for p in bucket)
            flat = torch.zeros(n, dtype=torch.float32, device=bucket[0].device)
            off = 0
            for p in bucket:
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
                self.bucket_of[p] = bi
            self.flat.append(flat)
            self.pending.append(0)
        cuda = bool(self.params) and self.params[0].is_cuda
        if self.world > 1 and cuda:
            self.comm_stream = torch.cuda.Stream(device=self.params[0].device)
        self.avg = dist.ReduceOp.AVG if (self.world > 1 and cuda) else None   # NCCL averages in the collective
        self.by_ptr = {p.data_ptr(): p for p in self.params}
        for p in self.params:
            p.register_post_accumulate_grad_hook(self.ready)

    def zero(self):
        """Clears the buckets (one multi-tensor launch): gradients that autograd accumulates (``+=``) and the
        identically-zero conv-bias gradients rely on it; kernel-written gradients overwrite their slice anyway."""
        for p in self.params:       # a caller may have dropped the views (optimizer.zero_grad(set_to_none=True))
            if p.grad is None:
                bi = self.bucket_of[p]
                off = sum(q.numel() for q in self.buckets[bi][:self.buckets[bi].index(p)])
                p.grad = self.flat[bi][off:off + p.numel()].view_as(p)
        if self.flat:
            torch._foreach_zero_(self.flat)

    def begin(self):
        """Arm the ready-counting for one backward pass."""
        self.pending = [len(b) for b in self.buckets]
        self.active = True

    def ready(self, p):
        """Parameter ``p`` has received its whole gradient for this backward pass."""
        if not self.active:
            return
        bi = self.bucket_of[p]
        self.pending[bi] -= 1
        if self.pending[bi] == 0:
            self._launch(bi)

    def ready_tensor(self, t):
        """``ready`` for a tensor that aliases one of the parameters (what the autograd functions hold)."""
        self.ready(self.by_ptr[t.data_ptr()])

    def _launch(self, bi):
        if self.world == 1:
            return
        flat = self.flat[bi]
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            if ops.STEP.forked:                       # weight gradients are produced on the side stream
                self.comm_stream.wait_stream(ops.STEP.wstream)
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(flat, op=self.avg, group=self.group)
        else:  # CPU / gloo (tests): no AVG in gloo
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.mul_(1.0 / self.world)

    def finish(self):
        """Reduce whatever did not report ready (unused parameters) and join the side stream."""
        if self.world > 1:
            for bi, left in enumerate(self.pending):
                if left > 0:
                    self._launch(bi)
                    self.pending[bi] = 0
            if self.comm_stream is not None:
                torch.cuda.current_stream().wait_stream(self.comm_stream)
        self.active = False


class GanTrainStep:
    """One ``optimize_params``-equivalent step; see the module docstring."""

    def __init__(self, netg, netd, lr=2e-5, beta1=0.5, w_adv=1, w_con=10, pos_weight=2, distributed=None,
                 bucket_mb=16.0, graph=None):
        self.netg, self.netd = netg, netd
        self.w_adv, self.w_con, self.pos_weight = w_adv, w_con, pos_weight
        if distributed is None:
            distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.red_g = GradAllReducer(list(netg.parameters()), bucket_mb)
        self.red_d = GradAllReducer(list(netd.parameters()), bucket_mb)
        dev = next(netg.parameters()).device
        fused = dev.type == "cuda"
        if graph is None:
            graph = fused and os.environ.get("VFD_CUDA_GRAPH", "1") != "0"
        self.use_graph = bool(graph) and fused
        # Same hyper-parameters as models/mygannet.py:270-273 (capturable: the step counters live on the device
        # so the update can be replayed from a CUDA graph)
        self.optimizer_g = torch.optim.Adam(netg.parameters(), lr=lr, betas=(beta1, 0.999), fused=fused,
                                            capturable=self.use_graph)
        self.optimizer_d = torch.optim.Adam(netd.parameters(), lr=lr, betas=(beta1, 0.999), fused=fused,
                                            capturable=self.use_graph)
        self._g_ptrs = set(self.red_g.by_ptr)
        self.losses = torch.zeros(len(LOSS_KEYS), dtype=torch.float32, device=dev)
        self.predict = None
        self.packer = ops.WeightPacker([netg, netd]) if fused else None
        self._graph, self._static_in, self._eager_steps = None, None, 0
        self.adopt_inputs = False
        self._step_counter = torch.zeros((), dtype=torch.int64, device=dev) if fused else None

    GRAPH_WARMUP_STEPS = 2

    def step(self, inp, gt, gt_flow=None, pre_flow=None, dropout_seeds=None):
        """inp (B,3,D,H,W) in [-1,1]; gt (B,1,D,H,W) in {0,1}; flows (B,3,D,H,W). Returns the device
        tensor of the 12 logged scalars in LOSS_KEYS order (no host synchronisation).

        A flow left as ``None`` is computed inside the step exactly where the reference computes it
        (``video_to_flow(gray2rgb(gt))`` / ``video_to_flow(gray2rgb(predict.detach()))``,
        models/mygannet.py:279-282) -- on the device (vfd_gan_b200.flow) instead of on the host."""
        args = (inp, gt, gt_flow, pre_flow)
        if not self.use_graph or dropout_seeds is not None:
            return self._step_impl(*args, dropout_seeds=dropout_seeds)
        if self._graph is None:
            if self._eager_steps < self.GRAPH_WARMUP_STEPS:
                self._eager_steps += 1
                return self._step_impl(*args, seed_dev=self._step_counter)
            self._capture(args)
        elif any((s is None) != (t is None) or (s is not None and s.shape != t.shape)
                 for s, t in zip(self._static_in, args)):
            return self._step_impl(*args, seed_dev=self._step_counter)   # other batch geometry: stay eager
        for s, t in zip(self._static_in, args):
            if s is not None and s.data_ptr() != t.data_ptr():
                s.copy_(t, non_blocking=True)
        self._graph.replay()
        _lib.LAUNCHES += self.graph_calls
        _lib.KERNEL_LAUNCHES += self.graph_kernels          # kernels of ours the replay just launched
        ops.invalidate_packed_weights()                     # the replay updated the fp32 masters
        return self.losses

    def _capture(self, args):
        """Record one step into a CUDA graph. The graph's static inputs are private copies (every call copies its
        tensors into them), unless the caller hands its buffers over with ``adopt_inputs`` (HostBatchStep does: it
        owns them), in which case calls with those very tensors skip the copy."""
        self._static_in = [None if t is None else
                           (t if (self.adopt_inputs and t.is_contiguous() and t.dtype == torch.float32)
                            else t.detach().clone().contiguous().float())
                           for t in args]
        self.netg.train()
        self.netd.train()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        c0, k0 = _lib.LAUNCHES, _lib.KERNEL_LAUNCHES
        with torch.cuda.graph(graph):
            self._step_impl(*self._static_in, seed_dev=self._step_counter)
        self.graph_calls, self.graph_kernels = _lib.LAUNCHES - c0, _lib.KERNEL_LAUNCHES - k0
        _lib.LAUNCHES, _lib.KERNEL_LAUNCHES = c0, k0        # capture records, it does not launch
        self._graph = graph

    def _step_impl(self, inp, gt, gt_flow, pre_flow, dropout_seeds=None, seed_dev=None):
        netg, netd = self.netg, self.netd
        netg.train()
        netd.train()
        if seed_dev is not None:
            seed_dev += 1   # in-place on the device: captured, so every replay advances the dropout stream
        if self.packer is not None:
            self.packer.pack_all()          # every conv weight's bf16 GEMM operands, one launch
        spatiotempconv.DEFER_BN_COUNTERS = counters = []
        ops.ARENA.begin(inp.device)       # one zero fill for all weight-gradient accumulators of the step
        ops.STEP.begin(inp.device, self._grad_ready)
        try:
            return self._step_body(inp, gt, gt_flow, pre_flow, dropout_seeds, seed_dev)
        finally:
            ops.STEP.end()
            ops.ARENA.end()
            spatiotempconv.DEFER_BN_COUNTERS = None
            if counters:   # num_batches_tracked of every BatchNorm call of the step (NetD runs twice)
                seen = {}
                for c in counters:
                    ent = seen.setdefault(c.data_ptr(), [c, 0])
                    ent[1] += 1
                for n in sorted({e[1] for e in seen.values()}):
                    torch._foreach_add_([e[0] for e in seen.values() if e[1] == n], n)

    def _grad_ready(self, p):
        red = self.red_g if p.data_ptr() in self._g_ptrs else self.red_d
        red.ready(red.by_ptr[p.data_ptr()])

    def _step_body(self, inp, gt, gt_flow, pre_flow, dropout_seeds, seed_dev):
        netg, netd = self.netg, self.netd

        # forward_g
        logits, _ = netg.forward_cl(ops.PackFn.apply(inp, 0), dropout_seeds, seed_dev=seed_dev)
        predict = ops.SigmoidHeadFn.apply(logits)
        self.predict = predict.detach()

        # optical flow of the mask and of the prediction, when not supplied (models/mygannet.py:281-282)
        if gt_flow is None:
            gt_flow = video_to_flow(gt.expand(-1, 3, -1, -1, -1))
        if pre_flow is None:
            pre_flow = video_to_flow(self.predict.expand(-1, 3, -1, -1, -1))

        # forward_d: gray2rgb folded into the layout pack (1 -> 3 replicated channels)
        gt_cl = ops.PackFn.apply(gt, 3)
        pre_cl = ops.PackFn.apply(self.predict, 3)
        s_pr, s_fr, t_pr, t_fr = netd.forward_cl(gt_cl, ops.PackFn.apply(gt_flow, 0))
        s_pf, s_ff, t_pf, t_ff = netd.forward_cl(pre_cl, ops.PackFn.apply(pre_flow, 0))

        # backward_g
        with torch.no_grad():
            adv_s = ops.mse_cl(s_fr, s_ff, netd.spatdisc.feat_channels)
            adv_t = ops.mse_cl(t_fr, t_ff, netd.tempdisc.feat_channels)
        err_g_con = ops.WeightedBceFn.apply(predict, gt, float(self.pos_weight))
        self.red_g.zero()
        self.red_g.begin()
        (err_g_con * self.w_con).backward()
        ops.STEP.join()                 # the weight gradients of the side stream
        self.red_g.finish()
        self.optimizer_g.step()

        # backward_d
        ones, zeros = torch.ones_like(s_pr), torch.zeros_like(s_pf)
        e_rs, e_rt = F.binary_cross_entropy(s_pr, ones), F.binary_cross_entropy(t_pr, ones)
        e_fs, e_ft = F.binary_cross_entropy(s_pf, zeros), F.binary_cross_entropy(t_pf, zeros)
        err_d_real, err_d_fake = (e_rs + e_rt) * 0.5, (e_fs + e_ft) * 0.5
        err_d = (err_d_real + err_d_fake) * 0.5
        self.red_d.zero()
        self.red_d.begin()
        err_d.backward()
        ops.STEP.join()
        self.red_d.finish()
        self.optimizer_d.step()

        with torch.no_grad():
            adv = adv_s + adv_t
            con = err_g_con.detach()
            torch.stack([adv * self.w_adv + con * self.w_con, adv, adv_s, adv_t, con,
                         e_rs.detach(), e_rt.detach(), e_fs.detach(), e_ft.detach(),
                         err_d_real.detach(), err_d_fake.detach(), err_d.detach()], out=self.losses)
        return self.losses

    def losses_dict(self):
        """One device->host read of the 12 scalars (keys as logged at models/mygannet.py:314-342)."""
        vals = self.losses.tolist()
        return dict(zip(LOSS_KEYS, vals))


class HostBatchStep:
    """End-to-end entry: pinned host buffers in, loss scalars out.

    Every call copies the step's inputs host->device (async, from pinned memory, on a copy stream), runs
    ``GanTrainStep.step`` and reads the 12 loss scalars back -- the same boundary as
    ``lib/train_gan.py:69-70`` (``d.to('cuda')`` then ``optimize_params()``).

    The call is software-pipelined by one step, like a ``DataLoader(pin_memory=True)`` feeding
    ``.to('cuda', non_blocking=True)``: the copy of step i runs on the copy stream while step i-1 is still
    computing (two device input sets), and the scalars returned by call i are those of step i-1 (``None``
    for the first call; ``flush()`` returns the last step's). Nothing is skipped: every step's host->device
    bytes and device->host read are issued inside the caller's loop."""

    def __init__(self, trainer, batch, nfr, isize, device):
        self.trainer = trainer
        trainer.adopt_inputs = True     # self.dev below are the CUDA graph's static inputs (no extra copy per step)
        shp3, shp1 = (batch, 3, nfr, isize, isize), (batch, 1, nfr, isize, isize)
        mk = lambda: [torch.empty(shp3, device=device), torch.empty(shp1, device=device),
                      torch.empty(shp3, device=device), torch.empty(shp3, device=device)]
        self.dev = mk()            # the step's (CUDA-graph static) inputs
        self.stage = mk()          # landing buffers of the copy stream
        self.host_losses = [torch.empty(len(LOSS_KEYS), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.h2d_bytes = sum(t.numel() * 4 for t in self.dev)
        self.d2h_bytes = self.host_losses[0].numel() * 4
        cuda = torch.device(device).type == "cuda"
        self.copy_stream = torch.cuda.Stream(device=device) if cuda else None
        self.copied = torch.cuda.Event() if cuda else None
        self.consumed = torch.cuda.Event() if cuda else None
        self.read_back = [torch.cuda.Event() if cuda else None for _ in range(2)]
        self.n = 0

    def __call__(self, host_inp, host_gt, host_gt_flow, host_pre_flow):
        hosts = (host_inp, host_gt, host_gt_flow, host_pre_flow)
        cur = torch.cuda.current_stream()
        # host -> device on the copy stream; it may overlap the previous step's kernels
        with torch.cuda.stream(self.copy_stream):
            if self.n:
                self.copy_stream.wait_event(self.consumed)   # the landing buffers were drained
            for d, h in zip(self.stage, hosts):
                d.copy_(h, non_blocking=True)
            self.copied.record()
        cur.wait_event(self.copied)
        for d, s in zip(self.dev, self.stage):               # device-side hand-over into the static inputs
            d.copy_(s, non_blocking=True)
        self.consumed.record()
        losses = self.trainer.step(*self.dev)
        slot = self.n & 1
        self.host_losses[slot].copy_(losses, non_blocking=True)
        self.read_back[slot].record()
        self.n += 1
        if self.n == 1:
            return None
        self.read_back[slot ^ 1].synchronize()               # previous step's scalars are on the host
        return self.host_losses[slot ^ 1]

    def flush(self):
        """Scalars of the most recent step (waits for it)."""
        slot = (self.n - 1) & 1
        self.read_back[slot].synchronize()
        return self.host_losses[slot]
