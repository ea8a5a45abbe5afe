#!/usr/bin/env python
"""Benchmark of the vfd_gan ``mygan`` training step on B200 (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one ``optimize_params``-equivalent GAN training step (models/mygannet.py:350-366) on one
batch of synthetic clips. Workload at every N: BASELINE.json configs[1] -- 16x3x112x112 clips,
batch 32 PER GPU (weak scaling), bf16 tensor-core convs with fp32 accumulation / fp32 master weights.
Prints ONE JSON line on rank 0.

``--impl reference`` times the reference's own CPU implementation of the step on the host cores: its unmodified
modules, staged under oracle/_ref by oracle/make_ref.py (git-ignored, shipped to the GPU box); the oracle port
only if that copy is missing.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import types

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NFR, ISIZE, BATCH_PER_GPU = 16, 112, 32
WORKLOAD = f"GANomaly-3D (mygan NetG+NetD, R(2+1)D) train step, synthetic {NFR}x3x{ISIZE}x{ISIZE} clips, batch {BATCH_PER_GPU}/GPU"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


class KernelProfiler:
    """CUDA-event timing of individual kernel calls on the launching stream (ops.PROFILER hook)."""

    def __init__(self):
        self.records = []

    def run(self, kind, work, thunk, nbytes=0.0):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        thunk()
        b.record()
        self.records.append((kind, work, a, b, nbytes))

    def dump(self, path, steps):
        """Per-call table (one step's worth): kind, work (FLOP or bytes), milliseconds."""
        n = len(self.records) // steps
        rows = [{"i": i, "kind": k, "work": w, "ms": a.elapsed_time(b), "bytes": nb} for i, (k, w, a, b, nb) in
                enumerate(self.records[-n:])]
        with open(path, "w") as f:
            json.dump(rows, f)

    def conv_by_bound(self, steps, peak_tflops, peak_gbs):
        """Every conv launch is bounded either by the tensor pipe or by HBM, whichever of
        FLOPs / peak and algorithmic bytes / peak is larger (ridge = peak_tflops / peak_gbs FLOP per byte).
        -> achieved fraction of the applicable roofline for each class, and the attainable time of the lot."""
        cls = {"tensor": {"ms": 0.0, "flop": 0.0, "bytes": 0.0, "launches": 0, "t_min_ms": 0.0},
               "hbm": {"ms": 0.0, "flop": 0.0, "bytes": 0.0, "launches": 0, "t_min_ms": 0.0}}
        for kind, work, a, b, nb in self.records:
            if not kind.startswith("conv"):
                continue
            t_tc, t_mem = work / (peak_tflops * 1e12) * 1e3, nb / (peak_gbs * 1e9) * 1e3
            c = cls["tensor" if t_tc >= t_mem else "hbm"]
            c["ms"] += a.elapsed_time(b)
            c["flop"] += work
            c["bytes"] += nb
            c["launches"] += 1
            c["t_min_ms"] += max(t_tc, t_mem)
        out = {}
        for name, c in cls.items():
            if not c["launches"]:
                continue
            ach = c["flop"] / (c["ms"] * 1e-3) / 1e12 if name == "tensor" else c["bytes"] / (c["ms"] * 1e-3) / 1e9
            peak = peak_tflops if name == "tensor" else peak_gbs
            out[name + "_bound_layers"] = {"launches_per_step": c["launches"] / steps, "ms_per_step": c["ms"] / steps,
                                           "achieved": ach, "unit": "TFLOP/s" if name == "tensor" else "GB/s",
                                           "peak": peak, "frac": ach / peak}
        tot_ms = sum(c["ms"] for c in cls.values())
        out["attainable_ms_per_step"] = sum(c["t_min_ms"] for c in cls.values()) / steps
        out["frac_of_attainable"] = sum(c["t_min_ms"] for c in cls.values()) / tot_ms if tot_ms else None
        return out

    def summary(self, steps):
        out = {}
        for kind, work, a, b, _nb in self.records:
            d = out.setdefault(kind, {"ms": 0.0, "work": 0.0, "launches": 0})
            d["ms"] += a.elapsed_time(b)
            d["work"] += work
            d["launches"] += 1
        for d in out.values():
            d["ms_per_step"] = d["ms"] / steps
            d["launches_per_step"] = d["launches"] / steps
        return out


def conv_macs_per_clip(net, shapes):
    """Algorithmic conv MACs per clip (no channel / tile padding): sum over Conv3d of voxels*Cin*Cout*taps."""
    total = 0
    for name, vox in shapes:
        conv = dict(net.named_modules())[name]
        w = conv.weight
        total += vox * w.shape[0] * w.shape[1] * w[0, 0].numel()
    return total


def model_conv_shapes(nfr, s):
    g, sd, td = [], [], []
    vox = lambda lvl: (nfr >> lvl) * (s >> lvl) * (s >> lvl)
    for i in range(1, 6):
        g += [(f"dconv{i}.conv.spatial_conv", vox(i - 1)), (f"dconv{i}.conv.temporal_conv", vox(i - 1))]
    for i, lvl in ((5, 4), (4, 3), (3, 2), (2, 1), (1, 0)):
        g += [(f"uconv{i}.conv.spatial_conv", vox(lvl)), (f"uconv{i}.conv.temporal_conv", vox(lvl))]
    g.append(("conv_last", vox(0)))
    hs = s
    for i in range(1, 7):
        sd += [(f"spatdisc.dconv{i}.conv.spatial_conv", nfr * hs * hs), (f"spatdisc.dconv{i}.conv.temporal_conv", nfr * hs * hs)]
        hs //= 2
    d = nfr
    for i in range(1, 4):
        td += [(f"tempdisc.dconv{i}.conv.spatial_conv", d * s * s), (f"tempdisc.dconv{i}.conv.temporal_conv", d * s * s)]
        d //= 2
    return g, sd + td


class CpuReferenceStep:
    """The reference's own CPU implementation of the step: its unmodified ``NetG`` / ``NetD`` modules and losses
    (staged under oracle/_ref by oracle/make_ref.py) driven by ``optimize_params``' sequence
    (models/mygannet.py:350-366; ``MyGAN`` itself hard-codes 'cuda', so the loop is restated around the modules).
    112 is not a size the reference's NetD accepts: its two Linears and TDisc's global pool are
    re-created for the workload's geometry, every conv / BatchNorm is the reference's. Falls back to the oracle
    port (validated bit-exact against these modules) when oracle/_ref is not staged."""

    def __init__(self, batch):
        import torch.nn as nn
        from oracle import make_ref
        from oracle import vfd_oracle as O
        self.batch = O.synthetic_batch(batch, NFR, ISIZE, seed=0)
        torch.manual_seed(0)
        if make_ref.ref_root() is not None:
            R = make_ref.import_ref()
            mg, self.lu = R.mygannet, R.utils
            self.netg = mg.NetG()
            self.netd = mg.NetD(types.SimpleNamespace(nfr=NFR, isize=128))
            self.netd.tempdisc.gpool = nn.AvgPool3d((1, ISIZE, ISIZE), stride=1)
            self.netd.spatdisc.linear = nn.Linear(32 * 32 * (ISIZE // 64) ** 2, 1)
            self.netd.tempdisc.linear = nn.Linear(32 * 4 * (NFR // 8), 1)
            self.netg.apply(self.lu.weights_init)
            self.netd.apply(self.lu.weights_init)
            self.opt_d = torch.optim.Adam(self.netd.parameters(), lr=2e-5, betas=(0.5, 0.999))
            self.opt_g = torch.optim.Adam(self.netg.parameters(), lr=2e-5, betas=(0.5, 0.999))
            self.bce = nn.BCELoss()
            self.kind = "reference"
            self.what = "the reference's own modules (oracle/_ref), fp32, torch CPU"
        else:
            import vfd_gan_b200 as V
            netg, netd = V.NetG(), V.NetD(types.SimpleNamespace(nfr=NFR, isize=ISIZE))
            netg.apply(V.weights_init)
            netd.apply(V.weights_init)
            self.port = O.OracleTrainer(netg.state_dict(), netd.state_dict())
            self.kind = "port"
            self.what = "oracle port, fp32, torch CPU"

    def step(self):
        if self.kind == "port":
            self.port.step(*self.batch)
            return
        lu, netg, netd = self.lu, self.netg, self.netd
        inp, gt, gt_flow, pre_flow = self.batch
        netg.train(), netd.train()
        predict = netg(inp)                                                       # forward_g
        pre_3ch, gt_3ch = lu.gray2rgb(predict.detach()), lu.gray2rgb(gt.detach())  # forward_d (flows are inputs)
        s_pr, s_fr, t_pr, t_fr = netd(gt_3ch, gt_flow)
        s_pf, s_ff, t_pf, t_ff = netd(pre_3ch.detach(), pre_flow)
        self.opt_g.zero_grad()                                                    # backward_g
        err_g = (lu.l2_loss(s_fr, s_ff) + lu.l2_loss(t_fr, t_ff)) * 1 + lu.weighted_bce(predict, gt) * 10
        err_g.backward(retain_graph=True)
        self.opt_g.step()
        self.opt_d.zero_grad()                                                    # backward_d
        ones, zeros = torch.ones_like(s_pr), torch.zeros_like(s_pf)
        err_d = ((self.bce(s_pr, ones) + self.bce(t_pr, ones)) * 0.5 +
                 (self.bce(s_pf, zeros) + self.bce(t_pf, zeros)) * 0.5) * 0.5
        err_d.backward()
        self.opt_d.step()


def run_reference(args):
    """CPU arm: the reference's own step on the host cores, bounded sample of the workload."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample_batch = 2
    ref = CpuReferenceStep(sample_batch)
    for _ in range(args.warmup):
        ref.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.step()
    dt = (time.perf_counter() - t0) / args.steps
    val = sample_batch / dt
    sample = (f"{args.steps} steps of batch {sample_batch} (of the {BATCH_PER_GPU}-clip workload), "
              f"{NFR}x3x{ISIZE}x{ISIZE}, {ref.what}")
    line = {"impl": "reference", "metric": "train_clips_per_sec", "value": val, "unit": "clips/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": sample},
            "cpu_baseline": {"value": val, "unit": "clips/s", "cores": cores, "kind": ref.kind, "sample": sample},
            "e2e": {"value": val, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_sample():
    """Bounded CPU sample timed next to the GPU number on rank 0 (N = 1 only)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref = CpuReferenceStep(2)
    ref.step()
    t0 = time.perf_counter()
    n = 2
    for _ in range(n):
        ref.step()
    dt = (time.perf_counter() - t0) / n
    return {"value": 2 / dt, "unit": "clips/s", "cores": cores, "kind": ref.kind,
            "sample": f"{n} steps of batch 2 of the same {NFR}x3x{ISIZE}x{ISIZE} workload ({ref.what})"}


_JSON_FD = None


def claim_stdout():
    """Libraries (NCCL's version banner, ...) write to the process's stdout; the contract is ONE JSON line there.
    Point file descriptor 1 at stderr for the duration of the run and keep the real stdout for emit()."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def finish(world):
    """Leave without tearing NCCL down: the captured CUDA graph still references the communicator and
    destroy_process_group() can block on it. All ranks synchronise, flush and exit 0."""
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="clips per GPU (default: the headline config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-flow", action="store_true", help="skip the supplementary in-step optical-flow measurement")
    ap.add_argument("--dump-kernels", default="", help="write the per-call CUDA-event table of one step to this file")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            run_reference(args)
        return

    import torch.distributed as dist
    import vfd_gan_b200 as V
    from vfd_gan_b200 import _lib, ops

    claim_stdout()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: vfd_gan_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    torch.manual_seed(0)
    a = types.SimpleNamespace(nfr=NFR, isize=ISIZE)
    netg, netd = V.NetG(), V.NetD(a)
    netg.apply(V.weights_init)
    netd.apply(V.weights_init)
    netg, netd = netg.to(dev), netd.to(dev)
    trainer = V.GanTrainStep(netg, netd)
    host = V.HostBatchStep(trainer, B, NFR, ISIZE, dev)

    g = torch.Generator().manual_seed(1 + rank)
    shp3, shp1 = (B, 3, NFR, ISIZE, ISIZE), (B, 1, NFR, ISIZE, ISIZE)
    h_inp = (torch.rand(shp3, generator=g) * 2 - 1).pin_memory()
    h_gt = (torch.rand(shp1, generator=g) > 0.9).float().pin_memory()
    h_gf = (torch.rand(shp3, generator=g) * 2 - 1).pin_memory()
    h_pf = (torch.rand(shp3, generator=g) * 2 - 1).pin_memory()
    # the device-resident arm runs on HostBatchStep's own device buffers (pre-filled once), the e2e arm
    # refills the same buffers from pinned host memory every step
    d_inp, d_gt, d_gf, d_pf = host.dev
    for d, h in zip(host.dev, (h_inp, h_gt, h_gf, h_pf)):
        d.copy_(h)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident inputs ("value")
    for _ in range(max(args.warmup, 3)):
        trainer.step(d_inp, d_gt, d_gf, d_pf)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = _lib.KERNEL_LAUNCHES
    ms = timed(lambda: trainer.step(d_inp, d_gt, d_gf, d_pf), args.steps)
    launches = _lib.KERNEL_LAUNCHES - l0
    losses = trainer.losses_dict()
    # ---- host buffers in, losses out ("e2e")
    for _ in range(2):
        host(h_inp, h_gt, h_gf, h_pf)

    def e2e_step():
        host(h_inp, h_gt, h_gf, h_pf)   # returns the previous step's 12 scalars (host memory, waited for)

    ms_e2e = timed(e2e_step, args.steps)
    host.flush()
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel CUDA-event pass (same step, instrumented; not part of the numbers above)
    roofline, kernels = None, None
    peaks = load_peaks()
    if not args.no_profile:
        prof = KernelProfiler()
        ops.PROFILER = prof
        psteps = 2
        for _ in range(psteps):   # eager (the timed steps above replay a CUDA graph; events need real launches)
            trainer._step_impl(d_inp, d_gt, d_gf, d_pf, seed_dev=trainer._step_counter)
        torch.cuda.synchronize()
        ops.PROFILER = None
        kernels = prof.summary(psteps)
        if args.dump_kernels and rank == 0:
            prof.dump(args.dump_kernels, psteps)
        conv = {k: v for k, v in kernels.items() if k.startswith("conv")}
        dom = max(conv, key=lambda k: conv[k]["ms"])
        dd = conv[dom]
        achieved = dd["work"] / (dd["ms"] * 1e-3) / 1e12
        conv_ms = sum(v["ms_per_step"] for v in conv.values())
        conv_flops = sum(v["work"] for v in conv.values()) / psteps
        kname = {"conv_fwd": "conv_fwd_res_kernel / conv_fwd_tc_kernel (forward)",
                 "conv_dgrad": "conv_fwd_res_kernel / conv_fwd_tc_kernel (dgrad)",
                 "conv_wgrad": "conv_wgrad{2,3,_t}_kernel / thin_wgrad_kernel (+ tap_gather for folded thin layers)"}[dom]
        # DRAM bytes per launch of the dominant kernel kind from the committed ncu pass (profiles/), if any
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "dram_traffic_per_launch.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get(dom)
        roofline = {"bound": "tensor", "kernel": kname, "achieved": achieved, "peak": peaks["bf16_tflops"],
                    "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops"], "traffic": traffic,
                    "peak_source": peaks["source"], "launches_per_step": dd["launches_per_step"],
                    "avg_launch_ms": dd["ms"] / dd["launches"],
                    "note": "aggregate over all launches of the dominant conv kernel kind in one step: algorithmic "
                            "FLOPs / CUDA-event time; most of these launches are thin HBM-bound layers",
                    "all_conv": {"tflops": conv_flops / (conv_ms * 1e-3) / 1e12, "ms_per_step": conv_ms,
                                 "frac": conv_flops / (conv_ms * 1e-3) / 1e12 / peaks["bf16_tflops"]},
                    "conv_by_bound": prof.conv_by_bound(psteps, peaks["bf16_tflops"], peaks["hbm_gbs"])}
        bn = {k: v for k, v in kernels.items() if k in ("bn_act_fwd", "bn_act_bwd")}
        if bn:
            bn_bytes = sum(v["work"] for v in bn.values())
            bn_ms = sum(v["ms"] for v in bn.values())
            roofline["hbm_kernels"] = {"bound": "hbm", "kernel": "bn_act_fwd / bn_act_bwd_{reduce,apply}",
                                       "achieved": bn_bytes / (bn_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                                       "unit": "GB/s", "frac": bn_bytes / (bn_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                       "ms_per_step": bn_ms / psteps}
        for k, v in kernels.items():
            if k.startswith("bn"):
                v["gbs"] = v["work"] / (v["ms"] * 1e-3) / 1e9
                v["frac_hbm"] = v["gbs"] / peaks["hbm_gbs"]
            else:
                v["tflops"] = v["work"] / (v["ms"] * 1e-3) / 1e12
            del v["work"]

    # ---- supplementary: the step with both optical flows computed inside it, where the reference computes them
    # (models/mygannet.py:281-282), on the device; plus cv2's Farneback on the host cores for scale (N = 1 only)
    flow_extra = None
    if world == 1 and not args.no_flow:
        tflow = V.GanTrainStep(netg, netd)
        for _ in range(4):
            tflow.step(d_inp, d_gt)
        ms_flow = timed(lambda: tflow.step(d_inp, d_gt), args.steps)
        for _ in range(2):
            V.video_to_flow(d_inp)
        ms_v2f = timed(lambda: V.video_to_flow(d_inp), args.steps)
        flow_extra = {"value": B * args.steps / (ms_flow * 1e-3), "unit": "clips/s", "ms_per_step": ms_flow / args.steps,
                      "cuda_graph": bool(tflow._graph is not None),
                      "video_to_flow_ms_per_call": ms_v2f / args.steps,
                      "note": "gt_flow and pre_flow computed in-step by vfd_gan_b200.video_to_flow (2 calls per step)"}
        try:
            import cv2
            import numpy as np
            rng = np.random.default_rng(0)
            a, b_ = (cv2.GaussianBlur(rng.random((ISIZE, ISIZE)).astype(np.float32), (0, 0), 2) for _ in range(2))
            cv2.calcOpticalFlowFarneback(a, b_, None, 0.5, 3, 15, 3, 5, 1.2, 0)
            t0 = time.perf_counter()
            npairs = 20
            for _ in range(npairs):
                cv2.calcOpticalFlowFarneback(a, b_, None, 0.5, 3, 15, 3, 5, 1.2, 0)
            per_pair = (time.perf_counter() - t0) / npairs
            flow_extra["host_cv2_farneback_ms_per_step"] = per_pair * 1e3 * 2 * B * (NFR - 1)
            flow_extra["host_cv2_sample"] = f"{npairs} calls of cv2.calcOpticalFlowFarneback on {ISIZE}x{ISIZE} float32 frames, " \
                                            f"scaled to the 2 x {B} x {NFR - 1} pairs of one step (the reference's single-threaded loop)"
        except Exception as e:   # cv2 missing on the box: report only the device side
            flow_extra["host_cv2_farneback_ms_per_step"] = None
            flow_extra["host_cv2_sample"] = f"unavailable: {e}"

    if rank != 0:
        finish(world)
        return

    gshapes, dshapes = model_conv_shapes(NFR, ISIZE)
    g_macs, d_macs = conv_macs_per_clip(netg, gshapes), conv_macs_per_clip(netd, dshapes)
    flop_per_clip = 2.0 * (3 * g_macs + 6 * d_macs)
    clips = B * world
    value = clips * args.steps / (ms * 1e-3)
    e2e = clips * args.steps / (ms_e2e * 1e-3)
    line = {
        "metric": "train_clips_per_sec", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": clips, "nfr": NFR, "isize": ISIZE, "parallelism": f"dp{world}",
                   "l2": "per-step inputs (257 MB) and activations (GBs) exceed the 126 MB L2; no explicit flush",
                   "algorithmic_conv_gflop_per_clip": flop_per_clip / 1e9,
                   "conv_tflops_whole_step": flop_per_clip * clips * args.steps / (ms * 1e-3) / 1e12 / world,
                   "optical_flow": "precomputed input (reference computes it on the host, SURVEY 8d)",
                   "cuda_graph": bool(trainer._graph is not None),
                   "e2e_pipeline": "H2D of step i overlaps compute of step i-1; scalars read one step late"},
        "e2e": {"value": e2e, "unit": "clips/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": host.h2d_bytes, "d2h_bytes_per_step": host.d2h_bytes},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "kernels": kernels,
        "with_device_flow": flow_extra,
        "losses_last_step": {k: round(v, 6) for k, v in losses.items()},
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample()
    emit(line)
    finish(world)


if __name__ == "__main__":
    main()
