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

# BASELINE.json configs[1..4]. cfg2 is the headline (the metric is quoted on it); the others are selectable so that the
# driver / a reader can time them on 1..8 GPUs with the same contract (--workload cfg3|cfg4|cfg5).
WORKLOADS = {
    "cfg2": {"kind": "gan", "nfr": 16, "isize": 112, "batch": 32, "netg": "NetG", "metric": "train_clips_per_sec",
             "scaling": "weak",
             "desc": "GANomaly-3D (mygan NetG+NetD, R(2+1)D) train step, synthetic 16x3x112x112 clips, batch 32/GPU"},
    "cfg3": {"kind": "gan", "nfr": 32, "isize": 128, "batch": 16, "netg": "NetGLstm", "metric": "train_clips_per_sec",
             "scaling": "weak",
             "desc": "GANomaly-3D + ConvLSTM bottleneck (NetGLstm+NetD) train step, synthetic 32x3x128x128 clips, batch 16/GPU"},
    "cfg4": {"kind": "stcnn", "nfr": 16, "isize": 112, "batch": 32, "metric": "train_clips_per_sec", "scaling": "weak",
             "desc": "STCNN (mystcnn AutoEncoder) supervised BCE train step, synthetic 16x3x112x112 clips, batch 32/GPU "
                     "(256 across 8 GPUs)"},
    "cfg5": {"kind": "score", "nfr": 16, "isize": 112, "batch": 32, "clips": 10000, "metric": "scored_clips_per_sec",
             "scaling": "strong",
             "desc": "enc-dec-enc anomaly-scoring sweep (latent score, min-max scaling, ROC area), 10k synthetic "
                     "16x3x112x112 clips sharded over the GPUs"},
}
NFR, ISIZE, BATCH_PER_GPU = 16, 112, 32          # set from the selected workload in main()
WORKLOAD = WORKLOADS["cfg2"]["desc"]


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
    """The reference's own CPU implementation of one step of the selected workload, on the host cores.

    cfg2: the reference's unmodified ``NetG`` / ``NetD`` modules and losses (staged under oracle/_ref by
    oracle/make_ref.py) driven by ``optimize_params``' sequence (models/mygannet.py:350-366; ``MyGAN`` itself
    hard-codes 'cuda', so the loop is restated around the modules). 112 is not a size the reference's NetD accepts
    (SURVEY D4): its two Linears and TDisc's global pool are re-created for the workload's geometry, every conv /
    BatchNorm is the reference's. cfg4: the reference's ``models.mystcnn.AutoEncoder`` + BCELoss + Adam
    (lib/train_stcnn.py:100-109). cfg3 / cfg5 are builder-defined compositions of reference modules:
    the oracle port of the composition. Falls back to the oracle port when oracle/_ref is not staged."""

    def __init__(self, batch, wl=None):
        import torch.nn as nn
        from oracle import make_ref
        from oracle import vfd_oracle as O
        import vfd_gan_b200 as V
        wl = wl or WORKLOADS["cfg2"]
        self.wl, self.O = wl, O
        nfr, isize = wl["nfr"], wl["isize"]
        self.batch = O.synthetic_batch(batch, nfr, isize, seed=0)
        torch.manual_seed(0)
        have_ref = make_ref.ref_root() is not None
        self.kind, self.what = "port", "oracle port, fp32, torch CPU"
        if wl["kind"] == "gan" and wl["netg"] == "NetG" and have_ref:
            R = make_ref.import_ref()
            mg, self.lu = R.mygannet, R.utils
            self.netg = mg.NetG()
            self.netd = mg.NetD(types.SimpleNamespace(nfr=nfr, isize=128))
            self.netd.tempdisc.gpool = nn.AvgPool3d((1, isize, isize), stride=1)
            self.netd.spatdisc.linear = nn.Linear(32 * 32 * (isize // 64) ** 2, 1)
            self.netd.tempdisc.linear = nn.Linear(32 * 4 * (nfr // 8), 1)
            self.netg.apply(self.lu.weights_init)
            self.netd.apply(self.lu.weights_init)
            self.opt_d = torch.optim.Adam(self.netd.parameters(), lr=2e-5, betas=(0.5, 0.999))
            self.opt_g = torch.optim.Adam(self.netg.parameters(), lr=2e-5, betas=(0.5, 0.999))
            self.bce = nn.BCELoss()
            self.mode, self.kind = "gan_ref", "reference"
            self.what = "the reference's own modules (oracle/_ref), fp32, torch CPU"
        elif wl["kind"] == "gan":
            netg = getattr(V, wl["netg"])(3, 32, isize=isize) if wl["netg"] == "NetGLstm" else V.NetG()
            netd = V.NetD(types.SimpleNamespace(nfr=nfr, isize=isize))
            netg.apply(V.weights_init)
            netd.apply(V.weights_init)
            fn = O.netg_lstm_forward if wl["netg"] == "NetGLstm" else None
            self.port = O.OracleTrainer(netg.state_dict(), netd.state_dict(), netg_fn=fn)
            self.mode = "gan_port"
        elif wl["kind"] == "stcnn" and have_ref:
            R = make_ref.import_ref()
            self.model = R.mystcnn.AutoEncoder()
            self.model.apply(R.utils.weights_init)
            self.opt = torch.optim.Adam(self.model.parameters(), lr=2e-5, betas=(0.5, 0.999))
            self.bce = nn.BCELoss()
            self.mode, self.kind = "stcnn_ref", "reference"
            self.what = "the reference's own models.mystcnn.AutoEncoder (oracle/_ref), fp32, torch CPU"
        elif wl["kind"] == "stcnn":
            m = V.AutoEncoder()
            m.apply(V.weights_init)
            self.sd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and "running" not in k)
                       for k, v in m.state_dict().items()}
            self.opt = torch.optim.Adam([v for v in self.sd.values() if v.requires_grad], lr=2e-5, betas=(0.5, 0.999))
            self.mode = "stcnn_port"
        else:
            m = V.EncDecEncG(3, 32)
            m.apply(V.weights_init)
            self.sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
            self.mode = "score_port"

    def step(self):
        O = self.O
        inp, gt, gt_flow, pre_flow = self.batch
        if self.mode == "gan_port":
            self.port.step(*self.batch)
        elif self.mode == "stcnn_ref":
            self.model.train()
            self.opt.zero_grad()
            err = self.bce(self.model(inp), gt)
            err.backward()
            self.opt.step()
        elif self.mode == "stcnn_port":
            self.opt.zero_grad()
            err = torch.nn.functional.binary_cross_entropy(O.autoencoder_forward(self.sd, inp, True, [1.0] * 4), gt)
            err.backward()
            self.opt.step()
        elif self.mode == "score_port":
            with torch.no_grad():
                _, li, lo = O.enc_dec_enc_forward(self.sd, inp, True, [1.0] * 4)
                O.anomaly_scores(li, lo)
        else:
            lu, netg, netd = self.lu, self.netg, self.netd
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


def run_reference(args, wl):
    """CPU arm: the reference's own step on the host cores, bounded sample of the workload."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample_batch = 2
    ref = CpuReferenceStep(sample_batch, wl)
    for _ in range(args.warmup):
        ref.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.step()
    dt = (time.perf_counter() - t0) / args.steps
    val = sample_batch / dt
    sample = (f"{args.steps} steps of batch {sample_batch} (of the {wl['batch']}-clip-per-GPU workload), "
              f"{wl['nfr']}x3x{wl['isize']}x{wl['isize']}, {ref.what}")
    line = {"impl": "reference", "metric": wl["metric"], "value": val, "unit": "clips/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "sample": sample},
            "cpu_baseline": {"value": val, "unit": "clips/s", "cores": cores, "kind": ref.kind, "sample": sample},
            "e2e": {"value": val, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(wl):
    """Bounded CPU sample timed next to the GPU number on rank 0 (N = 1 only)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref = CpuReferenceStep(2, wl)
    ref.step()
    t0 = time.perf_counter()
    n = 2
    for _ in range(n):
        ref.step()
    dt = (time.perf_counter() - t0) / n
    return {"value": 2 / dt, "unit": "clips/s", "cores": cores, "kind": ref.kind,
            "sample": f"{n} steps of batch 2 of the same {wl['nfr']}x3x{wl['isize']}x{wl['isize']} workload ({ref.what})"}


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


def finish(world, objs=()):
    """Orderly exit. The captured CUDA graphs hold references into the NCCL communicator: drop them (and everything
    that owns them) and drain the device BEFORE destroying the process group, then tear NCCL down normally."""
    if world > 1:
        import gc
        import torch.distributed as dist
        torch.cuda.synchronize()
        dist.barrier()
        for o in objs:
            for attr in ("_graph", "trainer"):
                if hasattr(o, attr):
                    setattr(o, attr, None)
        del objs
        gc.collect()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        try:
            dist.destroy_process_group()
        except Exception as e:      # never turn a finished measurement into a failure
            print(f"destroy_process_group: {e}", file=sys.stderr)


# ------------------------------------------------------------------------------------------------ workloads
class GanWorkload:
    """cfg2 / cfg3: one optimize_params-equivalent GAN step per call (models/mygannet.py:350-366)."""

    def __init__(self, wl, B, dev, rank):
        import vfd_gan_b200 as V
        self.wl, self.B, self.V = wl, B, V
        nfr, isize = wl["nfr"], wl["isize"]
        torch.manual_seed(0)
        a = types.SimpleNamespace(nfr=nfr, isize=isize)
        netg = V.NetGLstm(3, 32, isize=isize) if wl["netg"] == "NetGLstm" else V.NetG()
        netd = V.NetD(a)
        netg.apply(V.weights_init)
        netd.apply(V.weights_init)
        self.netg, self.netd = netg.to(dev), netd.to(dev)
        self.trainer = V.GanTrainStep(self.netg, self.netd)
        self.host = V.HostBatchStep(self.trainer, B, nfr, isize, dev)
        g = torch.Generator().manual_seed(1 + rank)
        shp3, shp1 = (B, 3, nfr, isize, isize), (B, 1, nfr, isize, isize)
        self.h = [(torch.rand(shp3, generator=g) * 2 - 1).pin_memory(), (torch.rand(shp1, generator=g) > 0.9).float().pin_memory(),
                  (torch.rand(shp3, generator=g) * 2 - 1).pin_memory(), (torch.rand(shp3, generator=g) * 2 - 1).pin_memory()]
        # the device-resident arm runs on HostBatchStep's own device buffers (pre-filled once), the e2e arm refills
        # the same buffers from pinned host memory every step
        for d, h in zip(self.host.dev, self.h):
            d.copy_(h)
        self.h2d_bytes, self.d2h_bytes = self.host.h2d_bytes, self.host.d2h_bytes
        gshapes, dshapes = model_conv_shapes(nfr, isize)
        g_macs, d_macs = conv_macs_per_clip(self.netg, gshapes), conv_macs_per_clip(self.netd, dshapes)
        if wl["netg"] == "NetGLstm":   # gate conv of the bottleneck: T = nfr/16 steps over (isize/16)^2 pixels
            w = self.netg.clstm.cell_list[0].conv.weight
            g_macs += (nfr // 16) * (isize // 16) ** 2 * w.shape[0] * w.shape[1] * 9
        self.flop_per_clip = 2.0 * (3 * g_macs + 6 * d_macs)
        self.clips_per_step = B

    def resident_step(self):
        self.trainer.step(*self.host.dev)

    def e2e_step(self):
        self.host(*self.h)          # returns the previous step's 12 scalars (host memory, waited for)

    def e2e_flush(self):
        self.host.flush()

    def profile_step(self):
        self.trainer._step_impl(*self.host.dev, seed_dev=self.trainer._step_counter)

    def result(self):
        return {"losses_last_step": {k: round(v, 6) for k, v in self.trainer.losses_dict().items()}}

    def graphed(self):
        return self.trainer._graph is not None

    def owners(self):
        return (self.trainer, self.host)


class StcnnWorkload:
    """cfg4: one supervised BCE step of the STCNN AutoEncoder per call (lib/train_stcnn.py:100-109)."""

    def __init__(self, wl, B, dev, rank):
        import vfd_gan_b200 as V
        self.wl, self.B = wl, B
        nfr, isize = wl["nfr"], wl["isize"]
        torch.manual_seed(0)
        m = V.AutoEncoder()
        m.apply(V.weights_init)
        self.model = m.to(dev)
        self.trainer = V.StcnnTrainStep(self.model)
        g = torch.Generator().manual_seed(2 + rank)
        self.h = [(torch.rand(B, 3, nfr, isize, isize, generator=g) * 2 - 1).pin_memory(),
                  (torch.rand(B, 1, nfr, isize, isize, generator=g) > 0.9).float().pin_memory()]
        self.dev = [t.to(dev) for t in self.h]
        self.stage = [torch.empty_like(t) for t in self.dev]
        self.host_loss = torch.empty((), dtype=torch.float32).pin_memory()
        self.h2d_bytes, self.d2h_bytes = sum(t.numel() * 4 for t in self.h), 4
        self.flop_per_clip = 2.0 * 3 * 112.3e9 * (isize / 112) ** 2 * (nfr / 16)   # 112.3 GMAC / clip forward
        self.clips_per_step = B

    def resident_step(self):
        self.trainer.step(*self.dev)

    def e2e_step(self):
        for d, h in zip(self.stage, self.h):
            d.copy_(h, non_blocking=True)
        loss = self.trainer.step(*self.stage)
        self.host_loss.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the step's loss is on the host

    def e2e_flush(self):
        torch.cuda.synchronize()

    def profile_step(self):
        self.trainer._step_impl(*self.dev, None, self.trainer._step_counter)

    def result(self):
        return {"loss_last_step": round(float(self.trainer.loss), 6)}

    def graphed(self):
        return self.trainer._graph is not None

    def owners(self):
        return (self.trainer,)


class ScoreWorkload:
    """cfg5: the anomaly-scoring sweep. One "step" = one batch of clips through enc-dec-enc + the per-clip latent
    score; the sweep (clips / world per rank) ends with the score all-gather, the global min-max scaling and the
    ROC area. The timed region of `value` is steps x batch clips per rank."""

    def __init__(self, wl, B, dev, rank):
        import vfd_gan_b200 as V
        self.wl, self.B, self.V, self.dev_ = wl, B, V, dev
        nfr, isize = wl["nfr"], wl["isize"]
        torch.manual_seed(0)
        model = V.EncDecEncG(3, 32)
        model.apply(V.weights_init)
        self.model = model.to(dev).train()          # the reference's test loops never call .eval()
        self.model.netg.dropout.p = 0.0
        self.scorer = V.AnomalyScorer(self.model)
        g = torch.Generator().manual_seed(3 + rank)
        self.h = [(torch.rand(B, 3, nfr, isize, isize, generator=g) * 2 - 1).pin_memory() for _ in range(2)]
        self.pool = [t.to(dev) for t in self.h]
        self.stage = [torch.empty_like(t) for t in self.pool]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.h2d_bytes, self.d2h_bytes = self.h[0].numel() * 4, B * 4
        gshapes = model_conv_shapes(nfr, isize)[0]
        enc = conv_macs_per_clip(self.model.netg, gshapes[:10])      # dconv1..5: the second encoder has the same shapes
        gmacs = conv_macs_per_clip(self.model.netg, gshapes)
        self.flop_per_clip = 2.0 * (gmacs + enc)     # NetG forward + the second encoder
        self.clips_per_step = B
        self.i = 0
        self.host_scores = torch.empty(B, dtype=torch.float32).pin_memory()

    def resident_step(self):
        self.scorer.score_batch(self.pool[self.i & 1])
        self.i += 1
        if len(self.scorer.chunks) >= 64:
            self.scorer.chunks = self.scorer.chunks[-1:]

    def e2e_step(self):
        # H2D of this batch (pinned host -> device), score, per-clip scores back to the host
        k = self.i & 1
        self.stage[k].copy_(self.h[k], non_blocking=True)
        s = self.scorer.score_batch(self.stage[k])
        self.host_scores.copy_(s, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self.i += 1
        if len(self.scorer.chunks) >= 64:
            self.scorer.chunks = self.scorer.chunks[-1:]

    def e2e_flush(self):
        torch.cuda.synchronize()

    def profile_step(self):
        self.resident_step()

    def result(self):
        scaled, raw = self.scorer.finish()
        labels = (torch.arange(raw.numel(), device=raw.device) % 7 == 0).float()
        area = self.V.evaluate.roc_auc(labels, scaled)
        return {"scores_gathered": int(raw.numel()), "score_min_max": [float(raw.min()), float(raw.max())],
                "roc_area_synthetic_labels": float(area[0])}

    def graphed(self):
        return False

    def owners(self):
        return ()


def kind_of_kernels(kernels, peaks):
    """Per-kind achieved rate and the roofline that bounds the kind (tensor pipe for convs whose aggregate
    arithmetic intensity is above the ridge, HBM otherwise)."""
    out = {}
    for k, v in kernels.items():
        if v["ms"] <= 0:
            continue
        sec = v["ms"] * 1e-3
        if k.startswith("conv"):
            t_tc, t_mem = v["work"] / (peaks["bf16_tflops"] * 1e12), v["bytes"] / (peaks["hbm_gbs"] * 1e9)
            bound = "tensor" if t_tc >= t_mem else "hbm"
        else:
            bound = "hbm"
        if bound == "tensor":
            ach, peak, unit = v["work"] / sec / 1e12, peaks["bf16_tflops"], "TFLOP/s"
        else:
            ach, peak, unit = (v["bytes"] if k.startswith("conv") else v["work"]) / sec / 1e9, peaks["hbm_gbs"], "GB/s"
        out[k] = {"bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak}
    return out


def main():
    global NFR, ISIZE, BATCH_PER_GPU, WORKLOAD
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS),
                    help="BASELINE.json config to time (default cfg2 = the headline configs[1])")
    ap.add_argument("--batch", type=int, default=0, help="clips per GPU (default: the workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-flow", action="store_true", help="skip the supplementary in-step optical-flow measurement")
    ap.add_argument("--dump-kernels", default="", help="write the per-call CUDA-event table of one step to this file")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    NFR, ISIZE, BATCH_PER_GPU, WORKLOAD = wl["nfr"], wl["isize"], wl["batch"], wl["desc"]

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            run_reference(args, wl)
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

    B = args.batch or wl["batch"]
    steps = args.steps
    if wl["kind"] == "score" and wl["scaling"] == "strong":
        # strong scaling: the 10k clips are sharded; `steps` batches per rank cover (a bounded sample of) the shard
        per_rank = (wl["clips"] + world - 1) // world
        steps = min(max(args.steps, 1) * 8, (per_rank + B - 1) // B) if args.steps < 40 else (per_rank + B - 1) // B
    work = {"gan": GanWorkload, "stcnn": StcnnWorkload, "score": ScoreWorkload}[wl["kind"]](wl, B, dev, rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident inputs ("value")
    warm = max(args.warmup, 3)
    for _ in range(warm):
        work.resident_step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = _lib.KERNEL_LAUNCHES
    ms = timed(work.resident_step, steps)
    launches = _lib.KERNEL_LAUNCHES - l0
    # ---- host buffers in, result scalars out ("e2e")
    for _ in range(2):
        work.e2e_step()
    ms_e2e = timed(work.e2e_step, steps)
    work.e2e_flush()
    clocks = sampler.stop() if rank == 0 else None
    result = work.result()

    # ---- per-kernel CUDA-event pass (same step, instrumented and eager with the weight gradients back on the main
    # stream so that event pairs do not overlap; not part of the numbers above)
    roofline, kernels = None, None
    peaks = load_peaks()
    if not args.no_profile:
        prof = KernelProfiler()
        ops.PROFILER = prof
        async_wgrad, ops.STEP.async_wgrad = ops.STEP.async_wgrad, False
        psteps = 2
        for _ in range(psteps):
            work.profile_step()
        torch.cuda.synchronize()
        ops.PROFILER = None
        ops.STEP.async_wgrad = async_wgrad
        kernels = prof.summary(psteps)
        if args.dump_kernels and rank == 0:
            prof.dump(args.dump_kernels, psteps)
        for kind, v in kernels.items():          # algorithmic bytes per kind (convs carry FLOPs in `work`)
            v["bytes"] = sum(nb for k2, _w, _a, _b, nb in prof.records if k2 == kind)
        per_kind = kind_of_kernels(kernels, peaks)
        timed_kinds = {k: v for k, v in kernels.items() if k != "bn_stats"}
        dom = max(timed_kinds, key=lambda k: timed_kinds[k]["ms"])
        dd, dk = kernels[dom], per_kind[dom]
        kname = {"conv_fwd": "conv_fwd_res_kernel / conv_fwd_tc_kernel (forward)",
                 "conv_dgrad": "conv_fwd_res_kernel / conv_fwd_tc_kernel (dgrad)",
                 "conv_wgrad": "conv_wgrad{2,3,_t,_tc}_kernel / thin_wgrad_kernel (+ tap_gather for folded thin layers)",
                 "bn_act_fwd": "bn_act_fwd_kernel", "bn_act_bwd": "bn_act_bwd{8,}_{reduce,apply}_kernel"}.get(dom, dom)
        # DRAM bytes per launch of each kernel kind from the committed ncu pass (profiles/), if any
        traffic_tab = {}
        tpath = os.path.join(ROOT, "profiles", "dram_traffic_per_launch.json")
        if os.path.exists(tpath) and args.workload == "cfg2":
            with open(tpath) as f:
                traffic_tab = json.load(f)
        alg_per_launch = (dd["bytes"] if dk["bound"] == "hbm" and dom.startswith("conv") else
                          (dd["work"] if dk["bound"] == "hbm" else dd["bytes"])) / dd["launches"]
        conv = {k: v for k, v in kernels.items() if k.startswith("conv")}
        conv_ms = sum(v["ms_per_step"] for v in conv.values())
        conv_flops = sum(v["work"] for v in conv.values()) / psteps
        roofline = {"bound": dk["bound"], "kernel": kname, "kind": dom, "achieved": dk["achieved"], "peak": dk["peak"],
                    "unit": dk["unit"], "frac": dk["frac"], "traffic": traffic_tab.get(dom),
                    "algorithmic_bytes_per_launch": alg_per_launch,
                    "peak_source": peaks["source"], "launches_per_step": dd["launches_per_step"],
                    "avg_launch_ms": dd["ms"] / dd["launches"], "ms_per_step": dd["ms_per_step"],
                    "note": "dominant kernel KIND of one step by CUDA-event time (all kinds considered): aggregate "
                            "algorithmic work of its launches / their summed time; per_kind lists every kind",
                    "per_kind": {k: {**per_kind[k], "ms_per_step": kernels[k]["ms_per_step"],
                                     "traffic_over_algorithmic": (traffic_tab[k] * kernels[k]["launches"] /
                                                                  (kernels[k]["bytes"] if k.startswith("conv") else kernels[k]["work"]))
                                     if traffic_tab.get(k) else None}
                                 for k in per_kind}}
        if conv:
            roofline["all_conv"] = {"tflops": conv_flops / (conv_ms * 1e-3) / 1e12, "ms_per_step": conv_ms,
                                    "frac": conv_flops / (conv_ms * 1e-3) / 1e12 / peaks["bf16_tflops"]}
            roofline["conv_by_bound"] = prof.conv_by_bound(psteps, peaks["bf16_tflops"], peaks["hbm_gbs"])
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
            del v["work"], v["bytes"]

    # ---- supplementary (headline workload, N = 1): the step with both optical flows computed inside it, where the
    # reference computes them (models/mygannet.py:281-282), on the device; plus cv2's Farneback on the host cores
    flow_extra = None
    if world == 1 and not args.no_flow and args.workload == "cfg2":
        d_inp, d_gt = work.host.dev[0], work.host.dev[1]
        tflow = V.GanTrainStep(work.netg, work.netd)
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
        finish(world, work.owners())
        return

    clips = work.clips_per_step * world
    value = clips * steps / (ms * 1e-3)
    e2e = clips * steps / (ms_e2e * 1e-3)
    line = {
        "metric": wl["metric"], "value": value, "unit": "clips/s", "n_gpus": world, "steps": steps,
        "warmup": warm, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": wl["scaling"],
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": wl["desc"], "workload_id": args.workload, "global_batch": clips, "nfr": NFR, "isize": ISIZE,
                   "parallelism": f"dp{world}",
                   "l2": "per-step inputs and activations (hundreds of MB to GBs) exceed the 126 MB L2; no explicit flush",
                   "algorithmic_conv_gflop_per_clip": work.flop_per_clip / 1e9,
                   "conv_tflops_whole_step": work.flop_per_clip * clips * steps / (ms * 1e-3) / 1e12 / world,
                   "optical_flow": "precomputed input (reference computes it on the host, SURVEY 8d)" if wl["kind"] == "gan" else None,
                   "cuda_graph": work.graphed(),
                   "e2e_pipeline": "H2D of step i overlaps compute of step i-1; scalars read one step late" if wl["kind"] == "gan"
                   else "H2D from pinned memory, step, result scalars to pinned host memory, stream sync -- every step"},
        "e2e": {"value": e2e, "unit": "clips/s", "ms_per_step": ms_e2e / steps,
                "h2d_bytes_per_step": work.h2d_bytes, "d2h_bytes_per_step": work.d2h_bytes},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "kernels": kernels,
        "with_device_flow": flow_extra,
    }
    line.update(result)
    if wl["kind"] == "score":
        line["config"]["sweep"] = (f"{wl['clips']} clips / {world} ranks; timed: {steps} batches of {B} per rank "
                                   f"({steps * B * world} clips in total)")
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample(wl)
    emit(line)
    finish(world, work.owners())


if __name__ == "__main__":
    main()
