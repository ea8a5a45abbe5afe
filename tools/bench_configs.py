"""Timing of the other BASELINE.json configs (parity-test cases, not the bench.py line) on one B200:
    config 3  GAN step with the ConvLSTM bottleneck, 32-frame 128x128 clips
    config 4  STCNN AutoEncoder supervised step, 16x3x112x112
    config 5  enc-dec-enc anomaly-scoring sweep over synthetic clips (+ a bounded CPU-oracle sample)
Prints one JSON line per config (CUDA events, warm-up first, inputs resident in HBM).
    python tools/bench_configs.py [3] [4] [5] [--batch B] [--clips N]"""
import argparse
import json
import os
import sys
import time
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vfd_gan_b200 as V
from vfd_gan_b200 import ops, _lib

ap = argparse.ArgumentParser()
ap.add_argument("configs", nargs="*", default=["3", "4", "5"])
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--clips", type=int, default=10000)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--cpu", action="store_true", help="also time a bounded CPU-oracle sample for config 5")
args = ap.parse_args()
dev = torch.device("cuda", 0)


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _lib.KERNEL_LAUNCHES
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, (_lib.KERNEL_LAUNCHES - l0) // steps


if "3" in args.configs:
    B, D, S = args.batch or 16, 32, 128
    torch.manual_seed(0)
    netg = V.NetGLstm(3, 32, isize=S)
    netd = V.NetD(types.SimpleNamespace(nfr=D, isize=S))
    netg.apply(V.weights_init)
    netd.apply(V.weights_init)
    step = V.GanTrainStep(netg.to(dev), netd.to(dev))
    g = torch.Generator().manual_seed(1)
    inp = (torch.rand(B, 3, D, S, S, generator=g) * 2 - 1).to(dev)
    gt = (torch.rand(B, 1, D, S, S, generator=g) > 0.9).float().to(dev)
    gf = (torch.rand(B, 3, D, S, S, generator=g) * 2 - 1).to(dev)
    pf = (torch.rand(B, 3, D, S, S, generator=g) * 2 - 1).to(dev)
    ms, launches = timed(lambda: step.step(inp, gt, gf, pf), args.steps, warmup=4)
    losses = step.losses_dict()
    print(json.dumps({"config": 3, "workload": f"GAN step, NetG + ConvLSTM bottleneck (T={D // 16}), {D}x3x{S}x{S} clips, batch {B}, 1 GPU",
                      "ms_per_step": ms, "clips_per_s": B / ms * 1e3, "gpu_launches_per_step": launches,
                      "cuda_graph": step._graph is not None, "finite": all(v == v for v in losses.values()),
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)
    del step, netg, netd, inp, gt, gf, pf
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()

if "4" in args.configs:
    B, D, S = args.batch or 32, 16, 112
    torch.manual_seed(0)
    m = V.AutoEncoder()
    m.apply(V.weights_init)
    m = m.to(dev)
    tr = V.StcnnTrainStep(m)
    g = torch.Generator().manual_seed(2)
    inp = (torch.rand(B, 3, D, S, S, generator=g) * 2 - 1).to(dev)
    gt = (torch.rand(B, 1, D, S, S, generator=g) > 0.9).float().to(dev)
    ms, launches = timed(lambda: tr.step(inp, gt), args.steps)
    macs = 112.3e9 * (S / 112) ** 2          # 112.3 GMAC / clip forward at 16x112x112
    print(json.dumps({"config": 4, "workload": f"STCNN AutoEncoder BCE step, {D}x3x{S}x{S} clips, batch {B}, 1 GPU (eager, no CUDA graph)",
                      "ms_per_step": ms, "clips_per_s": B / ms * 1e3, "gpu_launches_per_step": launches,
                      "conv_tflops": 2 * 3 * macs * B / (ms * 1e-3) / 1e12, "loss": float(tr.step(inp, gt)),
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)
    del tr, m, inp, gt
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()

if "5" in args.configs:
    B, D, S = args.batch or 32, 16, 112
    torch.manual_seed(0)
    model = V.EncDecEncG(3, 32)
    model.apply(V.weights_init)
    model = model.to(dev).train()          # the reference's test loops never call .eval()
    model.netg.dropout.p = 0.0
    nb = (args.clips + B - 1) // B
    g = torch.Generator().manual_seed(3)
    pool = [(torch.rand(B, 3, D, S, S, generator=g) * 2 - 1).to(dev) for _ in range(4)]   # resident synthetic clips
    scorer = V.AnomalyScorer(model)
    for i in range(3):
        scorer.score_batch(pool[i % 4])
    scorer.finish()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _lib.KERNEL_LAUNCHES
    e0.record()
    for i in range(nb):
        scorer.score_batch(pool[i % 4])
    scaled, raw = scorer.finish()
    labels = (torch.arange(raw.numel(), device=dev) % 7 == 0).float()
    area = V.evaluate.roc_auc(labels[:16384], scaled[:16384])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    line = {"config": 5, "workload": f"enc-dec-enc anomaly-score sweep, {nb * B} synthetic {D}x3x{S}x{S} clips, batch {B}, 1 GPU, "
                                     "BatchNorm on batch statistics like the reference's test loops; incl. min-max scaling and ROC area",
            "total_ms": ms, "clips_per_s": nb * B / ms * 1e3, "gpu_launches": _lib.KERNEL_LAUNCHES - l0,
            "score_min_max": [float(raw.min()), float(raw.max())], "auc_first_16384": float(area[0])}
    if args.cpu:
        from oracle import vfd_oracle as O
        torch.set_num_threads(os.cpu_count() or 1)
        sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        xb = pool[0][:2].cpu()
        with torch.no_grad():
            O.enc_dec_enc_forward(sd, xb, True, [1.0] * 4)
            t0 = time.perf_counter()
            _, li, lo = O.enc_dec_enc_forward(sd, xb, True, [1.0] * 4)
            O.anomaly_scores(li, lo)
            dt = time.perf_counter() - t0
        line["cpu_oracle"] = {"clips_per_s": 2 / dt, "cores": os.cpu_count(), "sample": "1 batch of 2 clips, fp32 torch CPU"}
    print(json.dumps(line), flush=True)
