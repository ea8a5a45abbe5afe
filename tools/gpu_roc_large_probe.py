"""Times the voxel-level ROC / PR pipeline (csrc/roc_large.cu) at test-sweep sizes and prints one JSON line per size:
algorithmic bytes = 8 B read (score, label) per pair; traffic estimate = keys 16 B + 4 sort passes x (8 hist + 8 + 8
scatter) + prefix 8 + 8 + 4 + areas 12 = 156 B per pair."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import vfd_gan_b200 as V  # noqa: E402

for n in (1 << 20, 10_000_000, 100_000_000):
    g = torch.Generator(device="cuda").manual_seed(1)
    lab = (torch.rand(n, device="cuda", generator=g) > 0.95).float()
    sc = torch.sigmoid(torch.randn(n, device="cuda", generator=g) + 2.0 * lab)
    for quant in (False, True):
        s = (sc * 255).round() / 255 if quant else sc
        V.evaluate.roc_auc(lab, s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            out = V.evaluate.roc_auc(lab, s)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        line = {"n": n, "scores": "256 levels" if quant else "continuous", "ms": round(ms, 3),
                "pairs_per_s": n / ms * 1e3, "est_traffic_GBps": 156.0 * n / ms / 1e6, "roc": out[0].item(),
                "pr": out[3].item()}
        if n <= 10_000_000:
            from sklearn.metrics import roc_auc_score
            t0 = time.time()
            ref = roc_auc_score(lab.cpu().numpy(), s.cpu().numpy())
            line["sklearn_roc_s"] = round(time.time() - t0, 3)
            line["roc_abs_diff"] = abs(ref - line["roc"])
        print(json.dumps(line), flush=True)
