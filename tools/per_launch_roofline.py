"""bench.py --dump-kernels JSON -> the per-launch roofline table under profiles/ (one eager, CUDA-event-timed step):
    python tools/per_launch_roofline.py dump.json out.csv [peak_tflops peak_gbs]
Conv launches carry algorithmic FLOPs and bytes and are judged against whichever roofline bounds them; the HBM-bound
kinds (bn_*) carry algorithmic bytes only."""
import json
import os
import sys

dump, out = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tf, gbs = 1383.5, 6547.2
pk = os.path.join(root, "MEASURED_PEAKS.json")
if len(sys.argv) > 4:
    tf, gbs = float(sys.argv[3]), float(sys.argv[4])
elif os.path.exists(pk):
    d = json.load(open(pk))
    tf, gbs = d.get("bf16_tflops_sustained", d["bf16_tflops"]), d["hbm_gbs"]
rows = json.load(open(dump))
tot_ms = tot_att = 0.0
with open(out, "w") as f:
    f.write(f"# One eager, CUDA-event-timed train step (`python bench.py --dump-kernels ...`, B200): every conv / BatchNorm launch with\n"
            f"# its algorithmic work, the roofline that bounds it (dense bf16 sustained {tf} TFLOP/s, copy {gbs} GB/s from\n"
            f"# MEASURED_PEAKS.json) and the fraction of that bound achieved. Eager launches carry a few microseconds of launch\n"
            f"# overhead each; the headline number replays a CUDA graph.\n")
    f.write("idx,kind,gflop,alg_mbytes,ms,bound,attainable_ms,frac_of_bound,tflops,gbs\n")
    for r in rows:
        conv = r["kind"].startswith("conv")
        flop = r["work"] if conv else 0.0
        nbytes = r["bytes"] if conv else r["work"]
        t_tensor = flop / (tf * 1e12) * 1e3
        t_hbm = nbytes / (gbs * 1e9) * 1e3
        bound = "tensor" if t_tensor > t_hbm else "hbm"
        att = max(t_tensor, t_hbm)
        ms = r["ms"]
        tot_ms += ms
        tot_att += att
        f.write(f"{r['i']},{r['kind']},{flop / 1e9:.3f},{nbytes / 1e6:.1f},{ms:.4f},{bound},{att:.4f},{att / ms if ms else 0:.3f},"
                f"{flop / ms / 1e9 if ms else 0:.1f},{nbytes / ms / 1e6 if ms else 0:.0f}\n")
    f.write(f"# total {tot_ms:.2f} ms timed, {tot_att:.2f} ms attainable\n")
print(f"{len(rows)} launches, {tot_ms:.2f} ms timed, {tot_att:.2f} ms attainable -> {out}")
