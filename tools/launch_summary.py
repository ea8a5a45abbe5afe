"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: takes the LAST complete train step (from one
vfd::pack_weights_batched_kernel to the next / the end) and writes per-launch and per-kernel tables."""
import csv, sys, collections, re

raw, out_prefix = sys.argv[1], sys.argv[2]
rows = []
with open(raw) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    t = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    t_us = t / 1e3 if unit in ("ns", "nsecond") else (t if unit in ("us", "usecond") else t * 1e3)
    rows.append((name, t_us))
marks = [i for i, (n, _) in enumerate(rows) if "pack_weights_batched" in n]
print("steps found:", len(marks), "kernels:", len(rows))
step = rows[marks[-2]:marks[-1]] if len(marks) >= 2 else rows[marks[-1]:]
short = lambda n: re.sub(r"\(.*", "", n).replace("void ", "")[:110]
with open(out_prefix + "_launches_one_step.csv", "w") as f:
    f.write("idx,kernel,time_us\n")
    for i, (n, t) in enumerate(step):
        f.write(f"{i},{short(n).replace(',', ';')},{t:.2f}\n")
agg = collections.defaultdict(lambda: [0, 0.0])
for n, t in step:
    k = re.sub(r"<.*", "", short(n))
    agg[k][0] += 1
    agg[k][1] += t
tot = sum(v[1] for v in agg.values())
with open(out_prefix + "_step_kernel_summary.csv", "w") as f:
    f.write(f"# kernels in the step: {len(step)}; sum of kernel durations: {tot / 1e3:.2f} ms\n")
    f.write("kernel,launches,total_ms,share_pct\n")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k},{c},{t / 1e3:.3f},{100 * t / tot:.1f}\n")
print(open(out_prefix + "_step_kernel_summary.csv").read())
