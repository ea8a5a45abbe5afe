"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` capture of
tools/gpu_one_step.py: the LAST step's kernels, DRAM bytes and time per kernel name, and the per-kind totals bench.py
reads from profiles/dram_traffic_per_launch.json.
    python tools/dram_summary.py raw.csv out_prefix conv_launches_fwd_dgrad conv_launches_wgrad"""
import collections
import csv
import json
import re
import sys

raw, out_prefix = sys.argv[1], sys.argv[2]
with open(raw) as f:
    lines = [l for l in f if not l.startswith("==")]
per_id = collections.OrderedDict()
for r in csv.DictReader(lines):
    key = r["ID"]
    ent = per_id.setdefault(key, {"name": r["Kernel Name"]})
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "")
    name = r["Metric Name"]
    if name == "gpu__time_duration.sum":
        ent["us"] = val / 1e3 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3)
    else:
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        ent[name] = val * mult
rows = list(per_id.values())
marks = [i for i, r in enumerate(rows) if "pack_weights_batched" in r["name"]]
step = rows[marks[-1]:]
short = lambda n: re.sub(r"<.*", "", re.sub(r"\(.*", "", n.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")).replace("void ", ""))
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for r in step:
    a = agg[short(r["name"])]
    a[0] += 1
    a[1] += r.get("us", 0.0)
    a[2] += r.get("dram__bytes_read.sum", 0.0)
    a[3] += r.get("dram__bytes_write.sum", 0.0)
with open(out_prefix + "_dram_per_kernel.csv", "w") as f:
    f.write("kernel,launches,total_ms,dram_read_MB,dram_write_MB,GBps\n")
    for k, (c, us, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k},{c},{us / 1e3:.3f},{rd / 1e6:.1f},{wr / 1e6:.1f},{(rd + wr) / max(us, 1e-9) / 1e3:.0f}\n")
kinds = {"conv_fwd+dgrad": ("conv_fwd_res_kernel", "conv_fwd_tc_kernel", "tiny_pointwise_kernel", "conv_narrow_fwd_kernel"),
         "conv_wgrad": ("conv_wgrad", "thin_wgrad_kernel", "tiny_wgrad_kernel", "tap_gather_kernel"),
         "bn_act_fwd": ("bn_act_fwd_kernel",), "bn_act_bwd": ("bn_act_bwd",)}
tot = {}
for kind, pats in kinds.items():
    sel = [v for k, v in agg.items() if any(p in k for p in pats)]
    tot[kind] = {"launches": sum(v[0] for v in sel), "ms": sum(v[1] for v in sel) / 1e3,
                 "dram_bytes": sum(v[2] + v[3] for v in sel)}
print(json.dumps(tot, indent=1))
json.dump(tot, open(out_prefix + "_dram_per_kind.json", "w"), indent=1)
