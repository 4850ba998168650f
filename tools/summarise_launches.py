"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches / time / share of the LAST step.
usage: python tools/summarise_launches.py launches.csv [n_steps_in_log]"""
import csv, re, sys
from collections import defaultdict
path = sys.argv[1]
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
    rows.append((r["Kernel Name"], us))
# the LAST step = everything after the previous step's optimizer launches (the first step of a process also holds one-off
# launches -- optimizer-state zero fills, initial weight casts -- so cutting the list into equal parts would mis-attribute them)
opt = [i for i, (n, _) in enumerate(rows) if "FusedOptimizer" in n or "adamw" in n.lower()]
runs = []   # contiguous groups of optimizer launches = one optimizer step each
for i in opt:
    if runs and i - runs[-1][1] <= 2:
        runs[-1][1] = i
    else:
        runs.append([i, i])
if len(runs) >= 2:
    last = rows[runs[-2][1] + 1: runs[-1][1] + 1]
else:
    per = len(rows) // nsteps
    last = rows[-per:]
agg = defaultdict(lambda: [0, 0.0])
for name, us in last:
    m = re.match(r"(?:void )?([\w:]+)", name)
    k = m.group(1) if m else name
    if k.startswith("at::") or k.startswith("at_cuda"):
        k = "torch: " + k.split("<")[0]
    agg[k][0] += 1
    agg[k][1] += us
tot = sum(v[1] for v in agg.values())
print(f"launches in the step: {len(last)}; summed kernel time {tot / 1e3:.2f} ms\n")
print("| kernel | launches | us | share |\n|---|---:|---:|---:|")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {us:.0f} | {100 * us / tot:.1f}% |")
