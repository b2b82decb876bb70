"""Markdown table of the round's multi-GPU bench lines (profiles/r2_bench_*_{N}gpu*.json)."""
import glob
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = {}
for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r2_bench_*gpu*.json"))):
    m = re.match(r"r2_bench_(.+?)_(\d)gpu(.*)\.json", os.path.basename(f))
    if not m:
        continue
    name, n, tag = m.group(1) + m.group(3), int(m.group(2)), m.group(3)
    try:
        d = [json.loads(l) for l in open(f) if l.startswith("{")][-1]
    except Exception:
        continue
    if "roofline" not in d:
        continue
    rows.setdefault(name, {})[n] = d
print("| workload | GPUs | value pair-cells/s | ms/step | e2e pair-cells/s | ms/step | roofline.frac | contraction ms | share of step | vs 1 GPU (value / e2e) |")
print("|---|---|---|---|---|---|---|---|---|---|")
for name in sorted(rows):
    base = rows[name].get(1) or rows.get(name.replace("_gridshard", "").replace("_nccl", ""), {}).get(1)
    for n in sorted(rows[name]):
        d = rows[name][n]
        r = d["roofline"]
        sp = "" if not base else "%.2f / %.2f" % (d["value"] / base["value"], d["e2e"]["value"] / base["e2e"]["value"])
        print(f"| {name} | {n} | {d['value']:.3g} | {d['ms_per_step']:.2f} | {d['e2e']['value']:.3g} | {d['e2e']['ms_per_step']:.2f} | {r['frac']:.3f} | {r.get('gram_ms', float('nan')):.3f} | "
              f"{r.get('gram_share_of_step', float('nan')):.2f} | {sp} |")
