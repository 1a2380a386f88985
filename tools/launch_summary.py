#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launch count, total time and share of the
LAST `--steps` bench steps (a step ends with coverage_stats_kernel).  usage: launch_summary.py launches.csv out.txt [--steps 3] [--title ...]"""
import argparse
import collections
import csv
import re

ap = argparse.ArgumentParser()
ap.add_argument("csv"); ap.add_argument("out"); ap.add_argument("--steps", type=int, default=3); ap.add_argument("--title", default="")
a = ap.parse_args()
rows = [r for r in csv.reader(open(a.csv)) if len(r) > 5]
hdr, data = rows[0], rows[1:]
ix = {h: i for i, h in enumerate(hdr)}
name = lambda r: re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("rsqd::<unnamed>::", "")
val = lambda r: float(r[ix["Metric Value"]].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ix["Metric Unit"]], 1e-3)
ours = [r for r in data if "rsqd" in r[ix["Kernel Name"]]]
# walk back over the last `steps` steps: a step = [... hot_path ... coverage_stats]
ends = [i for i, r in enumerate(ours) if name(r).startswith("coverage_stats_kernel")]
first = ends[-a.steps - 1] + 1 if len(ends) > a.steps else 0
sel = ours[first:ends[-1] + 1] if ends else ours
agg = collections.OrderedDict()
for r in sel:
    k = name(r)
    n, t = agg.get(k, (0, 0.0))
    agg[k] = (n + 1, t + val(r))
tot = sum(t for _, t in agg.values())
with open(a.out, "w") as f:
    f.write(a.title + "\n\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:42s} n={n:4d} total={t:10.1f} us  share={100 * t / tot:5.1f}%\n")
    f.write(f"{'all kernels of this library':42s} n={sum(n for n, _ in agg.values()):4d} total={tot:10.1f} us\n")
print(open(a.out).read())
