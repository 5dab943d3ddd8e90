"""Sums an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name, from the last witness kernel on
(= one batch run of scripts/profile_batch.py).  usage: python scripts/launch_summary.py launches.csv"""
import csv, re, sys
from collections import defaultdict
rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(u, 1e-6)
    rows.append((re.sub(r"\(.*", "", r["Kernel Name"]), v))
start = max(i for i, (k, _) in enumerate(rows) if "witness" in k)
agg, cnt = defaultdict(float), defaultdict(int)
for k, v in rows[start:]:
    agg[k] += v; cnt[k] += 1
tot = sum(agg.values())
for k in sorted(agg, key=agg.get, reverse=True):
    print(f"{agg[k]:9.3f} ms {cnt[k]:6d}  {100*agg[k]/tot:5.1f}%  {k}")
print(f"{tot:9.3f} ms total over {len(rows)-start} launches")
