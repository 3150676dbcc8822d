"""ncu CSV (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum) of one pipeline step -> per-kernel table:
launches, time, DRAM bytes moved, achieved GB/s and fraction of the measured HBM peak (MEASURED_PEAKS.json hbm_gbs).
usage: python tools/hbm_kernels_report.py <ncu.csv> [peak_gbs]"""
import csv
import json
import os
import re
import sys
from collections import defaultdict

path = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else None
if peak is None:
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6543.1
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
per = {}
for r in csv.DictReader([l for l in open(path) if not l.startswith("==")]):
    d = per.setdefault(int(r["ID"]), {"name": re.sub(r"^void |<unnamed>::|\(anonymous namespace\)::", "", re.sub(r"\(.*", "", r["Kernel Name"]))})
    key = {"dram__bytes_read.sum": "rd", "dram__bytes_write.sum": "wr", "gpu__time_duration.sum": "us"}.get(r["Metric Name"])
    if key:
        d[key] = float(r["Metric Value"].replace(",", "")) * scale.get(r["Metric Unit"], 1.0)
agg = defaultdict(lambda: [0, 0.0, 0.0])
for d in per.values():
    a = agg[d["name"]]
    a[0] += 1
    a[1] += d.get("us", 0.0)
    a[2] += d.get("rd", 0.0) + d.get("wr", 0.0)
print(f"HBM peak used: {peak:.1f} GB/s (MEASURED_PEAKS.json hbm_gbs); DRAM bytes = dram__bytes_read.sum + dram__bytes_write.sum (cold cache per launch)")
print(f"{'kernel':44s} {'launches':>8s} {'total_us':>10s} {'MB moved':>10s} {'GB/s':>8s} {'of peak':>8s}")
for k in sorted(agg, key=lambda k: -agg[k][1]):
    n, us, by = agg[k]
    gbs = by / us / 1e3 if us else 0.0
    print(f"{k[:44]:44s} {n:8d} {us:10.1f} {by / 1e6:10.1f} {gbs:8.0f} {gbs / peak:8.2f}")
