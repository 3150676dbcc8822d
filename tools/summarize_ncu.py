"""Summarises an `ncu --csv --metrics gpu__time_duration.sum` launch list: per-kernel count, total and share."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void |<unnamed>::|\(anonymous namespace\)::", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'share':>7s} {'avg_us':>10s}")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{k[:60]:60s} {cnt[k]:8d} {tot[k]:12.1f} {100 * tot[k] / total:6.1f}% {tot[k] / cnt[k]:10.2f}")
print(f"{'TOTAL':60s} {sum(cnt.values()):8d} {total:12.1f}")
