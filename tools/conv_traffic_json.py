"""ncu CSV (metrics dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum; conv_tc kernels of one network
forward) -> profiles/*_conv_traffic.json, the per-launch DRAM traffic bench.py quotes in `roofline.traffic`.
usage: python tools/conv_traffic_json.py <ncu.csv> <blocks | "8 x 1536x2016 frames"> <out.json>"""
import csv
import json
import sys

path, blocks, out = sys.argv[1], sys.argv[2], sys.argv[3]
shape = f"{blocks} blocks of 128x128 packed px" if blocks.isdigit() else blocks
lines = [l for l in open(path) if not l.startswith("==")]
per = {}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
for r in csv.DictReader(lines):
    if "conv_tc" not in r["Kernel Name"]:
        continue
    d = per.setdefault(int(r["ID"]), {"kernel": "pair" if "<1>" in r["Kernel Name"] or "true" in r["Kernel Name"] else "single"})
    v = float(r["Metric Value"].replace(",", "")) * scale.get(r["Metric Unit"], 1.0)
    key = {"dram__bytes_read.sum": "read", "dram__bytes_write.sum": "write", "gpu__time_duration.sum": "us"}.get(r["Metric Name"])
    if key:
        d[key] = v
rows = [dict(id=i, **per[k]) for i, k in enumerate(sorted(per))]
tot = sum(r["read"] + r["write"] for r in rows)
res = {"what": f"conv_tc_kernel DRAM traffic, GuidedResUnet forward on {shape} (one network forward of bench.py), "
               "ncu dram__bytes_read+write per launch, cold cache (ncu flushes between launches)",
       "note": f"one network forward on {shape}",
       "launches": len(rows), "dram_bytes_total": tot, "dram_bytes_per_launch": tot / max(1, len(rows)),
       "time_us_total": sum(r["us"] for r in rows), "per_launch": rows}
json.dump(res, open(out, "w"), indent=1)
print(f"{len(rows)} launches, {tot / 1e6:.0f} MB total, {tot / max(1, len(rows)) / 1e6:.1f} MB/launch, {res['time_us_total']:.0f} us")
