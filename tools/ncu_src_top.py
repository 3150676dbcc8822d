"""Top stall-sample SASS lines of one kernel from an `ncu --page source --csv` export (first kernel in the file)."""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = [i for i, r in enumerate(rows) if 'Source' in r][0]
body = []
for r in rows[hi + 1:]:
    if len(r) < 6 or r[0] == 'Address':
        break
    body.append(r)
tot = sum(int(r[2] or 0) for r in body)
print("instructions", len(body), "samples", tot)
by = collections.Counter()
for r in body:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[1])
    by[m.group(2) if m else '?'] += int(r[2] or 0)
print("by opcode:", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in by.most_common(14)))
idx = sorted(range(len(body)), key=lambda i: -int(body[i][2] or 0))[:topn]
for i in sorted(idx):
    r = body[i]
    print(f"{i:5d} {r[1][:100]:100s} samples {r[2]:>6s} exec {r[5]}")
