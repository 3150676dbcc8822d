"""Per-launch table (time, DRAM bytes) from an ncu --csv launch list; optional second file for a side-by-side."""
import csv, sys
def load(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows, order = {}, []
    for r in csv.DictReader(lines):
        k = r['ID']
        if k not in rows:
            rows[k] = {'name': r['Kernel Name']}
            order.append(k)
        rows[k][r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
    return [rows[k] for k in order]
a = load(sys.argv[1])
b = load(sys.argv[2]) if len(sys.argv) > 2 else None
CONVS = ["c1.conv1", "c1.conv2", "pool1", "c2.conv1", "c2.conv2", "pool2", "c3.conv1", "c3.conv2", "pool3", "c4.conv1", "c4.conv2", "pool4", "c5.conv1", "c5.conv2",
         "upv6", "sc6", "c6.conv1", "c6.conv2", "upv7", "sc7", "c7.conv1", "c7.conv2", "upsc8", "c8.conv1", "c8.conv2", "upsc9", "c9.conv1", "c9.conv2"]
def label(rows):  # by kernel name: the FiLM vectors take one or two launches, the output conv is fused or its own kernel
    out, nconv, nfilm = [], 0, 0
    for r in rows:
        n = r['name']
        if 'film_kernel' in n:
            nfilm += 1
            out.append(f"film.{nfilm}")
        elif 'head_conv' in n:
            out.append("head")
        elif 'tail_conv' in n:
            out.append("tail")
        elif 'conv_tc_kernel' in n:
            out.append(CONVS[nconv] if nconv < len(CONVS) else '?')
            nconv += 1
        else:
            out.append(n.split('(')[0][-9:])
    return out
NAMES = label(a)
ta = tb = 0
for i, r in enumerate(a):
    t = r['gpu__time_duration.sum'] / 1e3
    by = r.get('dram__bytes_read.sum', 0) + r.get('dram__bytes_write.sum', 0)
    ta += t
    line = f"{i:2d} {NAMES[i] if i < len(NAMES) else '?':9s} {t:8.1f} us {by / 1e6:8.1f} MB"
    if b and i < len(b):
        t2 = b[i]['gpu__time_duration.sum'] / 1e3
        tb += t2
        line += f"   | {t2:8.1f} us  {100 * (t2 - t) / t:+5.1f}%"
    print(line)
print(f"total {ta:.1f}" + (f" | {tb:.1f}" if b else ""))
