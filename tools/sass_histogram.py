"""Per-kernel SASS opcode counts of the in-tree library (cuobjdump -sass): the tcgen05 / TMEM / TMA instructions that prove
the Blackwell path (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UTCBAR = tcgen05.commit, SYNCS =
mbarrier), plus the top opcodes of every kernel.
usage: python tools/sass_histogram.py [path/to/libyond_b200.so]"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "yond_public_b200", "libyond_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEY = ("UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTCBAR", "SYNCS", "HMMA", "DADD", "DFMA", "F2F", "MUFU", "SHFL", "ATOMS", "RED", "STG", "LDG", "LDS", "STS", "LDL", "STL")
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "").replace("void ", ""))
        cur = counts.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1
print(f"{os.path.basename(lib)}: {len(counts)} kernels")
for name, c in counts.items():
    tot = sum(c.values())
    fam = collections.Counter()
    for op, n in c.items():
        base = op.split(".")[0]
        if op.startswith("UTCHMMA.2CTA"):
            fam["UTCHMMA.2CTA"] += n
        for k in KEY:
            if base == k:
                fam[k] += n
    keyed = ", ".join(f"{k} {fam[k]}" for k in KEY if fam[k])
    print(f"\n{name[:110]}\n  {tot} instructions; {keyed}")
