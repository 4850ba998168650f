"""Opcode histogram of every kernel in libegom2p_b200.so (cuobjdump -sass): the evidence that the hot kernels are tcgen05 /
TMEM / TMA code (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA load / store /
reduce, UTCBAR = tcgen05.commit, SYNCS = mbarrier). usage: python tools/sass_histogram.py > profiles/rNN_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "egom2p_b200", "libegom2p_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "SYNCS", "MUFU", "HMMA", "FFMA", "RED", "ATOM")
cur, hist = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
print("kernel | total | " + " | ".join(KEY))
tot = collections.Counter()
for k, c in hist.items():
    row = [sum(v for op, v in c.items() if op.split(".")[0] == key or op.startswith(key)) for key in KEY]
    for key, v in zip(KEY, row):
        tot[key] += v
    print(f"{k[:90]} | {sum(c.values())} | " + " | ".join(map(str, row)))
print("ALL | - | " + " | ".join(str(tot[k]) for k in KEY))
print("\nfull opcode list of the library (count >= 20):")
allc = collections.Counter()
for c in hist.values():
    allc.update(c)
for op, v in allc.most_common():
    if v >= 20:
        print(f"  {op:40s} {v}")
