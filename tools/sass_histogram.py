#!/usr/bin/env python
"""SASS instruction histogram per kernel of liblineax_b200.so (no GPU needed):
python tools/sass_histogram.py > profiles/r02_sass_histograms.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "lineax_b200/liblineax_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda s: subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip()
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        hist[kern][m.group(2)] += 1
MARK = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "FFMA2", "CREDUX", "LDGSTS", "UTCBAR", "SYNCS")
print(f"# cuobjdump -sass {lib}: static SASS instruction counts per kernel (sm_100a)")
print("# Blackwell markers: UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG/UTMASTG = TMA tensor load/store,")
print("# UBLKCP = bulk copy, FFMA2 = fma.rn.f32x2, CREDUX = redux.sync, LDGSTS = cp.async, SYNCS = mbarrier\n")
for k, c in hist.items():
    name = demangle(k)
    name = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", ""))[:110]
    tot = sum(c.values())
    marks = " ".join(f"{m}={c[m]}" for m in MARK if c[m])
    top = " ".join(f"{o}:{n}" for o, n in c.most_common(8))
    print(f"{name}\n    {tot} instr | {marks or '-'}\n    {top}")
