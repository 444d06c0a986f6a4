#!/usr/bin/env python
"""cuobjdump -sass opcode histogram of the kernels whose name contains a pattern (profiles/*_sass_histogram.txt).

    python tools/sass_histogram.py font-ocr_b200/libfocr_b200.so scan_tc_kernel > profiles/r2_scan_tc_sass_histogram.txt
"""
import collections
import re
import subprocess
import sys

so, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, hist = None, collections.defaultdict(collections.Counter)
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
    if m and cur:
        hist[cur][m.group(1)] += 1
print(f"# cuobjdump -sass {so}: opcode histogram of the kernels matching '{pat}'")
print("# UTCIMMA = tcgen05.mma kind::i8, UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (TMA bulk copy),")
print("# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, USETMAXREG = setmaxnreg, IDP = dp4a")
key = ["UTCIMMA", "UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR", "SYNCS", "USETMAXREG", "UTMALDG", "ELECT", "R2UR", "FMNMX3", "VOTE", "BAR", "IDP"]
for fn, c in sorted(hist.items()):
    if pat not in fn:
        continue
    print(f"\n{fn}: {sum(c.values())} instructions")
    print("  " + "  ".join(f"{k}={c.get(k, 0)}" for k in key))
    print("  top: " + ", ".join(f"{k}={v}" for k, v in c.most_common(14)))
