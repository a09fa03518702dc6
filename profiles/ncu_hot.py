#!/usr/bin/env python
"""Hot SASS lines of each kernel in an `ncu --page source --csv` export: top stall-sample sites
and where the executed-instruction count changes (loop structure)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1][:80], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks:
    H = b["rows"][0]
    ia, isrc, ie, ismp, it = (H.index(x) for x in ("Address", "Source", "Instructions Executed", "# Samples", "Avg. Threads Executed"))
    data = [r for r in b["rows"][1:] if len(r) > ie and r[ie]]
    tot = sum(int(r[ie]) for r in data)
    tots = sum(int(r[ismp] or 0) for r in data)
    print("=" * 100)
    print(b["name"], "|", len(data), "SASS instrs, executed", tot, "samples", tots)
    for r in sorted(data, key=lambda r: -int(r[ismp] or 0))[:top_n]:
        print("  %5.1f%% smp %11s exec thr=%-3s %s" % (100.0 * int(r[ismp] or 0) / max(tots, 1), r[ie], r[it], r[isrc][:84]))
    print("  -- executed-count profile along the program")
    prev = None
    for i, r in enumerate(data):
        e = int(r[ie])
        if prev is None or abs(e - prev) > 0.25 * max(prev, 1):
            print("  #%-5d %12d thr=%-3s %s" % (i, e, r[it], r[isrc][:70]))
        prev = e
