#!/usr/bin/env python
"""Attribute an `ncu --page source --csv` (SASS view) export to CUDA source lines, using the
line table of the matching cubin (nvdisasm -g -c): executed warp-instructions and stall samples
per source line.  usage: ncu_lines.py src.csv kernel_substring file.sass mangled_substring [min_pct]"""
import csv
import re
import sys
from collections import defaultdict

src_csv, kname, sass_file, mangled = sys.argv[1:5]
min_pct = float(sys.argv[5]) if len(sys.argv) > 5 else 0.5
# --- line table from nvdisasm: ordered list of (file, line) per SASS instruction of the function
lines, cur, infn = [], None, False
for l in open(sass_file):
    if l.startswith(".text."):
        infn = mangled in l
        continue
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
# --- executed counts from ncu, same order
rows = list(csv.reader(open(src_csv)))
blocks, b = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        b = {"name": r[1], "rows": []}
        blocks.append(b)
    elif b is not None:
        b["rows"].append(r)
blk = [x for x in blocks if kname in x["name"]][0]
H = blk["rows"][0]
ie, ismp = H.index("Instructions Executed"), H.index("# Samples")
data = [r for r in blk["rows"][1:] if len(r) > ie and r[ie] != ""]
assert len(data) == len(lines), (len(data), len(lines))
agg = defaultdict(lambda: [0, 0, 0])
for r, ln in zip(data, lines):
    a = agg[ln]
    a[0] += int(r[ie]); a[1] += int(r[ismp] or 0); a[2] += 1
tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
srcs = {}
print("%s: %d warp-instr, %d samples" % (kname, tot, tots))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    pct = 100.0 * a[0] / tot
    if pct < min_pct:
        continue
    if f not in srcs:
        try:
            srcs[f] = open("/root/repo/amv-codec-tools_b200/csrc/" + f).read().split("\n")
        except Exception:
            srcs[f] = []
    text = srcs[f][ln - 1].strip()[:90] if 0 < ln <= len(srcs[f]) else ""
    print("%5.1f%% instr %5.1f%% smp %4d sass  %s:%d  %s" % (pct, 100.0 * a[1] / max(tots, 1), a[2], f, ln, text))
