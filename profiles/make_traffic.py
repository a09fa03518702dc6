#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --page raw --csv` export of profiles/prof_target.py <frames> 0:
DRAM bytes per frame of each hot kernel (first instance of each).  usage: make_traffic.py raw.csv frames [note]"""
import csv
import json
import os
import sys

rows = list(csv.reader(open(sys.argv[1])))
frames = int(sys.argv[2])
H = rows[0]
U = rows[1]
kinds = {"k_encode16": "encode", "k_compact": "compact", "k_unstuff": "unstuff", "k_vlc_tokens": "tokens", "k_idct<": "idct"}


def val(r, key):
    i = H.index(key)
    v = float(r[i])
    unit = U[i].lower()
    return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)


out = {}
for r in rows[2:]:
    name = r[H.index("Kernel Name")]
    for pat, k in kinds.items():
        if pat in name and k not in out:
            rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
            i = H.index("gpu__time_duration.sum")
            dur = float(r[i]) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(U[i].lower().replace("msecond", "ms").replace("usecond", "us").replace("nsecond", "ns").replace("second", "s"), 1.0)
            out[k] = {"dram_bytes_per_frame": (rd + wr) / frames, "read": rd, "write": wr, "frames": frames, "duration_ms": dur}
out["_source"] = sys.argv[3] if len(sys.argv) > 3 else "ncu --set full --clock-control none, profiles/prof_target.py %d 0" % frames
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
