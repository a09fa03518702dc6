#!/usr/bin/env python
"""Host<->device copy bandwidth of the box (pinned memory, cudaMemcpyAsync through torch): H2D alone,
D2H alone and both directions at once.  The e2e leg of bench.py moves 133 KB per frame in EACH
direction, so `both` / 133 KB is the ceiling of the end-to-end frames/s on one GPU."""
import json
import sys

import torch

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = 8
dev = torch.device("cuda", 0)
n = mb << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device=dev)
d_out = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream())
    s2.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    return n * reps / (e0.elapsed_time(e1) / 1e3) / 1e9


for _ in range(2):
    run(True, True)
res = {"buffer_MiB": mb, "h2d_alone_GBs": run(True, False), "d2h_alone_GBs": run(False, True),
       "both_each_direction_GBs": run(True, True)}
res["e2e_frames_per_s_ceiling_320x240"] = res["both_each_direction_GBs"] * 1e9 / 133232.0
print(json.dumps(res))
