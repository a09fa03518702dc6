#!/usr/bin/env python
"""Concurrent host<->device copy bandwidth of ALL ranks of a torchrun job (pinned memory, cudaMemcpyAsync through torch), every
rank on its own GPU and its own slice of the host cores: H2D alone, D2H alone, both at once -- per rank and summed.  What the
box's host memory / PCIe fabric gives N GPUs at the same time is the ceiling of bench.py's e2e leg at N GPUs
(133 KB per 320x240 frame in EACH direction).
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/pcie_probe_multi.py [MiB]"""
import json
import os
import sys

import torch
import torch.distributed as dist

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = 8
world, rank, lrank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
try:
    cores = sorted(os.sched_getaffinity(0))
    per = len(cores) // world
    if per >= 1:
        os.sched_setaffinity(0, set(cores[lrank * per:(lrank + 1) * per]))
except Exception:
    pass
torch.cuda.set_device(lrank)
dev = torch.device("cuda", lrank)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
n = mb << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_in.fill_(1)
h_out = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out.fill_(2)
d_in = torch.empty(n, dtype=torch.uint8, device=dev)
d_out = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
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
mine = [run(True, False), run(False, True), run(True, True)]
t = torch.tensor(mine, dtype=torch.float64, device=dev)
allt = [torch.zeros_like(t) for _ in range(world)]
if world > 1:
    dist.all_gather(allt, t)
else:
    allt = [t]
if rank == 0:
    rows = [a.cpu().tolist() for a in allt]
    res = {"n_gpus": world, "buffer_MiB": mb, "host_cores": len(cores),
           "h2d_alone_GBs_per_gpu": [round(r[0], 1) for r in rows], "d2h_alone_GBs_per_gpu": [round(r[1], 1) for r in rows],
           "both_each_direction_GBs_per_gpu": [round(r[2], 1) for r in rows],
           "h2d_alone_GBs_sum": round(sum(r[0] for r in rows), 1), "d2h_alone_GBs_sum": round(sum(r[1] for r in rows), 1),
           "both_each_direction_GBs_sum": round(sum(r[2] for r in rows), 1)}
    res["e2e_frames_per_s_ceiling_320x240"] = round(res["both_each_direction_GBs_sum"] * 1e9 / 133232.0)
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
