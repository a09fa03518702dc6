#!/usr/bin/env python
"""Small fixed workload for ncu: one warm-up + one measured encode/decode pass of N 320x240 frames
(device resident), optionally ADPCM.  Usage: python profiles/prof_target.py [frames] [log2_lanes]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amv_codec_tools_b200 as amv  # noqa: E402
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
log2p = int(sys.argv[2]) if len(sys.argv) > 2 else -1
dev = torch.device("cuda", 0)
ctx = amv.AmvCuda(device=0)
ctx.set_option("decode_log2_lanes", log2p)
for kv in sys.argv[3:]:                      # further options as key=value
    ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
W, H, CW, CH = bench.W, bench.H, bench.CW, bench.CH
Y, U, V = bench.synth_frames_torch(n, 0, dev, 1)
cap = n * 24 * 1024
pk = torch.empty(cap, dtype=torch.uint8, device=dev)
off = torch.zeros(n, dtype=torch.int64, device=dev)
size = torch.zeros(n, dtype=torch.int32, device=dev)
st = torch.zeros(n, dtype=torch.int32, device=dev)
DY, DU, DV = torch.empty_like(Y), torch.empty_like(U), torch.empty_like(V)
for _ in range(2):
    ctx.encode_frames_raw(Y, U, V, W, CW, W * H, CW * CH, n, W, H, None, pk, cap, bench.PKT_CAP, amv.LAYOUT_PACKED, off, size,
                          st, amv.MEM_DEVICE)
    ctx.decode_frames_raw(pk, cap, off, size, n, W, H, DY, DU, DV, W, CW, W * H, CW * CH, st, amv.MEM_DEVICE)
ctx.sync()
# ADPCM: 65536 chunks of 1378 samples
nc, ns = 65536, 1378
pcm = (8000 * torch.sin(torch.arange(nc * ns, device=dev, dtype=torch.float32) * 0.1254)).to(torch.int16)
poff = (torch.arange(nc, device=dev, dtype=torch.int64) * ns)
nsam = torch.full((nc,), ns, dtype=torch.int32, device=dev)
ooff = (torch.arange(nc, device=dev, dtype=torch.int64) * (8 + ns // 2))
osz = torch.full((nc,), 8 + ns // 2, dtype=torch.int32, device=dev)
out = torch.zeros(nc * (8 + ns // 2), dtype=torch.uint8, device=dev)
so = torch.zeros(nc, dtype=torch.int16, device=dev)
ast = torch.zeros(nc, dtype=torch.int32, device=dev)
dec = torch.zeros(nc * ns, dtype=torch.int16, device=dev)
for _ in range(2):
    ctx.adpcm_enc_chunks_raw(pcm, nc * ns, poff, nsam, None, so, nc, out, out.numel(), ooff, ast, amv.MEM_DEVICE)
    ctx.adpcm_dec_chunks_raw(out, out.numel(), ooff, osz, nc, dec, nc * ns, poff, ast, amv.MEM_DEVICE)
ctx.sync()
assert int(st.abs().sum()) == 0 and int(ast.abs().sum()) == 0
print("ok", n, int(size.sum()))
