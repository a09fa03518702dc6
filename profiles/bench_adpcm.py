#!/usr/bin/env python
"""BASELINE config 3: IMA-ADPCM-AMV chunk encode/decode, 1M independent 22050 Hz mono chunks of
1378 samples (n = 689 nibble bytes), device resident; prints one JSON line with chunks/s per
direction and the HBM roofline fraction (algorithmic bytes 5n+8 = 3453 per chunk).
The library is audited against the committed golden chunks of the reference."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amv_codec_tools_b200 as amv  # noqa: E402
import bench  # noqa: E402

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ns = 1378
dev = torch.device("cuda", 0)
ctx = amv.AmvCuda(device=0)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
ctx.set_option("profile_events", 1)
g = torch.Generator(device=dev); g.manual_seed(3)
t = torch.arange(nc * ns, device=dev, dtype=torch.float32)
pcm = (8000 * torch.sin(t * (2 * np.pi * 440 / 22050)) + 2000 * torch.sin(t * (2 * np.pi * 1234 / 22050))
       + 300 * torch.randn(nc * ns, device=dev, generator=g)).round().clamp(-32768, 32767).to(torch.int16)
del t
poff = torch.arange(nc, device=dev, dtype=torch.int64) * ns
nsam = torch.full((nc,), ns, dtype=torch.int32, device=dev)
csz = 8 + ns // 2
ooff = torch.arange(nc, device=dev, dtype=torch.int64) * csz
osz = torch.full((nc,), csz, dtype=torch.int32, device=dev)
out = torch.zeros(nc * csz, dtype=torch.uint8, device=dev)
so = torch.zeros(nc, dtype=torch.int16, device=dev)
st = torch.zeros(nc, dtype=torch.int32, device=dev)
dec = torch.zeros(nc * ns, dtype=torch.int16, device=dev)


def step():
    ctx.adpcm_enc_chunks_raw(pcm, nc * ns, poff, nsam, None, so, nc, out, out.numel(), ooff, st, amv.MEM_DEVICE)
    ctx.adpcm_dec_chunks_raw(out, out.numel(), ooff, osz, nc, dec, nc * ns, poff, st, amv.MEM_DEVICE)


for _ in range(3):
    step()
ctx.sync()
assert int(st.abs().sum()) == 0
for k in ("adpcm_enc", "adpcm_dec"):
    ctx.get_stat(k + "_kernel_ns")
for _ in range(steps):
    step()
ctx.sync()
res = {}
for k in ("adpcm_enc", "adpcm_dec"):
    cnt = ctx.get_stat(k + "_kernel_launches")
    ms = ctx.get_stat(k + "_kernel_ns") / 1e6 / max(cnt, 1)
    res[k] = ms
peak, src = bench.load_peaks()
bytes_per_chunk = 5 * (ns // 2) + 8
# audit without the oracle: the committed golden chunks of the reference (tests/golden/amv_golden.npz)
G = np.load(os.path.join(ROOT, "tests", "golden", "amv_golden.npz"))
ok = True
for kind in ("tones", "noise", "square"):
    k = "adpcm_%s/" % kind
    gout, goff, gsz, gcons = G[k + "out"], G[k + "off"], G[k + "sz"], G[k + "cons"]
    first = np.array([0, len(gcons)], np.uint32)
    gpoff = np.concatenate([[0], np.cumsum(gcons)[:-1]]).astype(np.uint64)
    step0 = np.array([int(gout[2]) | (int(gout[3]) << 8)], np.int16)
    eo, _, esz, _, est = ctx.adpcm_encode_streams(G[k + "src"], gpoff, gcons, first, step0)
    dp, _, dst = ctx.adpcm_decode(gout, goff, gsz)
    ok = ok and bool((est == 0).all()) and np.array_equal(eo, gout) and bool((dst == 0).all()) and np.array_equal(dp, G[k + "dec"])
na = 3 * 6
print(json.dumps({
    "metric": "IMA-ADPCM-AMV chunks/sec (1378-sample chunks)", "chunks": nc, "steps": steps,
    "encode_chunks_per_s": nc / (res["adpcm_enc"] / 1e3), "decode_chunks_per_s": nc / (res["adpcm_dec"] / 1e3),
    "encode_ms": res["adpcm_enc"], "decode_ms": res["adpcm_dec"], "bytes_per_chunk": bytes_per_chunk,
    "roofline_encode": {"achieved_GBs": bytes_per_chunk * nc / (res["adpcm_enc"] / 1e3) / 1e9, "peak": peak,
                        "frac": bytes_per_chunk * nc / (res["adpcm_enc"] / 1e3) / 1e9 / peak},
    "roofline_decode": {"achieved_GBs": bytes_per_chunk * nc / (res["adpcm_dec"] / 1e3) / 1e9, "peak": peak,
                        "frac": bytes_per_chunk * nc / (res["adpcm_dec"] / 1e3) / 1e9 / peak},
    "audit_vs_golden": {"chunks": na, "ok": bool(ok)}}))
