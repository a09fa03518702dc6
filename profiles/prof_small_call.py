"""One frame per call through AMV_MEM_HOST (what the AVCodec shims do): wall time per call and the device time of each kernel.
Usage: python profiles/prof_small_call.py [width height [calls]]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

import amv_codec_tools_b200 as amv
from oracle_lib import synth_frames, chroma_dims

w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (320, 240)
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 300
ctx = amv.AmvCuda(device=0)
y, u, v = synth_frames(1, w, h, seed=5)
cw, ch = chroma_dims(w, h)
pk, off, sz, st = ctx.encode_frames(y, u, v, qscale=2)
assert (st == 0).all()
dy = np.zeros((1, h, w), np.uint8); du = np.zeros((1, ch, cw), np.uint8); dv = np.zeros((1, ch, cw), np.uint8)
st = np.zeros(1, np.int32)
out = {"width": w, "height": h, "packet_bytes": int(sz[0]), "calls": calls}
KINDS = ("unstuff", "sync", "tokens", "idct", "encode", "compact")
for prof in (0, 1):
    ctx.set_option("profile_events", prof)
    for name, fn in (("decode", lambda: ctx.decode_frames_raw(pk, pk.nbytes, off, sz, 1, w, h, dy, du, dv, w, cw, w * h, cw * ch, st, amv.MEM_HOST)),
                     ("encode", lambda: ctx.encode_frames(y, u, v, qscale=2))):
        for _ in range(20): fn()
        if prof:
            for k in KINDS: ctx.get_stat(k + "_kernel_ns")       # the ns query consumes the samples so far
        t0 = time.perf_counter()
        for _ in range(calls): fn()
        dt = (time.perf_counter() - t0) / calls
        rec = {"wall_us_per_call": round(dt * 1e6, 1)}
        if prof:
            for k in KINDS:
                ln, ns = ctx.get_stat(k + "_kernel_launches"), ctx.get_stat(k + "_kernel_ns")
                if ln > 0: rec[k + "_us"] = round(ns / calls / 1e3, 1)
            rec["kernel_us_sum"] = round(sum(v for k, v in rec.items() if k.endswith("_us")), 1)
        out[name + ("_profiled" if prof else "")] = rec
print(json.dumps(out))
