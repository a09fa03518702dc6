#!/usr/bin/env python
"""Digest of BASELINE config 1 as the REFERENCE makes it (run where oracle/_ref is built, i.e. where /root/reference
is mounted): 1 000 frames 160x120 + 1 000 ADPCM chunks (22050 Hz, 16 fps), encoded by the reference's amv /
adpcm_ima_amv encoders, muxed by its amv muxer, demuxed and decoded by its own demuxer / decoders.
Writes tests/golden/config1_digest.json (sha256 of the file, of the decoded planes, of the decoded PCM)."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import Ref, pack, synth_frames, synth_pcm  # noqa: E402

w, h, n, ns = 160, 120, 1000, 1378
ref = Ref()
y, u, v = synth_frames(n, w, h, seed=1, kind="sinus")
pcm = synth_pcm(ns * n + 4096, seed=1, kind="tones")
vpk, voff, vsz = ref.encode_frames(y, u, v, w, h, quality=0)
apk, aoff, asz, cons = ref.adpcm_encode_stream(pcm, ns, max_chunks=n)
data = ref.mux(w, h, 16, 22050, vpk, voff, vsz, apk, aoff[:n], asz[:n])
info, rv, ra = ref.demux(data)
rpk, roff, rsz = pack(rv)
ry, ru, rvv, got, _ = ref.decode_frames(rpk, roff, rsz, w, h)
ak, ao, az = pack(ra)
rpcm, _, _ = ref.adpcm_decode(ak, ao, az)
out = {"what": "BASELINE config 1: reference-encoded, reference-muxed 160x120 clip, 1000 frames + 1000 ADPCM chunks",
       "reference": ref.version(), "file_bytes": len(data), "file_sha256": hashlib.sha256(data).hexdigest(),
       "planes_sha256": hashlib.sha256(ry.tobytes() + ru.tobytes() + rvv.tobytes()).hexdigest(),
       "pcm_sha256": hashlib.sha256(rpcm.tobytes()).hexdigest(), "nvideo": int(info[4]), "naudio": int(info[5])}
json.dump(out, open(os.path.join(HERE, "config1_digest.json"), "w"), indent=1)
print(out)
