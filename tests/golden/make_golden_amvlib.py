#!/usr/bin/env python
"""Regenerate tests/golden/amvlib_golden.npz from the UNMODIFIED reference amvlib.

Run in the build container (needs /root/reference, oracle/_ref/libamvref.so and libamvlibref.so):
    python tests/golden/make_golden_amvlib.py
Inputs are AMV packets made by the reference ffmpeg-fork encoder from our synthetic frames, plus the
head of the reference's own fixture C-AMVDecoder/bin/AMV1.amv; outputs are what the reference amvlib
(AmvVideoDecode / AmvAudioDecode, compiled in place) makes of them.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import FIXTURE_AMV, AmvlibRef, Ref, pack, synth_frames, walk_amv  # noqa: E402


def main():
    ref, alib = Ref(), AmvlibRef()
    out, names = {}, []
    for kind, w, h, n, quality in [("sinus", 160, 120, 2, 0), ("sinus", 208, 176, 1, 5 * 118), ("edges", 64, 48, 2, 12 * 118),
                                   ("flat", 32, 32, 2, 0), ("sinus", 72, 24, 2, 31 * 118), ("noise", 48, 40, 1, 10 * 118)]:
        y, u, v = synth_frames(n, w, h, seed=23, kind=kind)
        pk, off, sz = ref.encode_frames(y, u, v, w, h, quality=quality)
        bgr, ret = alib.video_decode(pk, off, sz, w, h)
        assert (ret == 0).all()
        key = "%s_%dx%d_q%d" % (kind, w, h, quality)
        names.append(key)
        out[key + "/pk"], out[key + "/off"], out[key + "/sz"], out[key + "/bgr"] = pk, off, sz, bgr
    out["video_cases"] = np.frombuffer("\n".join(names).encode(), np.uint8)
    w, h, fps, vids, auds = walk_amv(open(FIXTURE_AMV, "rb").read())
    pk, off, sz = pack(vids)
    bgr, ret = alib.video_decode(pk, off, sz, w, h)
    assert (ret == 0).all()
    out["AMV1/dims"] = np.array([w, h, fps, len(vids)], np.int32)
    out["AMV1/bgr_md5"] = np.frombuffer(hashlib.md5(bgr.tobytes()).hexdigest().encode(), np.uint8)
    k = 3
    out["AMV1/pk"], out["AMV1/off"], out["AMV1/sz"] = pack(vids[:k])
    out["AMV1/bgr"] = bgr[:k]
    ak, aoff, asz = pack(auds)
    pcm, poff, ns, aret = alib.audio_decode(ak, aoff, asz)
    assert (aret == 0).all()
    out["AMV1/pcm_md5"] = np.frombuffer(hashlib.md5(pcm.tobytes()).hexdigest().encode(), np.uint8)
    out["AMV1/ak"], out["AMV1/aoff"], out["AMV1/asz"] = pack(auds[:k])
    out["AMV1/pcm"] = pcm[: int(ns[:k].sum())]
    out["AMV1/nsamp"] = ns[:k]
    path = os.path.join(HERE, "amvlib_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
