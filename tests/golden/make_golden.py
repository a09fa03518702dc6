#!/usr/bin/env python
"""Regenerate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference and oracle/_ref/libamvref.so):
    python tests/golden/make_golden.py
Every array in the files is either an input we synthesised or an output of the
reference codecs (AMVmuxer libavcodec 51.47.1, generic C paths) on that input;
`AMV1` entries are the first packets of the reference's own fixture
C-AMVDecoder/bin/AMV1.amv with the reference decoder's output.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import FIXTURE_AMV, Ref, pack, synth_frames, synth_pcm, walk_amv  # noqa: E402


def main():
    ref = Ref()
    out = {"ref_version": np.frombuffer(ref.version().encode(), np.uint8)}
    cases = [("sinus", 160, 120, 2, 0), ("noise", 48, 40, 2, 0), ("edges", 64, 48, 2, 0), ("flat", 32, 32, 2, 0),
             ("sinus", 208, 176, 1, 5 * 118), ("sinus", 72, 24, 2, 31 * 118)]
    names = []
    for kind, w, h, n, quality in cases:
        y, u, v = synth_frames(n, w, h, seed=21, kind=kind)
        pk, off, sz = ref.encode_frames(y, u, v, w, h, quality=quality)
        dy, du, dv, got, _ = ref.decode_frames(pk, off, sz, w, h)
        assert (got != 0).all()
        key = "%s_%dx%d_q%d" % (kind, w, h, quality)
        names.append(key)
        for nm, a in (("y", y), ("u", u), ("v", v), ("pk", pk), ("off", off), ("sz", sz), ("dy", dy), ("du", du), ("dv", dv)):
            out["%s/%s" % (key, nm)] = a
    out["video_cases"] = np.frombuffer("\n".join(names).encode(), np.uint8)

    # reference fixture: first 4 video packets / audio chunks + whole-clip hashes
    w, h, fps, vids, auds = walk_amv(open(FIXTURE_AMV, "rb").read())
    pk, off, sz = pack(vids)
    dy, du, dv, got, _ = ref.decode_frames(pk, off, sz, w, h)
    m = hashlib.md5()
    for i in range(len(vids)):
        m.update(dy[i].tobytes()); m.update(du[i].tobytes()); m.update(dv[i].tobytes())
    out["AMV1/dims"] = np.array([w, h, fps, len(vids)], np.int32)
    out["AMV1/planes_md5"] = np.frombuffer(m.hexdigest().encode(), np.uint8)
    k = 4
    pk4, off4, sz4 = pack(vids[:k])
    out["AMV1/pk"], out["AMV1/off"], out["AMV1/sz"] = pk4, off4, sz4
    out["AMV1/dy"], out["AMV1/du"], out["AMV1/dv"] = dy[:k], du[:k], dv[:k]
    ak, aoff, asz = pack(auds)
    pcm, poff, ns = ref.adpcm_decode(ak, aoff, asz)
    out["AMV1/pcm_md5"] = np.frombuffer(hashlib.md5(pcm.tobytes()).hexdigest().encode(), np.uint8)
    ak4, aoff4, asz4 = pack(auds[:k])
    out["AMV1/ak"], out["AMV1/aoff"], out["AMV1/asz"] = ak4, aoff4, asz4
    out["AMV1/pcm"] = pcm[: int(ns[:k].sum())]

    # ADPCM encoder: one chained stream per signal kind
    for kind in ("tones", "noise", "square"):
        src = synth_pcm(1378 * 6 + 100, seed=22, kind=kind)
        eo, eoff, esz, cons = ref.adpcm_encode_stream(src, 1378)
        dp, _, _ = ref.adpcm_decode(eo, eoff, esz)
        out["adpcm_%s/src" % kind] = src
        out["adpcm_%s/out" % kind], out["adpcm_%s/off" % kind] = eo, eoff
        out["adpcm_%s/sz" % kind], out["adpcm_%s/cons" % kind] = esz, cons
        out["adpcm_%s/dec" % kind] = dp
    np.savez_compressed(os.path.join(HERE, "amv_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "amv_golden.npz"), os.path.getsize(os.path.join(HERE, "amv_golden.npz")), "bytes")


if __name__ == "__main__":
    main()
