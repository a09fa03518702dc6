#!/usr/bin/env python
"""Regenerate tests/golden/mjpeg_golden.npz from the UNMODIFIED reference (needs oracle/_ref/libamvref.so):
plain JPEG frames from the reference's mjpeg_encoder (tables in the stream), some with the quantiser table of
their DQT segment overwritten (same scan, other tables), and what the reference's mjpeg_decoder makes of them."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import Oracle, Ref, jpeg_encode_simple, mjpeg_with_dqt, pack, resample_chroma, synth_frames  # noqa: E402


def main():
    ref = Ref()
    out, names = {}, []
    for kind, w, h, n, dqt_seed, s422 in [("sinus", 160, 120, 2, None, 0), ("sinus", 208, 176, 1, 3, 0), ("flat", 32, 32, 2, None, 0),
                                          ("sinus", 72, 24, 2, 4, 0), ("noise", 48, 40, 1, None, 0),
                                          ("sinus", 160, 120, 2, None, 1), ("noise", 104, 56, 2, 5, 1), ("edges", 72, 24, 1, None, 1)]:
        y, u, v = synth_frames(n, w, h, seed=27, kind=kind)
        chroma = None
        if s422:        # YUVJ422P input: full-height chroma planes (the encoder writes 2x2 / 1x2 / 1x2 sampling)
            u, v = (np.repeat(a, 2, axis=1)[:, :h, :].copy() for a in (u, v))
            u[:, 1::2, :] = np.clip(u[:, 1::2, :].astype(int) + 3, 0, 255).astype(np.uint8)      # rows differ
            chroma = (u.shape[2], h)
        pk, off, sz = ref.mjpeg_encode_frames(y, u, v, w, h)
        if dqt_seed is not None:
            pk = mjpeg_with_dqt(pk, off, sz, dqt_seed)
        dy, du, dv, got, _ = ref.decode_frames(pk, off, sz, w, h, mjpeg=True, chroma=chroma)
        assert (got != 0).all()
        key = "%s_%dx%d_d%s%s" % (kind, w, h, "x" if dqt_seed is None else dqt_seed, "s422" if s422 else "")
        names.append(key)
        for nm, a in (("pk", pk), ("off", off), ("sz", sz), ("dy", dy), ("du", du), ("dv", dv)):
            out["%s/%s" % (key, nm)] = a
    # the other layouts the decoder accepts (mjpegdec.c:283-311) -- the reference has no encoder for them, so the
    # frames come from a minimal JPEG writer and the reference DECODER says what they mean
    o = Oracle()
    for tag, samp in (("s211", ((2, 1), (1, 1))), ("s444", ((1, 1), (1, 1)))):
        for kind, w, h, n in (("sinus", 160, 120, 2), ("noise", 72, 40, 2), ("edges", 102, 56, 1)):
            y, u, v = synth_frames(n, w, h, seed=28, kind=kind)
            U, V = resample_chroma(u, w, h, samp), resample_chroma(v, w, h, samp)
            pk, off, sz = pack([jpeg_encode_simple(o, y[i], U[i], V[i], samp).tobytes() for i in range(n)])
            dy, du, dv, got, _ = ref.decode_frames(pk, off, sz, w, h, mjpeg=True, chroma=(U.shape[2], U.shape[1]))
            assert (got != 0).all()
            key = "%s_%dx%d_dx%s" % (kind, w, h, tag)
            names.append(key)
            for nm, a in (("pk", pk), ("off", off), ("sz", sz), ("dy", dy), ("du", du), ("dv", dv)):
                out["%s/%s" % (key, nm)] = a
    # restart intervals (DRI + RSTn), again from the minimal writer and the reference decoder
    for tag, samp, restart in (("r1", ((2, 2), (1, 1)), 1), ("r5", ((2, 2), (1, 1)), 5), ("r3s211", ((2, 1), (1, 1)), 3)):
        for kind, w, h, n in (("sinus", 160, 120, 2), ("noise", 72, 40, 2)):
            y, u, v = synth_frames(n, w, h, seed=29, kind=kind)
            U, V = resample_chroma(u, w, h, samp), resample_chroma(v, w, h, samp)
            pk, off, sz = pack([jpeg_encode_simple(o, y[i], U[i], V[i], samp, restart=restart).tobytes() for i in range(n)])
            dy, du, dv, got, _ = ref.decode_frames(pk, off, sz, w, h, mjpeg=True, chroma=(U.shape[2], U.shape[1]))
            assert (got != 0).all()
            key = "%s_%dx%d_dx%s" % (kind, w, h, tag)
            names.append(key)
            for nm, a in (("pk", pk), ("off", off), ("sz", sz), ("dy", dy), ("du", du), ("dv", dv)):
                out["%s/%s" % (key, nm)] = a
    out["cases"] = np.frombuffer("\n".join(names).encode(), np.uint8)
    path = os.path.join(HERE, "mjpeg_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
