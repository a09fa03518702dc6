#!/usr/bin/env python
"""Regenerate tests/golden/mjpeg_golden.npz from the UNMODIFIED reference (needs oracle/_ref/libamvref.so):
plain JPEG frames from the reference's mjpeg_encoder (tables in the stream), some with the quantiser table of
their DQT segment overwritten (same scan, other tables), and what the reference's mjpeg_decoder makes of them."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import Ref, mjpeg_with_dqt, synth_frames  # noqa: E402


def main():
    ref = Ref()
    out, names = {}, []
    for kind, w, h, n, dqt_seed in [("sinus", 160, 120, 2, None), ("sinus", 208, 176, 1, 3), ("flat", 32, 32, 2, None),
                                    ("sinus", 72, 24, 2, 4), ("noise", 48, 40, 1, None)]:
        y, u, v = synth_frames(n, w, h, seed=27, kind=kind)
        pk, off, sz = ref.mjpeg_encode_frames(y, u, v, w, h)
        if dqt_seed is not None:
            pk = mjpeg_with_dqt(pk, off, sz, dqt_seed)
        dy, du, dv, got, _ = ref.decode_frames(pk, off, sz, w, h, mjpeg=True)
        assert (got != 0).all()
        key = "%s_%dx%d_d%s" % (kind, w, h, "x" if dqt_seed is None else dqt_seed)
        names.append(key)
        for nm, a in (("pk", pk), ("off", off), ("sz", sz), ("dy", dy), ("du", du), ("dv", dv)):
            out["%s/%s" % (key, nm)] = a
    out["cases"] = np.frombuffer("\n".join(names).encode(), np.uint8)
    path = os.path.join(HERE, "mjpeg_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
