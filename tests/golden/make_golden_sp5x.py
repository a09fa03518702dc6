#!/usr/bin/env python
"""Regenerate tests/golden/sp5x_golden.npz from the UNMODIFIED reference (needs oracle/_ref/libamvref.so):
SP5X packets built from reference-encoded AMV scans (14 header bytes + scan with literal FF bytes; the
reference has no SP5X encoder) and what the reference's sp5x_decoder makes of them."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import Oracle, Ref, sp5x_from_amv, synth_frames  # noqa: E402


def main():
    ref, o = Ref(), Oracle()
    out, names = {}, []
    for kind, w, h, n, quality in [("sinus", 160, 120, 2, 0), ("sinus", 208, 176, 1, 5 * 118), ("flat", 32, 32, 2, 0),
                                   ("sinus", 72, 24, 2, 0), ("noise", 48, 40, 1, 20 * 118)]:
        y, u, v = synth_frames(n, w, h, seed=25, kind=kind)
        pk, off, sz = ref.encode_frames(y, u, v, w, h, quality=quality)
        sp, soff, ssz = sp5x_from_amv(o, pk, off, sz)
        assert all(sp[int(a) + 14:int(a) + int(b)].tobytes().count(b"\xff") <= 400 for a, b in zip(soff, ssz))
        dy, du, dv, got, _ = ref.decode_frames(sp, soff, ssz, w, h, sp5x=True)
        assert (got != 0).all()
        key = "%s_%dx%d_q%d" % (kind, w, h, quality)
        names.append(key)
        for nm, a in (("pk", sp), ("off", soff), ("sz", ssz), ("dy", dy), ("du", du), ("dv", dv)):
            out["%s/%s" % (key, nm)] = a
    out["cases"] = np.frombuffer("\n".join(names).encode(), np.uint8)
    path = os.path.join(HERE, "sp5x_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
