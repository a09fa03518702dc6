#!/usr/bin/env python
"""Regenerate tests/golden/resample_golden.npz from the UNMODIFIED reference (needs oracle/_ref/libamvref.so):
small pictures through img_resample_init + img_resample (what the fork's sws_scale runs for `-s WxH`), short
PCM streams through audio_resample_init(1, ch, out_rate, in_rate) + audio_resample fed in packets, and the
filter banks av_build_filter made for them."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import Ref, chroma_dims, synth_frames, synth_pcm  # noqa: E402

SCALE_CASES = [(64, 48, 32, 24), (48, 32, 104, 88), (51, 37, 40, 24), (96, 80, 33, 21), (40, 24, 40, 56)]
AUDIO_CASES = [(44100, 2, 6000, 1152), (48000, 1, 5000, 1024), (8000, 1, 1500, 320), (22050, 2, 3000, 512),
               (32000, 2, 4000, 4000)]


def main():
    ref = Ref()
    out = {}
    rng = np.random.default_rng(77)
    for (iw, ih, ow, oh) in SCALE_CASES:
        y, u, v = synth_frames(2, iw, ih, seed=31, kind="sinus")
        icw, ich = chroma_dims(iw, ih)
        y[1] = rng.integers(0, 256, (ih, iw), dtype=np.uint8)
        u[1] = rng.integers(0, 256, (ich, icw), dtype=np.uint8)
        v[1] = rng.integers(0, 256, (ich, icw), dtype=np.uint8)
        oy, ou, ov = ref.scale_frames(y, u, v, ow, oh, fill=7)
        key = "scale_%dx%d_%dx%d" % (iw, ih, ow, oh)
        for nm, a in (("y", y), ("u", u), ("v", v), ("oy", oy), ("ou", ou), ("ov", ov)):
            out["%s/%s" % (key, nm)] = a
    for (rate, ch, n, chunk) in AUDIO_CASES:
        pcm = synth_pcm(n * ch, seed=rate + ch, kind="tones")
        pcm[: 64 * ch] = rng.integers(-32768, 32768, 64 * ch).astype(np.int16)        # full-scale start: the mirrored taps
        pcm[-200 * ch:] = np.where(rng.random(200 * ch) < 0.5, 32767, -32768).astype(np.int16)   # saturation
        key = "audio_%d_%d" % (rate, ch)
        out[key + "/pcm"] = pcm
        out[key + "/out"] = ref.audio_resample(pcm, ch, rate, 22050, chunk=chunk)
        out[key + "/bank_rows"] = ref.resample_bank(rate, 22050)[[0, 1, 511, 512, 1023]]
    path = os.path.join(HERE, "resample_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
