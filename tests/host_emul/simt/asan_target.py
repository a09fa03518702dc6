"""TEST-ONLY: every hot kernel of the emulated library once, small shapes, ragged tails, every lane count, the three encoder
and token-pass forms, the ADPCM forms, plus bit-flipped packets -- run by tests/test_simt_emul.py under AddressSanitizer
(LD_PRELOAD of libasan, the library built with ASAN=1), where the library's "device" and pinned buffers are heap blocks and
its shared memory is static storage: a kernel that reads or writes outside them is reported.
Usage: python asan_target.py <libamvcuda_emul.so built with ASAN=1>"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(TESTS))
sys.path.insert(0, TESTS)
import amv_codec_tools_b200 as amv  # noqa: E402
from oracle_lib import Oracle, offsets_of, pack, synth_frames, synth_pcm  # noqa: E402

ctx = amv.AmvCuda(device=0, lib_path=sys.argv[1])
o = Oracle()
for (w, h, n, kind, q) in ((160, 120, 3, "sinus", 2), (72, 24, 4, "noise", 2), (48, 40, 5, "flat", 7), (208, 176, 2, "edges", 3)):
    y, u, v = synth_frames(n, w, h, seed=3, kind=kind)
    wpk, woff, wsz = o.encode_frames(y, u, v, w, h, q)
    for form in (2, 1, 0):
        ctx.set_option("encode_rounds", form)
        pk, off, sz, st = ctx.encode_frames(y, u, v, qscale=q)
        assert (st == 0).all() and np.array_equal(pk, wpk), ("encode", w, h, form)
    wy, wu, wv, _ = o.decode_frames(wpk, woff, wsz, w, h)
    rng = np.random.default_rng(w)
    units = [bytearray(wpk[int(woff[i]): int(woff[i]) + int(wsz[i])].tobytes()) for i in range(n)]
    for b in units[1:]:
        for _ in range(6):
            b[int(rng.integers(2, len(b) - 2))] ^= 1 << int(rng.integers(0, 8))
    bpk, boff, bsz = pack([bytes(b) for b in units])
    for tp in (2, 1, 0):
        ctx.set_option("decode_token_pass", tp)
        for lp in (0, 2, 5):
            ctx.set_option("decode_log2_lanes", lp)
            dy, du, dv, st = ctx.decode_frames(wpk, woff, wsz, w, h)
            assert (st == 0).all() and np.array_equal(dy, wy) and np.array_equal(du, wu) and np.array_equal(dv, wv), ("decode", w, h, tp, lp)
            dy, du, dv, st = ctx.decode_frames(bpk, boff, bsz, w, h)           # damaged scans: must end, inside their buffers
            assert st[0] == 0 and np.array_equal(dy[0], wy[0])
ctx.set_option("decode_token_pass", 2)
ctx.set_option("decode_log2_lanes", -1)
ctx.set_option("encode_rounds", 4)
rng = np.random.default_rng(5)
nsamp = (rng.integers(0, 900, 70) * 2).astype(np.uint32)
nsamp[:3] = [0, 2, 1378]
pcm = synth_pcm(int(nsamp.sum()) + 2, seed=6, kind="noise")
poff = offsets_of(nsamp)
step_in = (np.arange(70) * 7 % 89).astype(np.int16)
wout, _, wsz, wso = o.adpcm_encode(pcm, poff, nsamp, step_in)
for form in (0, 1, 2):
    ctx.set_option("adpcm_form", form)
    out, ooff, osz, so, st = ctx.adpcm_encode(pcm, poff, nsamp, step_in)
    assert (st == 0).all() and np.array_equal(out, wout) and np.array_equal(so, wso), ("adpcm encode", form)
    dec, _, dst = ctx.adpcm_decode(out, ooff, osz)
    assert (dst == 0).all() and np.array_equal(dec, o.adpcm_decode(out, ooff, osz)[0]), ("adpcm decode", form)
# range conversion and the scaler with its range steps (their GPU tests hold torch tensors; here through host buffers)
y, u, v = synth_frames(2, 102, 56, seed=8, kind="noise")
for direction in (0, 1):
    cy, cu, cv = ctx.convert_range(y, u, v, direction)
    wy, wu, wv = o.convert_range(y, u, v, direction)
    assert np.array_equal(cy, wy) and np.array_equal(cu, wu) and np.array_equal(cv, wv), ("range", direction)
for flags in (0, 1, 2, 3):
    sy, su, sv = ctx.scale_frames(y, u, v, 64, 48, flags=flags)
    wy, wu, wv = o.sws_scale(y, u, v, 64, 48, flags & 1, flags & 2)
    assert np.array_equal(sy, wy) and np.array_equal(su, wu) and np.array_equal(sv, wv), ("scale", flags)
print("asan target ok")
