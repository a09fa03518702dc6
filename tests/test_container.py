"""AMV container on the host: libamvcuda's amv_file_mux / amv_file_index against the
reference's own muxer and demuxer (libavformat amvenc.c / avidec.c compiled in place and driven like
ffmpeg.c drives them), against the reference's fixture, and against committed golden files.  Pure
host code: runs without a GPU."""
import hashlib
import os

import numpy as np
import pytest

import amv_codec_tools_b200 as amv
from oracle_lib import FIXTURE_AMV, Oracle, Ref, pack, synth_frames, synth_pcm, walk_amv

GOLD = os.path.join(os.path.dirname(__file__), "golden", "container_golden.npz")
needs_ref = pytest.mark.skipif(not Ref.available(), reason="oracle/_ref/libamvref.so not built")


@pytest.fixture(scope="module")
def lib():
    amv.build()
    return amv.load_library()


def make_units(w, h, n, seed=1):
    o = Oracle()
    y, u, v = synth_frames(n, w, h, seed=seed)
    pk, off, sz = o.encode_frames(y, u, v, w, h, 2)
    pcm = synth_pcm(1378 * n, seed=seed)
    ck, coff, csz, _ = o.adpcm_encode(pcm, np.arange(n, dtype=np.uint64) * 1378, np.full(n, 1378, np.uint32), np.zeros(n, np.int16))
    return pk, off, sz, ck, coff, csz


@needs_ref
@pytest.mark.parametrize("w,h,fps,n", [(160, 120, 16, 7), (320, 240, 12, 3), (208, 176, 16, 1), (128, 96, 10, 0), (160, 120, 16, 1000)])
def test_mux_bytes_identical_to_reference(lib, w, h, fps, n):
    if n == 1000:                                    # config 1 length: packets repeated, durations > 60 s
        pk, off, sz, ck, coff, csz = make_units(w, h, 8)
        idx = np.arange(n) % 8
        off, sz, coff, csz = off[idx], sz[idx], coff[idx], csz[idx]
    else:
        pk, off, sz, ck, coff, csz = make_units(w, h, max(n, 1))
        off, sz, coff, csz = off[:n], sz[:n], coff[:n], csz[:n]
    want = Ref().mux(w, h, fps, 22050, pk, off, sz, ck, coff, csz)
    got = amv.file_mux(w, h, fps, 22050, pk, off, sz, ck, coff, csz, lib=lib)
    assert got == want
    assert got.find(b"movi") == 0x138 and got.endswith(b"AMV_END_")        # compare_amv.c:30-41


@needs_ref
def test_index_matches_reference_demuxer(lib):
    w, h, fps, n = 208, 176, 16, 9
    pk, off, sz, ck, coff, csz = make_units(w, h, n, seed=5)
    data = amv.file_mux(w, h, fps, 22050, pk, off, sz, ck, coff, csz, lib=lib)
    rinfo, rv, ra = Ref().demux(data)
    info, vo, vs, ao, as_ = amv.file_index(data, lib=lib)
    assert (info.width, info.height, info.fps, info.sample_rate, info.nvideo, info.naudio) == tuple(rinfo.tolist())
    assert [data[int(o):int(o) + int(s)] for o, s in zip(vo, vs)] == rv
    assert [data[int(o):int(o) + int(s)] for o, s in zip(ao, as_)] == ra
    assert info.movi_offset == 0x138 and info.has_end_marker == 1 and info.truncated == 0 and info.channels == 1
    assert info.us_per_frame == 62500 and info.nb_frames_header == n


@needs_ref
@pytest.mark.skipif(not os.path.exists(FIXTURE_AMV), reason="reference fixture not mounted")
def test_index_reference_fixture(lib):
    """the real device clip: list sizes are all zero there, the walker must not trust them"""
    data = open(FIXTURE_AMV, "rb").read()
    rinfo, rv, ra = Ref().demux(data)
    info, vo, vs, ao, as_ = amv.file_index(data, lib=lib)
    assert (info.width, info.height, info.fps, info.sample_rate, info.nvideo, info.naudio) == (128, 96, 12, 16000, 252, 252)
    assert tuple(rinfo.tolist()) == (128, 96, 12, 16000, 252, 252)
    assert [data[int(o):int(o) + int(s)] for o, s in zip(vo, vs)] == rv
    assert [data[int(o):int(o) + int(s)] for o, s in zip(ao, as_)] == ra
    assert info.movi_offset == 0x138


def test_mux_golden_and_roundtrip(lib):
    """needs neither the reference tree nor _ref: the committed header/hash of a reference-muxed file"""
    g = np.load(GOLD)
    w, h, fps, n = g["dims"].tolist()
    data = amv.file_mux(w, h, fps, 22050, g["pk"], g["off"], g["sz"], g["ck"], g["coff"], g["csz"], lib=lib)
    assert data[:0x13c] == g["header"].tobytes()
    assert hashlib.md5(data).hexdigest() == bytes(g["md5"]).decode()
    info, vo, vs, ao, as_ = amv.file_index(data, lib=lib)
    assert (info.nvideo, info.naudio) == (n, n)
    assert np.array_equal(vs, g["sz"]) and np.array_equal(as_, g["csz"])
    W, H, FPS, vids, auds = walk_amv(data)           # the compare_amv.c-style walker agrees
    assert (W, H, FPS, len(vids), len(auds)) == (w, h, fps, n, n)


def test_index_rejects_and_flags(lib):
    with pytest.raises(amv.AmvError):
        amv.file_index(b"RIFF\0\0\0\0AVI LIST", lib=lib)
    g = np.load(GOLD)
    w, h, fps, n = g["dims"].tolist()
    data = amv.file_mux(w, h, fps, 22050, g["pk"], g["off"], g["sz"], g["ck"], g["coff"], g["csz"], lib=lib)
    cut = data[: len(data) - 300]
    info, vo, vs, ao, as_ = amv.file_index(cut, lib=lib)
    assert info.truncated == 1 and info.has_end_marker == 0 and info.nvideo + info.naudio < 2 * n
