"""GPU parity on the BASELINE shapes round 1 left untested, against the oracle AND -- where oracle/_ref travelled to
the box -- against the unmodified reference itself, plus the regressions of the round-1 advisor findings.

  config 1   a 160x120 clip of 1 000 frames + 1 000 ADPCM chunks, encoded and muxed by the REFERENCE, indexed by
             amv_file_index and decoded from the file buffer; compared with the reference's own demux + decode
             (avidec.c:429-434 -> sp5xdec.c:33-93, adpcm.c:1268-1292)
  config 5   1280x720: every lane count on single frames, and 8 192 frames at once (encode and decode), every frame
             compared with the oracle by index replication like config 2
  Ref        encode / decode / ADPCM straight against oracle/_ref/libamvref.so (no restatement in between)
"""
import hashlib
import json
import os

import numpy as np
import pytest

import amv_codec_tools_b200 as amv
from oracle_lib import Oracle, Ref, chroma_dims, offsets_of, pack, synth_frames, synth_pcm

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
EMUL = os.environ.get("AMV_EMUL") == "1"


@pytest.fixture(scope="module")
def ctx():
    c = amv.AmvCuda(device=0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


needs_ref = pytest.mark.skipif(not Ref.available(), reason="oracle/_ref/libamvref.so not built on this box")


# ------------------------------------------------------------------ 1280x720 (config 5 shape)
@pytest.mark.parametrize("log2p", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("kind", ["sinus", "noise"])
def test_decode_1280x720_all_lane_counts(ctx, oracle, kind, log2p):
    w, h, n = 1280, 720, 2
    y, u, v = synth_frames(n, w, h, seed=71, kind=kind)
    pk, off, sz = oracle.encode_frames(y, u, v, w, h, 2)
    ctx.set_option("decode_log2_lanes", log2p)
    try:
        dy, du, dv, st = ctx.decode_frames(pk, off, sz, w, h)
    finally:
        ctx.set_option("decode_log2_lanes", -1)
    wy, wu, wv, wst = oracle.decode_frames(pk, off, sz, w, h)
    assert (st == 0).all() and (wst == 0).all()
    assert np.array_equal(dy, wy) and np.array_equal(du, wu) and np.array_equal(dv, wv)


@pytest.mark.parametrize("kind", ["sinus", "noise", "flat", "edges"])
def test_encode_1280x720_identical(ctx, oracle, kind):
    w, h, n = 1280, 720, 2
    y, u, v = synth_frames(n, w, h, seed=72, kind=kind)
    pk, off, sz, st = ctx.encode_frames(y, u, v)
    wpk, woff, wsz = oracle.encode_frames(y, u, v, w, h, 2)
    assert (st == 0).all() and np.array_equal(sz, wsz) and np.array_equal(pk, wpk)


@pytest.mark.skipif(EMUL, reason="device-resident full-size run")
def test_config5_8192_frames_1280x720(ctx, oracle):
    import torch
    w, h, n, nbase = 1280, 720, 8192, 24
    cw, ch = chroma_dims(w, h)
    dev = torch.device("cuda", 0)
    by, bu, bv = synth_frames(nbase, w, h, seed=501, kind="sinus")
    ny, nu, nv = synth_frames(2, w, h, seed=502, kind="noise")
    fy, fu, fv = synth_frames(2, w, h, seed=503, kind="flat")
    by[-4:-2], bu[-4:-2], bv[-4:-2] = ny, nu, nv
    by[-2:], bu[-2:], bv[-2:] = fy, fu, fv
    bpk, boff, bsz = oracle.encode_frames(by, bu, bv, w, h, 2)
    wy, wu, wv, wst = oracle.decode_frames(bpk, boff, bsz, w, h)
    assert (wst == 0).all()
    idx = np.random.default_rng(504).integers(0, nbase, n)
    tidx = torch.from_numpy(idx).to(dev)
    Y, U, V = (torch.from_numpy(a).to(dev)[tidx] for a in (by, bu, bv))             # 11.3 GB of planes
    pkt_cap = (int(bsz.max()) + 4096 + 4095) // 4096 * 4096
    out = torch.empty(int(bsz[idx].astype(np.int64).sum()) + 65536, dtype=torch.uint8, device=dev)
    off = torch.zeros(n, dtype=torch.int64, device=dev)
    size = torch.zeros(n, dtype=torch.int32, device=dev)
    st = torch.zeros(n, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ctx.encode_frames_raw(Y, U, V, w, cw, w * h, cw * ch, n, w, h, None, out, out.numel(), pkt_cap, amv.LAYOUT_PACKED, off, size, st,
                          amv.MEM_DEVICE)
    ctx.sync()
    assert int(st.abs().sum().item()) == 0
    sz = size.cpu().numpy().astype(np.int64)
    assert np.array_equal(sz, bsz[idx].astype(np.int64))
    assert np.array_equal(off.cpu().numpy(), offsets_of(sz).astype(np.int64))
    views = [bpk[int(o): int(o) + int(z)] for o, z in zip(boff, bsz)]
    want = np.concatenate([views[j] for j in idx])
    total = int(sz.sum())
    assert torch.equal(out[:total], torch.from_numpy(want).to(dev)), "1280x720 packets differ from the oracle's"
    del want, Y, U, V
    torch.cuda.empty_cache()
    DY = torch.empty((n, h, w), dtype=torch.uint8, device=dev)
    DU = torch.empty((n, ch, cw), dtype=torch.uint8, device=dev)
    DV = torch.empty((n, ch, cw), dtype=torch.uint8, device=dev)
    st.zero_()
    torch.cuda.synchronize()
    ctx.decode_frames_raw(out, total, off, size, n, w, h, DY, DU, DV, w, cw, w * h, cw * ch, st, amv.MEM_DEVICE)
    ctx.sync()
    assert int(st.abs().sum().item()) == 0
    for got_t, want_np in ((DY, wy), (DU, wu), (DV, wv)):
        assert torch.equal(got_t, torch.from_numpy(want_np).to(dev)[tidx])
    del DY, DU, DV, out
    torch.cuda.empty_cache()


# ------------------------------------------------------------------ straight against the reference
@needs_ref
@pytest.mark.parametrize("w,h,kind", [(320, 240, "sinus"), (208, 176, "noise"), (160, 120, "edges"), (1280, 720, "sinus")])
def test_encode_decode_vs_reference_itself(ctx, w, h, kind):
    ref = Ref()
    n = 2 if w > 1000 else 6
    y, u, v = synth_frames(n, w, h, seed=81, kind=kind)
    rpk, roff, rsz = ref.encode_frames(y, u, v, w, h, quality=0)                    # quality 0 -> qscale 2 (update_qscale)
    pk, off, sz, st = ctx.encode_frames(y, u, v)
    assert (st == 0).all() and np.array_equal(sz, rsz) and np.array_equal(pk, rpk), "packets differ from the reference encoder's"
    ry, ru, rv, got, _ = ref.decode_frames(rpk, roff, rsz, w, h)
    assert (got != 0).all()
    dy, du, dv, dst = ctx.decode_frames(rpk, roff, rsz, w, h)
    assert (dst == 0).all()
    # pixels where the reference itself indexes outside its clamp table (undefined in C, SURVEY 9.3: only noise gets there) are excluded
    masks = Oracle().decode_frames(rpk, roff, rsz, w, h, undef=True)[4]
    for got, want, m in zip((dy, du, dv), (ry, ru, rv), masks):
        assert np.array_equal(got[m == 0], want[m == 0]), "planes differ from the reference decoder's"


@needs_ref
def test_adpcm_vs_reference_itself(ctx):
    ref = Ref()
    ns, nchunks = 1378, 64
    pcm = synth_pcm(ns * nchunks, seed=82, kind="tones")
    rout, roff, rsz, cons = ref.adpcm_encode_stream(pcm, ns)                        # one chained stream, as ffmpeg.c feeds it
    k = len(rsz)
    nsamp = cons.astype(np.uint32)
    poff = offsets_of(nsamp)
    fc = np.array([0, k], np.uint32)
    out, ooff, osz, so, st = ctx.adpcm_encode_streams(pcm, poff, nsamp, fc, np.zeros(1, np.int16))
    assert (st == 0).all() and np.array_equal(osz, rsz) and np.array_equal(out, rout), "chunks differ from the reference encoder's"
    rdec, _, _ = ref.adpcm_decode(rout, roff, rsz)
    dec, _, dst = ctx.adpcm_decode(rout, roff, rsz)
    assert (dst == 0).all() and np.array_equal(dec, rdec), "PCM differs from the reference decoder's"


# ------------------------------------------------------------------ config 1 as written
def _config1_clip(oracle):
    """1 000 frames 160x120 + 22050 Hz audio at 16 fps, encoded and muxed by the reference when it is on the box, else by
    the oracle + our muxer (both pinned against the reference: tests/test_oracle_vs_ref.py, tests/test_container.py);
    the digest of the file the REFERENCE makes is committed (tests/golden/config1_digest.json, make_golden_config1.py)."""
    w, h, n, ns = 160, 120, 1000, 1378
    y, u, v = synth_frames(n, w, h, seed=1, kind="sinus")
    pcm = synth_pcm(ns * n + 4096, seed=1, kind="tones")
    if Ref.available():
        ref = Ref()
        vpk, voff, vsz = ref.encode_frames(y, u, v, w, h, quality=0)
        apk, aoff, asz, cons = ref.adpcm_encode_stream(pcm, ns, max_chunks=n)
        data = ref.mux(w, h, 16, 22050, vpk, voff, vsz, apk, aoff[:n], asz[:n])
        made_by = "reference"
    else:
        vpk, voff, vsz = oracle.encode_frames(y, u, v, w, h, 2)
        nsamp = np.full(n, ns, np.uint32)
        apk, aoff, asz, _ = oracle.adpcm_encode(pcm[: ns * n], offsets_of(nsamp), nsamp, np.zeros(n, np.int16))
        data = amv.file_mux(w, h, 16, 22050, vpk, voff, vsz, apk, aoff, asz)
        made_by = "oracle"
    return w, h, n, data, made_by


def test_config1_reference_clip_from_file_buffer(ctx, oracle):
    w, h, n, data, made_by = _config1_clip(oracle)
    digest = json.load(open(os.path.join(HERE, "golden", "config1_digest.json")))
    if made_by == "reference":
        assert hashlib.sha256(data).hexdigest() == digest["file_sha256"], "the reference-made clip differs from the committed digest"
    info, voff, vsz, aoff, asz = amv.file_index(data)
    assert (info.width, info.height, info.nvideo, info.naudio) == (w, h, n, n)
    buf = np.frombuffer(data, np.uint8)
    dy, du, dv, st = ctx.decode_frames(buf, voff, vsz, w, h)                        # packets addressed inside the file buffer
    assert (st == 0).all()
    dpcm, dpoff, ast = ctx.adpcm_decode(buf, aoff, asz)
    assert (ast == 0).all()
    if made_by == "reference":
        ref = Ref()
        rinfo, rv, ra = ref.demux(data)
        assert int(rinfo[4]) == n and int(rinfo[5]) == n
        rpk, roff, rsz = pack(rv)
        assert np.array_equal(rsz, vsz)
        ry, ru, rvv, got, _ = ref.decode_frames(rpk, roff, rsz, w, h)
        assert np.array_equal(dy, ry) and np.array_equal(du, ru) and np.array_equal(dv, rvv), "planes differ from the reference's decode of its own file"
        apk, ao, az = pack(ra)
        rpcm, _, _ = ref.adpcm_decode(apk, ao, az)
        assert np.array_equal(dpcm, rpcm), "PCM differs from the reference's decode of its own file"
        assert hashlib.sha256(ry.tobytes() + ru.tobytes() + rvv.tobytes()).hexdigest() == digest["planes_sha256"]
        assert hashlib.sha256(rpcm.tobytes()).hexdigest() == digest["pcm_sha256"]
    else:
        wy, wu, wv, wst = oracle.decode_frames(buf, voff, vsz, w, h)
        assert np.array_equal(dy, wy) and np.array_equal(du, wu) and np.array_equal(dv, wv)
        # same frames, same encoder arithmetic: the planes are the ones the reference decodes from its own clip
        assert hashlib.sha256(dy.tobytes() + du.tobytes() + dv.tobytes()).hexdigest() == digest["planes_sha256"]
        wpcm, _, _ = oracle.adpcm_decode(buf, aoff, asz)
        assert np.array_equal(dpcm, wpcm)


# ------------------------------------------------------------------ advisor findings of round 1
def test_adpcm_host_path_leaves_uncovered_bytes_alone(ctx, oracle):
    ns = 64
    pcm = synth_pcm(ns * 4, seed=91, kind="noise")
    nsamp = np.full(4, ns, np.uint32)
    out, ooff, osz, _ = oracle.adpcm_encode(pcm, offsets_of(nsamp), nsamp, np.zeros(4, np.int16))
    chunks = [out[int(o): int(o) + int(z)].tobytes() for o, z in zip(ooff, osz)]
    chunks[2] = chunks[2][:5]                                                       # rejected: shorter than its header
    ck, coff, csz = pack(chunks)
    dst = np.full(1000, 0x5a5a, np.int16)
    poff = np.array([10, 100, 300, 500], np.uint64)                                 # gaps between the chunks
    st = np.zeros(4, np.int32)
    r = ctx.lib.amv_adpcm_dec_chunks(ctx.ctx, ck.ctypes.data, ck.nbytes, coff.ctypes.data, csz.ctypes.data, 4, dst.ctypes.data,
                                     dst.size, poff.ctypes.data, st.ctypes.data, amv.MEM_HOST)
    assert r == 0 and st[2] & amv.ST_SHORT and st[0] == st[1] == st[3] == 0
    want, _, _ = oracle.adpcm_decode(out, ooff, osz)
    expect = np.full(1000, 0x5a5a, np.int16)
    for i in (0, 1, 3):
        expect[int(poff[i]): int(poff[i]) + ns] = want[i * ns: (i + 1) * ns]
    assert np.array_equal(dst, expect), "bytes outside the decoded chunks changed"


def test_packed_encode_without_room_reports_size_zero(ctx):
    y, u, v = synth_frames(4, 160, 120, seed=92, kind="sinus")
    cw, ch = chroma_dims(160, 120)
    out = np.zeros(9000, np.uint8)                                                  # room for one packet (~6 KB), not four
    off, size, st = np.zeros(4, np.uint64), np.zeros(4, np.uint32), np.zeros(4, np.int32)
    r = ctx.lib.amv_encode_frames(ctx.ctx, y.ctypes.data, u.ctypes.data, v.ctypes.data, 160, cw, 160 * 120, cw * ch, 4, 160, 120, None,
                                  out.ctypes.data, out.nbytes, 32768, amv.LAYOUT_PACKED, off.ctypes.data, size.ctypes.data,
                                  st.ctypes.data, amv.MEM_HOST)
    assert r == 0
    assert st[0] == 0 and size[0] > 0
    bad = st != 0
    assert bad.any() and (st[bad] & amv.ST_NOSPACE).all() and (size[bad] == 0).all()


def test_adpcm_stream_table_is_validated(ctx):
    pcm = synth_pcm(64 * 4, seed=93)
    nsamp = np.full(4, 64, np.uint32)
    poff = offsets_of(nsamp)
    out = np.zeros(4 * 40, np.uint8)
    ooff = (np.arange(4) * 40).astype(np.uint64)
    st = np.zeros(4, np.int32)
    so = np.zeros(2, np.int16)
    for fc in ([0, 3, 2], [0, 2, 9]):                                               # not monotonic / past nchunks
        fca = np.array(fc, np.uint32)
        r = ctx.lib.amv_adpcm_enc_streams(ctx.ctx, pcm.ctypes.data, pcm.size, poff.ctypes.data, nsamp.ctypes.data, fca.ctypes.data, 2, 4,
                                          None, so.ctypes.data, out.ctypes.data, out.nbytes, ooff.ctypes.data, st.ctypes.data, amv.MEM_HOST)
        assert r < 0


def test_hostile_offsets_are_range_errors_not_faults(ctx, oracle):
    y, u, v = synth_frames(2, 160, 120, seed=94)
    pk, off, sz = oracle.encode_frames(y, u, v, 160, 120, 2)
    bad = off.copy()
    bad[1] = np.uint64(2 ** 64 - 8)                                                 # off + size wraps
    dy, du, dv, st = ctx.decode_frames(pk, bad, sz, 160, 120)
    assert st[0] == 0 and st[1] == amv.ST_RANGE
    wy, _, _, _ = oracle.decode_frames(pk, off, sz, 160, 120)
    assert np.array_equal(dy[0], wy[0])
    # ... and the context is still alive
    dy2, _, _, st2 = ctx.decode_frames(pk, off, sz, 160, 120)
    assert (st2 == 0).all() and np.array_equal(dy2, wy)


@pytest.mark.skipif(EMUL, reason="needs two CUDA devices")
def test_two_devices_in_one_process():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("one device on this box")
    oracle = Oracle()
    y, u, v = synth_frames(3, 160, 120, seed=95)
    wpk, woff, wsz = oracle.encode_frames(y, u, v, 160, 120, 2)
    wy, wu, wv, _ = oracle.decode_frames(wpk, woff, wsz, 160, 120)
    torch.cuda.set_device(0)
    ctxs = [amv.AmvCuda(device=d) for d in (0, 1)]
    for c in ctxs:          # the second device needs its own tables and shared-memory opt-in
        pk, off, sz, st = c.encode_frames(y, u, v)
        assert (st == 0).all() and np.array_equal(pk, wpk)
        dy, du, dv, dst = c.decode_frames(pk, off, sz, 160, 120)
        assert (dst == 0).all() and np.array_equal(dy, wy) and np.array_equal(du, wu) and np.array_equal(dv, wv)
        assert torch.cuda.current_device() == 0, "an entry point left the caller on another device"
    for c in ctxs:
        c.close()
