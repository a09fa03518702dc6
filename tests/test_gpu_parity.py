"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle on the
same seeded inputs, against the committed golden vectors, and through size-independent
properties at larger sizes.  Bit-exact everywhere (integer / byte work)."""
import os

import numpy as np
import pytest

import amv_codec_tools_b200 as amv
from oracle_lib import Oracle, chroma_dims, offsets_of, pack, synth_frames, synth_pcm

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "amv_golden.npz"))
VIDEO_CASES = bytes(G["video_cases"]).decode().split("\n")


@pytest.fixture(scope="module")
def ctx():
    c = amv.AmvCuda(device=0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


# ------------------------------------------------------------------ ADPCM
@pytest.mark.parametrize("kind", ["tones", "noise", "square", "silence"])
@pytest.mark.parametrize("nchunks,ns", [(1, 1378), (33, 1378), (257, 1378), (64, 2), (40, 64), (17, 4000)])
def test_adpcm_roundtrip_vs_oracle(ctx, oracle, kind, nchunks, ns):
    pcm = synth_pcm(nchunks * ns, seed=41, kind=kind)
    nsamp = np.full(nchunks, ns, np.uint32)
    poff = offsets_of(nsamp)
    step_in = (np.arange(nchunks) * 7 % 89).astype(np.int16)
    out, ooff, osz, so, st = ctx.adpcm_encode(pcm, poff, nsamp, step_in)
    wout, _, wsz, wso = oracle.adpcm_encode(pcm, poff, nsamp, step_in)
    assert (st == 0).all()
    assert np.array_equal(osz, wsz) and np.array_equal(out, wout) and np.array_equal(so, wso)
    dec, _, dst = ctx.adpcm_decode(out, ooff, osz)
    wdec, _, wst = oracle.adpcm_decode(out, ooff, osz)
    assert (dst == 0).all() and np.array_equal(dec, wdec)


def test_adpcm_ragged_and_edge_cases(ctx, oracle):
    rng = np.random.default_rng(42)
    nsamp = (rng.integers(0, 900, 200) * 2).astype(np.uint32)      # includes empty chunks
    nsamp[:3] = [0, 2, 4]
    pcm = synth_pcm(int(nsamp.sum()) + 2, seed=43, kind="noise")
    poff = offsets_of(nsamp)
    out, ooff, osz, so, st = ctx.adpcm_encode(pcm, poff, nsamp)
    wout, _, wsz, wso = oracle.adpcm_encode(pcm, poff, nsamp, np.zeros(200, np.int16))
    assert (st == 0).all() and np.array_equal(out, wout) and np.array_equal(so, wso)
    dec, _, dst = ctx.adpcm_decode(out, ooff, osz)
    wdec, _, _ = oracle.adpcm_decode(out, ooff, osz)
    assert (dst == 0).all() and np.array_equal(dec, wdec)
    # arbitrary nibbles, every legal header step index, odd chunk sizes / unaligned offsets
    chunks = []
    for idx in range(89):
        body = rng.integers(0, 256, int(rng.integers(1, 200)), dtype=np.uint8).tobytes()
        pred = int(rng.integers(-32768, 32768)) & 0xFFFF
        chunks.append(pred.to_bytes(2, "little") + idx.to_bytes(2, "little") + (2 * len(body)).to_bytes(4, "little") + body)
    ck, coff, csz = pack(chunks)
    dec, _, dst = ctx.adpcm_decode(ck, coff, csz)
    wdec, _, _ = oracle.adpcm_decode(ck, coff, csz)
    assert (dst == 0).all() and np.array_equal(dec, wdec)
    # rejected inputs: short chunk, step index 89
    bad = [b"\x00\x00\x00", (0).to_bytes(2, "little") + (89).to_bytes(2, "little") + bytes(8)]
    ck, coff, csz = pack(bad)
    _, _, dst = ctx.adpcm_decode(ck, coff, csz)
    assert dst[0] == amv.ST_SHORT and dst[1] == amv.ST_RANGE


def test_adpcm_stream_chaining(ctx, oracle):
    """amv_adpcm_enc_streams carries the step index inside a stream like repeated adpcm_encode_frame calls."""
    k = "adpcm_tones/"
    for kind in ("tones", "noise", "square"):
        k = "adpcm_%s/" % kind
        cons = G[k + "cons"]
        poff = offsets_of(cons)
        # three copies of the golden stream as three independent streams
        nchunk = len(cons)
        src = np.concatenate([G[k + "src"]] * 3)
        pcm_off = np.concatenate([poff + i * len(G[k + "src"]) for i in range(3)]).astype(np.uint64)
        nsamp = np.concatenate([cons] * 3).astype(np.uint32)
        first = np.array([0, nchunk, 2 * nchunk, 3 * nchunk], np.uint32)
        out, ooff, osz, so, st = ctx.adpcm_encode_streams(src, pcm_off, nsamp, first)
        assert (st == 0).all()
        one = G[k + "out"]
        assert np.array_equal(out, np.concatenate([one] * 3))


@pytest.mark.parametrize("kind", ["tones", "noise", "square"])
def test_adpcm_golden(ctx, kind):
    k = "adpcm_%s/" % kind
    out, off, sz = G[k + "out"], G[k + "off"], G[k + "sz"]
    dec, _, st = ctx.adpcm_decode(out, off, sz)
    assert (st == 0).all() and np.array_equal(dec, G[k + "dec"])


# ------------------------------------------------------------------ video encode
@pytest.mark.parametrize("w,h", [(160, 120), (320, 240), (208, 176), (128, 96), (48, 40), (16, 16), (72, 24), (1280, 720)])
@pytest.mark.parametrize("kind", ["sinus", "noise", "flat", "edges"])
def test_encode_byte_identical(ctx, oracle, w, h, kind):
    n = 2 if w * h > 200000 else (5 if w * h > 40000 else 9)
    y, u, v = synth_frames(n, w, h, seed=51, kind=kind)
    pk, off, sz, st = ctx.encode_frames(y, u, v)
    wpk, woff, wsz = oracle.encode_frames(y, u, v, w, h, 2)
    assert (st == 0).all()
    assert np.array_equal(sz, wsz) and np.array_equal(off, woff)
    assert np.array_equal(pk, wpk)


@pytest.mark.parametrize("case", VIDEO_CASES)
def test_encode_golden(ctx, case):
    kind, dims, q = case.split("_")
    w, h = map(int, dims.split("x"))
    qs = ctx.qscale_from_quality(int(q[1:]))
    pk, off, sz, st = ctx.encode_frames(G[case + "/y"], G[case + "/u"], G[case + "/v"], qscale=qs)
    assert (st == 0).all() and np.array_equal(sz, G[case + "/sz"]) and np.array_equal(pk, G[case + "/pk"])


def test_encode_qscale_per_frame_and_layouts(ctx, oracle):
    w, h = 64, 48
    y, u, v = synth_frames(30, w, h, seed=52, kind="sinus")
    qs = (2 + np.arange(30) % 30).astype(np.int32)
    pk, off, sz, st = ctx.encode_frames(y, u, v, qscale=qs)
    assert (st == 0).all()
    for i in range(30):
        wpk, _, wsz = oracle.encode_frames(y[i:i + 1], u[i:i + 1], v[i:i + 1], w, h, int(qs[i]))
        assert np.array_equal(pk[int(off[i]): int(off[i]) + int(sz[i])], wpk)
    # slot layout: same packets at i*pkt_cap
    cap = 8192
    spk, soff, ssz, sst = ctx.encode_frames(y, u, v, qscale=qs, pkt_cap=cap, layout=amv.LAYOUT_SLOTS)
    assert np.array_equal(ssz, sz) and np.array_equal(soff, np.arange(30, dtype=np.uint64) * cap)
    for i in range(30):
        assert np.array_equal(spk[i * cap: i * cap + int(sz[i])], pk[int(off[i]): int(off[i]) + int(sz[i])])
    # capacity too small -> per-frame NOSPACE, size 0, neighbours unaffected
    tiny = int(sz.min()) + 1
    tpk, toff, tsz, tst = ctx.encode_frames(y, u, v, qscale=qs, pkt_cap=tiny, layout=amv.LAYOUT_SLOTS)
    for i in range(30):
        if sz[i] <= tiny:
            assert tst[i] == 0 and tsz[i] == sz[i]
        else:
            assert tst[i] == amv.ST_NOSPACE and tsz[i] == 0


def test_encode_rejects_unsupported(ctx):
    y, u, v = synth_frames(1, 32, 20, seed=1)            # (h/2)%8 == 2: the reference reads outside the picture
    with pytest.raises(amv.AmvError):
        ctx.encode_frames(y, u, v)
    y, u, v = synth_frames(1, 32, 32, seed=1)
    with pytest.raises(amv.AmvError):
        ctx.encode_frames(y, u, v, qscale=1)
    assert ctx.encode_frames(y[:0], u[:0], v[:0])[1].shape == (0,)      # empty batch


def test_encode_strided_device_buffers(ctx, oracle):
    """Device-resident input with padded rows / frame strides and an odd width (generic, non-8-byte path)."""
    import torch
    for (w, h, pad) in [(320, 240, 32), (72, 24, 3), (34, 16, 5)]:
        n = 4
        cw, ch = chroma_dims(w, h)
        y, u, v = synth_frames(n, w, h, seed=53, kind="noise")
        ls_y, ls_c = w + pad, cw + pad
        fs_y, fs_c = ls_y * h + 7 * pad, ls_c * ch + 3 * pad
        Y = np.zeros((n, fs_y), np.uint8); U = np.zeros((n, fs_c), np.uint8); V = np.zeros((n, fs_c), np.uint8)
        for i in range(n):
            Y[i, : ls_y * h].reshape(h, ls_y)[:, :w] = y[i]
            U[i, : ls_c * ch].reshape(ch, ls_c)[:, :cw] = u[i]
            V[i, : ls_c * ch].reshape(ch, ls_c)[:, :cw] = v[i]
        dY, dU, dV = (torch.from_numpy(a).cuda() for a in (Y, U, V))
        cap = w * h * 3 + 4096
        out = torch.zeros(n * cap, dtype=torch.uint8, device="cuda")
        off = torch.zeros(n, dtype=torch.int64, device="cuda")
        size = torch.zeros(n, dtype=torch.int32, device="cuda")
        st = torch.zeros(n, dtype=torch.int32, device="cuda")
        ctx.encode_frames_raw(dY, dU, dV, ls_y, ls_c, fs_y, fs_c, n, w, h, None, out, out.numel(), cap, amv.LAYOUT_PACKED,
                              off, size, st, amv.MEM_DEVICE)
        ctx.sync()
        wpk, woff, wsz = oracle.encode_frames(y, u, v, w, h, 2)
        assert (st.cpu().numpy() == 0).all()
        assert np.array_equal(size.cpu().numpy().astype(np.uint32), wsz)
        assert np.array_equal(out.cpu().numpy()[: len(wpk)], wpk)


# ------------------------------------------------------------------ video decode
@pytest.mark.parametrize("log2p", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("w,h,kind", [(160, 120, "sinus"), (320, 240, "sinus"), (208, 176, "noise"), (128, 96, "edges"),
                                      (48, 40, "flat"), (16, 16, "sinus"), (72, 24, "noise")])
def test_decode_identical_all_lane_counts(ctx, oracle, w, h, kind, log2p):
    n = 5 if w * h > 40000 else 11
    y, u, v = synth_frames(n, w, h, seed=61, kind=kind)
    pk, off, sz = oracle.encode_frames(y, u, v, w, h, 2)
    ctx.set_option("decode_log2_lanes", log2p)
    try:
        dy, du, dv, st = ctx.decode_frames(pk, off, sz, w, h)
    finally:
        ctx.set_option("decode_log2_lanes", -1)
    wy, wu, wv, wst = oracle.decode_frames(pk, off, sz, w, h)
    assert (st == 0).all() and (wst == 0).all()
    assert np.array_equal(dy, wy) and np.array_equal(du, wu) and np.array_equal(dv, wv)
    if log2p:
        assert 1 <= ctx.get_stat("decode_sync_rounds") <= (1 << log2p) + 1


@pytest.mark.parametrize("case", VIDEO_CASES)
def test_decode_golden(ctx, oracle, case):
    kind, dims, q = case.split("_")
    w, h = map(int, dims.split("x"))
    dy, du, dv, st = ctx.decode_frames(G[case + "/pk"], G[case + "/off"], G[case + "/sz"], w, h)
    assert (st == 0).all()
    # pixels where the reference itself indexes outside its clamp table are excluded
    _, _, _, _, masks = oracle.decode_frames(G[case + "/pk"], G[case + "/off"], G[case + "/sz"], w, h, undef=True)
    for got, want, m in zip((dy, du, dv), (G[case + "/dy"], G[case + "/du"], G[case + "/dv"]), masks):
        assert np.array_equal(got[m == 0], want[m == 0])


def test_decode_reference_fixture_head(ctx):
    """First packets of the reference's own AMV1.amv (real device stream) + its audio."""
    w, h, fps, n = G["AMV1/dims"].tolist()
    for log2p in (0, 5):
        ctx.set_option("decode_log2_lanes", log2p)
        dy, du, dv, st = ctx.decode_frames(G["AMV1/pk"], G["AMV1/off"], G["AMV1/sz"], w, h)
        ctx.set_option("decode_log2_lanes", -1)
        assert (st == 0).all()
        assert np.array_equal(dy, G["AMV1/dy"]) and np.array_equal(du, G["AMV1/du"]) and np.array_equal(dv, G["AMV1/dv"])
    pcm, _, ast = ctx.adpcm_decode(G["AMV1/ak"], G["AMV1/aoff"], G["AMV1/asz"])
    assert (ast == 0).all() and np.array_equal(pcm, G["AMV1/pcm"])


def test_decode_container_style_offsets_and_strides(ctx, oracle):
    """Packets at arbitrary (unaligned, gapped, unordered) offsets; padded device output planes."""
    import torch
    w, h, n = 208, 176, 6
    cw, ch = chroma_dims(w, h)
    y, u, v = synth_frames(n, w, h, seed=62, kind="sinus")
    pk, off, sz = oracle.encode_frames(y, u, v, w, h, 2)
    rng = np.random.default_rng(63)
    order = rng.permutation(n)
    blob = bytearray(rng.integers(0, 256, 13, dtype=np.uint8).tobytes())
    noff = np.zeros(n, np.uint64)
    for i in order:                                      # 00dc-style 8-byte headers between packets, no padding
        blob += b"00dc" + int(sz[i]).to_bytes(4, "little")
        noff[i] = len(blob)
        blob += pk[int(off[i]): int(off[i]) + int(sz[i])].tobytes()
    blob = np.frombuffer(bytes(blob), np.uint8)
    wy, wu, wv, _ = oracle.decode_frames(pk, off, sz, w, h)
    dy, du, dv, st = ctx.decode_frames(blob, noff, sz, w, h)
    assert (st == 0).all() and np.array_equal(dy, wy) and np.array_equal(du, wu) and np.array_equal(dv, wv)
    # padded device planes (generic store path: strides not multiples of 8)
    ls_y, ls_c = w + 3, cw + 5
    fs_y, fs_c = ls_y * h + 11, ls_c * ch + 1
    Y = torch.full((n * fs_y,), 7, dtype=torch.uint8, device="cuda")
    U = torch.full((n * fs_c,), 7, dtype=torch.uint8, device="cuda")
    V = torch.full((n * fs_c,), 7, dtype=torch.uint8, device="cuda")
    st = torch.zeros(n, dtype=torch.int32, device="cuda")
    dblob, doff, dsz = torch.from_numpy(blob.copy()).cuda(), torch.from_numpy(noff.astype(np.int64)).cuda(), \
        torch.from_numpy(sz.astype(np.int32)).cuda()
    ctx.decode_frames_raw(dblob, dblob.numel(), doff, dsz, n, w, h, Y, U, V, ls_y, ls_c, fs_y, fs_c, st, amv.MEM_DEVICE)
    ctx.sync()
    Yh = Y.cpu().numpy().reshape(n, fs_y)
    for i in range(n):
        assert np.array_equal(Yh[i, : ls_y * h].reshape(h, ls_y)[:, :w], wy[i])
        assert (Yh[i, : ls_y * h].reshape(h, ls_y)[:, w:] == 7).all()      # padding untouched
    Uh = U.cpu().numpy().reshape(n, fs_c)
    for i in range(n):
        assert np.array_equal(Uh[i, : ls_c * ch].reshape(ch, ls_c)[:, :cw], wu[i])


def test_decode_corrupt_streams_flagged_not_fatal(ctx, oracle):
    w, h = 160, 120
    y, u, v = synth_frames(4, w, h, seed=64, kind="sinus")
    pk, off, sz = oracle.encode_frames(y, u, v, w, h, 2)
    good = [pk[int(off[i]): int(off[i]) + int(sz[i])].tobytes() for i in range(4)]
    bad = [good[0], good[1][:300] + b"\xff\xd9", b"\xff\xd8\xff\xd9", b"\xff", good[2][:100] + b"\xff\xc4" + good[2][100:],
           good[3]]
    bk, boff, bsz = pack(bad)
    for log2p in (0, 3):
        ctx.set_option("decode_log2_lanes", log2p)
        dy, du, dv, st = ctx.decode_frames(bk, boff, bsz, w, h)
        ctx.set_option("decode_log2_lanes", -1)
        wy, wu, wv, _ = oracle.decode_frames(pk, off, sz, w, h)
        assert st[0] == 0 and st[5] == 0
        assert np.array_equal(dy[0], wy[0]) and np.array_equal(dy[5], wy[3])
        assert st[1] != 0 and st[2] != 0 and st[3] & amv.ST_SHORT and st[4] & amv.ST_MARKER
    # out-of-range offsets are rejected per frame
    st = ctx.decode_frames(pk, off + np.uint64(10 ** 9), sz, w, h)[3]
    assert (st == amv.ST_RANGE).all()


# ------------------------------------------------------------------ size-independent properties at scale
def test_roundtrip_properties_large_batch(ctx, oracle):
    """4096 frames of 320x240: (1) the packets of a frame do not depend on its position in the batch
    or the batch size; (2) decode(encode(x)) is identical whichever lane count decodes it and equals
    the oracle on an audited subset; (3) a checksum of checksums over the whole batch is stable."""
    import zlib
    w, h, n = 320, 240, 4096
    base_y, base_u, base_v = synth_frames(64, w, h, seed=71, kind="sinus")
    idx = np.random.default_rng(72).integers(0, 64, n)
    y, u, v = base_y[idx], base_u[idx], base_v[idx]
    pk, off, sz, st = ctx.encode_frames(y, u, v, pkt_cap=65536)
    assert (st == 0).all()
    bpk, boff, bsz, _ = ctx.encode_frames(base_y, base_u, base_v)
    wpk, woff, wsz = oracle.encode_frames(base_y[:8], base_u[:8], base_v[:8], w, h, 2)
    assert np.array_equal(bpk[: len(wpk)], wpk)
    crc_base = [zlib.crc32(bpk[int(boff[j]): int(boff[j]) + int(bsz[j])].tobytes()) for j in range(64)]
    for i in range(n):
        assert zlib.crc32(pk[int(off[i]): int(off[i]) + int(sz[i])].tobytes()) == crc_base[idx[i]]
    outs = []
    for log2p in (0, 2):
        ctx.set_option("decode_log2_lanes", log2p)
        dy, du, dv, dst = ctx.decode_frames(pk, off, sz, w, h)
        ctx.set_option("decode_log2_lanes", -1)
        assert (dst == 0).all()
        outs.append(zlib.crc32(dy.tobytes()) ^ zlib.crc32(du.tobytes()) ^ zlib.crc32(dv.tobytes()))
    assert outs[0] == outs[1]
    wy, wu, wv, _ = oracle.decode_frames(bpk, boff, bsz, w, h)
    for i in range(0, n, 37):
        assert np.array_equal(dy[i], wy[idx[i]]) and np.array_equal(du[i], wu[idx[i]]) and np.array_equal(dv[i], wv[idx[i]])


# ------------------------------------------------------------------ the drop-in boundary itself
DROPIN = os.path.join(amv.PKG_DIR, "glue", "_build", "dropin_check")


@pytest.mark.skipif(not os.path.exists(DROPIN), reason="glue/_build/dropin_check not built (needs the reference tree)")
@pytest.mark.parametrize("w,h,n,extra", [(160, 120, 24, []), (320, 240, 6, []), (208, 176, 5, []), (160, 120, 24, ["50", "7"]),
                                         (320, 240, 150, ["100", "64"])])
def test_avcodec_dropin_matches_reference_codecs(w, h, n, extra):
    """The reference's own libavcodec, driven like ffmpeg.c drives it, once with the libamvcuda AVCodec
    shims registered first (avcodec_find_* returns them) and once with its CPU codecs: identical
    packets, planes, ADPCM chunks and PCM (glue/ffmpeg/dropin_check.c).  The same frames also go through the
    look-ahead (CODEC_CAP_DELAY) shims -- extra = [timing effort %, queue depth]: depth 7 / 64 with more frames than
    the queue holds runs the steady state (a batch handed out while the next one fills), the default depth the drain."""
    import subprocess
    out = subprocess.run([DROPIN, str(w), str(h), str(n)] + extra, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "DROP-IN CHECK OK" in out.stdout and "drop-in frames/s" in out.stdout
    print([ln for ln in out.stdout.splitlines() if ln.startswith("drop-in frames/s")][0])


def test_host_path_pinned_zero_copy_and_chunked(ctx, oracle):
    """AMV_MEM_HOST with page-locked buffers: metadata, input packets and the packed output are read /
    written in place by the kernels (no staging copies), planes go through the chunked DMA ring.
    Small chunks force several ring turns; results must equal the pageable (staged) path and the oracle."""
    import torch
    w, h, n = 160, 120, 37
    cw, ch = chroma_dims(w, h)
    y, u, v = synth_frames(n, w, h, seed=91, kind="sinus")
    wpk, woff, wsz = oracle.encode_frames(y, u, v, w, h, 2)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    hy, hu, hv = pin(y), pin(u), pin(v)
    cap = n * 65536
    out = torch.zeros(cap, dtype=torch.uint8).pin_memory()
    off = torch.zeros(n, dtype=torch.int64).pin_memory()
    size = torch.zeros(n, dtype=torch.int32).pin_memory()
    st = torch.zeros(n, dtype=torch.int32).pin_memory()
    ctx.set_option("host_chunk_frames", 5)
    try:
        ctx.encode_frames_raw(hy, hu, hv, w, cw, w * h, cw * ch, n, w, h, None, out, cap, 65536, amv.LAYOUT_PACKED, off, size, st,
                              amv.MEM_HOST)
        assert (st.numpy() == 0).all()
        assert np.array_equal(size.numpy().astype(np.uint32), wsz) and np.array_equal(off.numpy().astype(np.uint64), woff)
        assert np.array_equal(out.numpy()[: len(wpk)], wpk)
        dy = torch.zeros((n, h, w), dtype=torch.uint8).pin_memory()
        du = torch.zeros((n, ch, cw), dtype=torch.uint8).pin_memory()
        dv = torch.zeros((n, ch, cw), dtype=torch.uint8).pin_memory()
        ctx.decode_frames_raw(out, cap, off, size, n, w, h, dy, du, dv, w, cw, w * h, cw * ch, st, amv.MEM_HOST)
        wy, wu, wv, _ = oracle.decode_frames(wpk, woff, wsz, w, h)
        assert (st.numpy() == 0).all()
        assert np.array_equal(dy.numpy(), wy) and np.array_equal(du.numpy(), wu) and np.array_equal(dv.numpy(), wv)
        # pageable inputs, same chunking: staged path
        pk2, off2, sz2, st2 = ctx.encode_frames(y, u, v)
        assert np.array_equal(pk2, wpk) and (st2 == 0).all()
        qy, qu, qv, qst = ctx.decode_frames(wpk, woff, wsz, w, h)
        assert np.array_equal(qy, wy) and np.array_equal(qu, wu) and np.array_equal(qv, wv)
    finally:
        ctx.set_option("host_chunk_frames", 0)


# ------------------------------------------------------------------ amvlib flavour
GA = np.load(os.path.join(os.path.dirname(__file__), "golden", "amvlib_golden.npz"))
AMVLIB_CASES = bytes(GA["video_cases"]).decode().split("\n")


@pytest.mark.parametrize("log2p", [0, 1, 3, 5])
@pytest.mark.parametrize("w,h,kind", [(160, 120, "sinus"), (320, 240, "sinus"), (208, 176, "noise"), (128, 96, "edges"),
                                      (48, 40, "flat"), (16, 16, "sinus"), (72, 24, "noise"), (102, 56, "sinus")])
def test_amvlib_decode_identical(ctx, oracle, w, h, kind, log2p):
    """amv_decode_frames_bgr24 (AmvVideoDecode mirror) vs the amvlib oracle: every byte of the bottom-up BGR24
    bitmaps, all lane counts, partial macroblocks, widths whose rows are padded (102*3 = 306 -> 308 bytes)"""
    n = 5 if w * h > 40000 else 11
    y, u, v = synth_frames(n, w, h, seed=71, kind=kind)
    pk, off, sz = oracle.encode_frames(y, u, v, w, h, 2 if kind == "sinus" else 7)
    ctx.set_option("decode_log2_lanes", log2p)
    try:
        bgr, st = ctx.decode_frames_bgr24(pk, off, sz, w, h)
    finally:
        ctx.set_option("decode_log2_lanes", -1)
    want, wst, um = oracle.amvlib_decode_frames(pk, off, sz, w, h, undef=True)
    assert (st == 0).all() and (wst == 0).all()
    # outside amvlib's clamp table the reference reads foreign memory; both sides saturate there, so the
    # comparison holds on every byte (the mask only says where the REFERENCE is undefined)
    assert np.array_equal(bgr, want)
    if kind == "sinus":
        assert not um.any()


@pytest.mark.parametrize("case", AMVLIB_CASES)
def test_amvlib_decode_golden(ctx, oracle, case):
    kind, dims, q = case.split("_")
    w, h = map(int, dims.split("x"))
    bgr, st = ctx.decode_frames_bgr24(GA[case + "/pk"], GA[case + "/off"], GA[case + "/sz"], w, h)
    assert (st == 0).all()
    _, _, um = oracle.amvlib_decode_frames(GA[case + "/pk"], GA[case + "/off"], GA[case + "/sz"], w, h, undef=True)
    assert np.array_equal(bgr[um == 0], GA[case + "/bgr"][um == 0])


def test_amvlib_fixture_head_and_audio(ctx):
    """First packets / chunks of the reference's AMV1.amv through the AmvVideoDecode / AmvAudioDecode mirrors.
    amvlib's audio is the ffmpeg decode of the chunk (one-byte step index) followed by the samples of its
    4-byte group padding: the chunk is handed over zero-padded to a multiple of four data bytes."""
    w, h, fps, n = GA["AMV1/dims"].tolist()
    bgr, st = ctx.decode_frames_bgr24(GA["AMV1/pk"], GA["AMV1/off"], GA["AMV1/sz"], w, h)
    assert (st == 0).all() and np.array_equal(bgr, GA["AMV1/bgr"])
    ak, aoff, asz = GA["AMV1/ak"], GA["AMV1/aoff"], GA["AMV1/asz"]
    padded = []
    for o, s in zip(aoff, asz):
        c = bytearray(ak[int(o):int(o) + int(s)].tobytes())
        c[3] = 0
        c += bytes((-(int(s) - 8)) % 4)
        padded.append(bytes(c))
    pk, po, ps = pack(padded)
    pcm, _, ast = ctx.adpcm_decode(pk, po, ps)
    assert (ast == 0).all() and np.array_equal(pcm, GA["AMV1/pcm"])


def test_amvlib_device_buffers_strides_and_corrupt(ctx, oracle):
    import torch
    w, h, n = 208, 176, 33
    y, u, v = synth_frames(n, w, h, seed=72, kind="sinus")
    pk, off, sz = oracle.encode_frames(y, u, v, w, h, 3)
    lb, fs = 3 * w + 40, (3 * w + 40) * h + 128            # padded rows and frames; padding must stay untouched
    dev = torch.device("cuda", 0)
    d_pk = torch.from_numpy(pk).to(dev)
    d_off = torch.from_numpy(off.astype(np.int64)).to(dev)
    d_sz = torch.from_numpy(sz.astype(np.int32)).to(dev)
    d_bgr = torch.full((n * fs,), 0xA5, dtype=torch.uint8, device=dev)
    d_st = torch.zeros(n, dtype=torch.int32, device=dev)
    ctx.decode_frames_bgr24_raw(d_pk, d_pk.numel(), d_off, d_sz, n, w, h, d_bgr, lb, fs, d_st, amv.MEM_DEVICE)
    ctx.sync()
    got = d_bgr.cpu().numpy().reshape(n, fs)
    want, wst = oracle.amvlib_decode_frames(pk, off, sz, w, h, line_bytes=lb)
    assert int(d_st.abs().sum()) == 0
    img = got[:, : lb * h].reshape(n, h, lb)
    assert np.array_equal(img[:, :, : 3 * w], want[:, :, : 3 * w])
    assert (img[:, :, 3 * w:] == 0xA5).all() and (got[:, lb * h:] == 0xA5).all()
    # corrupt packets are flagged, never fatal
    bad = pk.copy()
    bad[int(off[3]) + 40: int(off[3]) + 60] = 0xFF
    cut_sz = sz.copy()
    cut_sz[5] = 30
    bgr, st = ctx.decode_frames_bgr24(bad, off, cut_sz, w, h)
    assert st[3] != 0 and st[5] != 0 and (np.delete(st, [3, 5]) == 0).all()
    ok = [i for i in range(n) if i not in (3, 5)]
    assert np.array_equal(bgr[ok][:, :, : 3 * w], want[ok][:, :, : 3 * w])


AMVLIB_DROPIN = os.path.join(amv.PKG_DIR, "glue", "_build", "amvlib_dropin_check")


@pytest.mark.skipif(not os.path.exists(AMVLIB_DROPIN), reason="glue/_build/amvlib_dropin_check not built (needs the reference tree)")
def test_amvlib_dropin_matches_reference_amvlib(oracle, tmp_path):
    """The reference's own AmvVideoDecode / AmvAudioDecode (unmodified amvlib objects) next to the libamvcuda
    bindings of the same two entry points (glue/amvlib/amvcuda_amvlib.c), driven through an AMVDecoder like
    amvlib's callers do: identical bitmaps, PCM, lengths and return codes (glue/amvlib/amvlib_dropin_check.c)."""
    import struct
    import subprocess
    w, h, fps, n = GA["AMV1/dims"].tolist()
    units = [(w, h, [GA["AMV1/pk"][int(o):int(o) + int(s)].tobytes() for o, s in zip(GA["AMV1/off"], GA["AMV1/sz"])],
              [GA["AMV1/ak"][int(o):int(o) + int(s)].tobytes() for o, s in zip(GA["AMV1/aoff"], GA["AMV1/asz"])])]
    for (ww, hh) in ((160, 120), (320, 240), (208, 176)):
        y, u, v = synth_frames(6, ww, hh, seed=81, kind="sinus")
        pk, off, sz = oracle.encode_frames(y, u, v, ww, hh, 2)
        pcm = synth_pcm(1378 * 4, seed=82)
        ck, coff, csz, _ = oracle.adpcm_encode(pcm, np.arange(4, dtype=np.uint64) * 1378, np.full(4, 1378, np.uint32),
                                               np.zeros(4, np.int16))
        units.append((ww, hh, [pk[int(o):int(o) + int(s)].tobytes() for o, s in zip(off, sz)],
                      [ck[int(o):int(o) + int(s)].tobytes() for o, s in zip(coff, csz)]))
    for i, (ww, hh, vids, auds) in enumerate(units):
        path = tmp_path / ("units%d.bin" % i)
        with open(path, "wb") as f:
            f.write(b"AMVP" + struct.pack("<4i", ww, hh, len(vids), len(auds)))
            for b in vids + auds:
                f.write(struct.pack("<I", len(b)) + b)
        out = subprocess.run([AMVLIB_DROPIN, str(path)], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "AMVLIB DROP-IN CHECK OK" in out.stdout


# ------------------------------------------------------------------ SP5X
GS = np.load(os.path.join(os.path.dirname(__file__), "golden", "sp5x_golden.npz"))
SP5X_CASES = bytes(GS["cases"]).decode().split("\n")


@pytest.mark.parametrize("log2p", [0, 2, 5])
@pytest.mark.parametrize("w,h,kind", [(160, 120, "sinus"), (320, 240, "sinus"), (208, 176, "noise"), (128, 96, "edges"),
                                      (48, 40, "flat"), (16, 16, "sinus"), (72, 24, "noise"), (102, 56, "sinus")])
def test_sp5x_decode_identical(ctx, oracle, w, h, kind, log2p):
    """amv_decode_frames_sp5x vs the oracle: literal FF bytes in the scan (no un-stuffing), 14-byte header skipped,
    top-down placement, partial macroblocks cropped; packets at unaligned offsets"""
    from oracle_lib import sp5x_from_amv
    n = 5 if w * h > 40000 else 11
    y, u, v = synth_frames(n, w, h, seed=91, kind=kind)
    pk, off, sz = oracle.encode_frames(y, u, v, w, h, 2 if kind == "sinus" else 6)
    sp, soff, ssz = sp5x_from_amv(oracle, pk, off, sz)
    sp = np.concatenate([np.zeros(3, np.uint8), sp])           # shift every packet off its natural alignment
    soff = soff + 3
    ctx.set_option("decode_log2_lanes", log2p)
    try:
        dy, du, dv, st = ctx.decode_frames(sp, soff, ssz, w, h, sp5x=True)
    finally:
        ctx.set_option("decode_log2_lanes", -1)
    wy, wu, wv, wst = oracle.sp5x_decode_frames(sp, soff, ssz, w, h)
    assert (st == 0).all() and (wst == 0).all()
    assert np.array_equal(dy, wy) and np.array_equal(du, wu) and np.array_equal(dv, wv)


@pytest.mark.parametrize("case", SP5X_CASES)
def test_sp5x_decode_golden(ctx, oracle, case):
    kind, dims, q = case.split("_")
    w, h = map(int, dims.split("x"))
    dy, du, dv, st = ctx.decode_frames(GS[case + "/pk"], GS[case + "/off"], GS[case + "/sz"], w, h, sp5x=True)
    assert (st == 0).all()
    _, _, _, _, masks = oracle.sp5x_decode_frames(GS[case + "/pk"], GS[case + "/off"], GS[case + "/sz"], w, h, undef=True)
    for got, want, m in zip((dy, du, dv), (GS[case + "/dy"], GS[case + "/du"], GS[case + "/dv"]), masks):
        assert np.array_equal(got[m == 0], want[m == 0])


# ------------------------------------------------------------------ plain MJPEG
GMJ = np.load(os.path.join(os.path.dirname(__file__), "golden", "mjpeg_golden.npz"))
MJPEG_CASES = bytes(GMJ["cases"]).decode().split("\n")


@pytest.mark.parametrize("case", MJPEG_CASES)
@pytest.mark.parametrize("log2p", [0, 3])
def test_mjpeg_decode_golden(ctx, oracle, case, log2p):
    """amv_mjpeg_configure + amv_decode_frames_mjpeg vs what the reference's mjpeg_decoder made of the same frames
    (tables from the stream's own DQT / DHT segments, some with non-standard quantisers)"""
    kind, dims, q = case.split("_")
    w, h = map(int, dims.split("x"))
    pk, off, sz = GMJ[case + "/pk"], GMJ[case + "/off"], GMJ[case + "/sz"]
    assert ctx.mjpeg_configure(pk[int(off[0]): int(off[0]) + int(sz[0])]) == (w, h)
    ctx.set_option("decode_log2_lanes", log2p)
    try:
        dy, du, dv, st = ctx.decode_frames(pk, off, sz, w, h, mjpeg=True)
    finally:
        ctx.set_option("decode_log2_lanes", -1)
    assert (st == 0).all()
    _, _, _, _, masks = oracle.mjpeg_decode_frames(pk, off, sz, w, h, undef=True)
    for got, want, m in zip((dy, du, dv), (GMJ[case + "/dy"], GMJ[case + "/du"], GMJ[case + "/dv"]), masks):
        assert np.array_equal(got[m == 0], want[m == 0])


def test_mjpeg_device_buffers_mixed_headers_and_errors(ctx, oracle):
    """device-resident packets at unaligned offsets; a frame with another header is flagged, not decoded;
    unsupported headers are refused by amv_mjpeg_configure"""
    import torch
    a, b = "sinus_160x120_dx", "sinus_208x176_d3"
    w, h = 160, 120
    pk, off, sz = GMJ[a + "/pk"], GMJ[a + "/off"], GMJ[a + "/sz"]
    other = GMJ[b + "/pk"][: int(GMJ[b + "/sz"][0])]
    frames = [pk[int(o): int(o) + int(s)].tobytes() for o, s in zip(off, sz)]
    bk, boff, bsz = pack([b"\x00" * 5 + frames[0], other.tobytes(), frames[1], frames[0][:300]])
    boff[0] += 5; bsz[0] -= 5
    assert ctx.mjpeg_configure(frames[0]) == (w, h)
    n = 4
    cw, ch = chroma_dims(w, h)
    Y = torch.zeros((n, h, w), dtype=torch.uint8, device="cuda")
    U = torch.zeros((n, ch, cw), dtype=torch.uint8, device="cuda")
    V = torch.zeros((n, ch, cw), dtype=torch.uint8, device="cuda")
    st = torch.zeros(n, dtype=torch.int32, device="cuda")
    dk = torch.from_numpy(bk.copy()).cuda()
    doff, dsz = torch.from_numpy(boff.astype(np.int64)).cuda(), torch.from_numpy(bsz.astype(np.int32)).cuda()
    ctx.decode_frames_raw(dk, dk.numel(), doff, dsz, n, w, h, Y, U, V, w, cw, w * h, cw * ch, st, amv.MEM_DEVICE, mjpeg=True)
    ctx.sync()
    st = st.cpu().numpy()
    assert st[0] == 0 and st[2] == 0 and st[1] & amv.ST_HEADER and st[3] != 0
    assert np.array_equal(Y[0].cpu().numpy(), GMJ[a + "/dy"][0]) and np.array_equal(Y[2].cpu().numpy(), GMJ[a + "/dy"][1])
    assert np.array_equal(V[2].cpu().numpy(), GMJ[a + "/dv"][1])
    # wrong dimensions for the configured header; headers outside the contract
    with pytest.raises(amv.AmvError):
        ctx.decode_frames(bk, boff, bsz, w + 16, h, mjpeg=True)
    good = np.frombuffer(frames[0], np.uint8)
    j = frames[0].find(b"\xff\xc0")
    for pos, val in ((j + 4, 12), (j + 11, 0x41), (j + 1, 0xc2)):
        bad = good.copy()
        bad[pos] = val
        with pytest.raises(amv.AmvError):
            ctx.mjpeg_configure(bad)
    with pytest.raises(amv.AmvError):
        ctx.mjpeg_configure(good[:100])


@pytest.mark.parametrize("w,h,kind", [(320, 240, "sinus"), (208, 176, "noise"), (102, 56, "edges")])
def test_mjpeg_decode_vs_oracle_larger_batch(ctx, oracle, w, h, kind):
    """JPEG frames assembled from the oracle's AMV scans behind a reference-made header with the fixed AMV tables
    replaced by the stream's own: checks the table path against the oracle on more frames than the golden set holds"""
    case = "sinus_160x120_dx"
    hdr_src = GMJ[case + "/pk"][: int(GMJ[case + "/sz"][0])]
    ow, oh, start = oracle.mjpeg_header(hdr_src)
    hdr = hdr_src[:start].copy()
    j = hdr.tobytes().find(b"\xff\xc0")
    hdr[j + 5], hdr[j + 6], hdr[j + 7], hdr[j + 8] = h >> 8, h & 255, w >> 8, w & 255
    n = 24
    y, u, v = synth_frames(n, w, h, seed=95, kind=kind)
    pk, off, sz = oracle.encode_frames(y, u, v, w, h, 3)
    frames = [hdr.tobytes() + pk[int(o) + 2: int(o) + int(s)].tobytes() for o, s in zip(off, sz)]   # header | scan | EOI
    mk, moff, msz = pack(frames)
    assert ctx.mjpeg_configure(frames[0]) == (w, h)
    dy, du, dv, st = ctx.decode_frames(mk, moff, msz, w, h, mjpeg=True)
    wy, wu, wv, wst, masks = oracle.mjpeg_decode_frames(mk, moff, msz, w, h, undef=True)
    assert (st == 0).all() and (wst == 0).all()
    for got, want, m in zip((dy, du, dv), (wy, wu, wv), masks):
        assert np.array_equal(got[m == 0], want[m == 0])


# ------------------------------------------------------------------ range conversion
@pytest.mark.parametrize("direction", [0, 1])
@pytest.mark.parametrize("w,h", [(320, 240), (208, 176), (102, 56), (16, 16)])
def test_range_conversion_identical(ctx, oracle, direction, w, h):
    import torch
    rng = np.random.default_rng(12)
    cw, ch = chroma_dims(w, h)
    n = 5
    y = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
    u = rng.integers(0, 256, (n, ch, cw), dtype=np.uint8)
    v = rng.integers(0, 256, (n, ch, cw), dtype=np.uint8)
    y[0].reshape(-1)[:256] = np.arange(256)
    want = oracle.convert_range(y, u, v, direction)
    got = ctx.convert_range(y, u, v, direction)                       # host buffers
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    # device buffers, padded rows, in place
    dev = torch.device("cuda", 0)
    ls_y, ls_c = w + 16 - w % 16 + 16, cw + 16 - cw % 16 + 16
    dY = torch.full((n, h, ls_y), 7, dtype=torch.uint8, device=dev); dY[:, :, :w] = torch.from_numpy(y).to(dev)
    dU = torch.full((n, ch, ls_c), 7, dtype=torch.uint8, device=dev); dU[:, :, :cw] = torch.from_numpy(u).to(dev)
    dV = torch.full((n, ch, ls_c), 7, dtype=torch.uint8, device=dev); dV[:, :, :cw] = torch.from_numpy(v).to(dev)
    torch.cuda.synchronize()
    ctx.convert_range_raw(dY, dU, dV, ls_y, ls_c, h * ls_y, ch * ls_c, n, w, h, direction, dY, dU, dV, ls_y, ls_c, h * ls_y,
                          ch * ls_c, amv.MEM_DEVICE)
    ctx.sync()
    for t, ww, wn in ((dY, w, want[0]), (dU, cw, want[1]), (dV, cw, want[2])):
        a = t.cpu().numpy()
        assert np.array_equal(a[:, :, :ww], wn) and (a[:, :, ww:] == 7).all()


# ------------------------------------------------------------------ ADPCM -trellis N
@pytest.mark.parametrize("trellis", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("kind", ["tones", "noise", "square"])
def test_adpcm_trellis_vs_oracle(ctx, oracle, kind, trellis):
    """option adpcm_trellis = N: independent chunks (ragged lengths, start states) and chained streams"""
    rng = np.random.default_rng(41)
    ns = np.array([1378, 1378, 2, 0, 130, 256, 258, 4000, 64, 1378] * 4, np.uint32)
    pcm = synth_pcm(int(ns.sum()), seed=42, kind=kind)
    poff = offsets_of(ns)
    step_in = rng.integers(0, 89, len(ns)).astype(np.int16)
    ctx.set_option("adpcm_trellis", trellis)
    try:
        eo, eoff, esz, so, st = ctx.adpcm_encode(pcm, poff, ns, step_in)
        first = np.array([0, 7, 8, 20, len(ns)], np.uint32)
        co, coff, csz, cso, cst = ctx.adpcm_encode_streams(pcm, poff, ns, first, step_in[:4])
    finally:
        ctx.set_option("adpcm_trellis", 0)
    wo, woff, wsz, wso = oracle.adpcm_encode_trellis(pcm, poff, ns, step_in, trellis)
    assert (st == 0).all() and np.array_equal(esz, wsz) and np.array_equal(eo, wo) and np.array_equal(so, wso)
    # chained: the state after a chunk feeds the next chunk of its stream
    want = np.zeros_like(co)
    for sidx in range(4):
        state = int(step_in[sidx])
        for c in range(int(first[sidx]), int(first[sidx + 1])):
            o1, _, _, s1 = oracle.adpcm_encode_trellis(pcm, poff[c:c + 1], ns[c:c + 1], np.array([state], np.int16), trellis)
            want[int(coff[c]):int(coff[c]) + len(o1)] = o1
            state = int(s1[0])
        assert int(cso[sidx]) == state
    assert (cst == 0).all() and np.array_equal(co, want)


@pytest.mark.parametrize("case", sorted(set(k.split("/")[0] for k in np.load(os.path.join(os.path.dirname(__file__), "golden", "adpcm_trellis_golden.npz")).files)))
def test_adpcm_trellis_golden(ctx, case):
    GT = np.load(os.path.join(os.path.dirname(__file__), "golden", "adpcm_trellis_golden.npz"))
    trellis = int(case[1])
    out, off, sz, cons = (GT[case + "/" + k] for k in ("out", "off", "sz", "cons"))
    first = np.array([0, len(cons)], np.uint32)
    ctx.set_option("adpcm_trellis", trellis)
    try:
        eo, _, esz, _, st = ctx.adpcm_encode_streams(GT[case + "/src"], offsets_of(cons), cons, first, np.zeros(1, np.int16))
    finally:
        ctx.set_option("adpcm_trellis", 0)
    assert (st == 0).all() and np.array_equal(esz, sz) and np.array_equal(eo, out)


# ------------------------------------------------------------------ picture scaler / audio resampler
GR = np.load(os.path.join(os.path.dirname(__file__), "golden", "resample_golden.npz"))


@pytest.mark.parametrize("case", sorted(set(k.split("/")[0] for k in GR.files if k.startswith("scale_"))))
def test_scaler_matches_golden(ctx, case):
    """amv_scale_frames against pictures the reference's img_resample produced"""
    ow, oh = map(int, case.split("_")[2].split("x"))
    got = ctx.scale_frames(GR[case + "/y"], GR[case + "/u"], GR[case + "/v"], ow, oh, fill=7)
    for a, nm in zip(got, ("oy", "ou", "ov")):
        assert np.array_equal(a, GR[case + "/" + nm])


@pytest.mark.parametrize("dims", [(640, 480, 320, 240), (352, 288, 208, 176), (160, 120, 320, 240), (321, 243, 160, 120),
                                  (1280, 720, 128, 96), (100, 100, 101, 99), (720, 576, 208, 176), (64, 48, 640, 360),
                                  (16, 16, 2, 2), (5, 3, 17, 9), (320, 240, 320, 120), (2, 2, 8, 8), (8, 8, 1, 1)])
@pytest.mark.parametrize("form", [0, 1, 2])
def test_scaler_identical(ctx, oracle, dims, form):
    """host buffers and padded device planes; down, up, odd sizes (chroma at sizes >> 1), extreme ratios; every kernel
    form (option scale_form: direct, tiles, tiles with staged source rows)"""
    import torch
    iw, ih, ow, oh = dims
    ctx.set_option("scale_form", form)
    try:
        _scaler_identical(ctx, oracle, torch, iw, ih, ow, oh)
    finally:
        ctx.set_option("scale_form", 1)


def _scaler_identical(ctx, oracle, torch, iw, ih, ow, oh):
    rng = np.random.default_rng(iw * 7 + oh)
    n = 3
    icw, ich = chroma_dims(iw, ih)
    ocw, och = chroma_dims(ow, oh)
    y = rng.integers(0, 256, (n, ih, iw), dtype=np.uint8)
    u = rng.integers(0, 256, (n, ich, icw), dtype=np.uint8)
    v = rng.integers(0, 256, (n, ich, icw), dtype=np.uint8)
    y[1] = np.where(rng.random((ih, iw)) < 0.5, 0, 255)
    want = oracle.scale_frames(y, u, v, ow, oh, fill=9)
    got = ctx.scale_frames(y, u, v, ow, oh, fill=9)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    dev = torch.device("cuda", 0)
    for pad in (16, 3):                                   # 4-byte aligned rows (32-bit stores) and odd pitches (byte stores)
        ils_y, ils_c, ols_y, ols_c = iw + pad, icw + pad, ow + pad, ocw + pad
        dY = torch.full((n, ih, ils_y), 5, dtype=torch.uint8, device=dev); dY[:, :, :iw] = torch.from_numpy(y).to(dev)
        dU = torch.full((n, ich, ils_c), 5, dtype=torch.uint8, device=dev); dU[:, :, :icw] = torch.from_numpy(u).to(dev)
        dV = torch.full((n, ich, ils_c), 5, dtype=torch.uint8, device=dev); dV[:, :, :icw] = torch.from_numpy(v).to(dev)
        oY = torch.full((n, oh, ols_y), 9, dtype=torch.uint8, device=dev)
        oU = torch.full((n, och, ols_c), 9, dtype=torch.uint8, device=dev)
        oV = torch.full((n, och, ols_c), 9, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        ctx.scale_frames_raw(dY, dU, dV, ils_y, ils_c, ih * ils_y, ich * ils_c, n, iw, ih, oY, oU, oV, ols_y, ols_c, oh * ols_y,
                             och * ols_c, ow, oh, amv.MEM_DEVICE)
        ctx.sync()
        for t, ww, wn in ((oY, ow, want[0]), (oU, ocw, want[1]), (oV, ocw, want[2])):
            a = t.cpu().numpy()
            assert np.array_equal(a[:, :, :ww], wn) and (a[:, :, ww:] == 9).all()


@pytest.mark.parametrize("flags", [1, 2, 3])
@pytest.mark.parametrize("dims", [(352, 288, 208, 176), (160, 120, 320, 240)])
def test_scaler_with_range_steps(ctx, oracle, dims, flags):
    """amv_scale_frames_ex: the img_convert steps the fork's sws_scale puts around the scaler for YUVJ420P sides;
    host buffers and device planes (the source planes stay as they were)"""
    import torch
    iw, ih, ow, oh = dims
    rng = np.random.default_rng(ow + flags)
    n = 3
    y = rng.integers(0, 256, (n, ih, iw), dtype=np.uint8)
    u = rng.integers(0, 256, (n, ih // 2, iw // 2), dtype=np.uint8)
    v = rng.integers(0, 256, (n, ih // 2, iw // 2), dtype=np.uint8)
    want = oracle.sws_scale(y, u, v, ow, oh, flags & 1, flags & 2)
    got = ctx.scale_frames(y, u, v, ow, oh, flags=flags)
    assert all(np.array_equal(a, b) for a, b in zip(got, want))
    dev = torch.device("cuda", 0)
    dY, dU, dV = (torch.from_numpy(a).to(dev) for a in (y, u, v))
    oY = torch.zeros((n, oh, ow), dtype=torch.uint8, device=dev)
    oU = torch.zeros((n, oh // 2, ow // 2), dtype=torch.uint8, device=dev)
    oV = torch.zeros((n, oh // 2, ow // 2), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    ctx.scale_frames_raw(dY, dU, dV, iw, iw // 2, iw * ih, iw * ih // 4, n, iw, ih, oY, oU, oV, ow, ow // 2, ow * oh, ow * oh // 4,
                         ow, oh, amv.MEM_DEVICE, flags)
    ctx.sync()
    assert all(np.array_equal(t.cpu().numpy(), b) for t, b in zip((oY, oU, oV), want))
    assert all(np.array_equal(t.cpu().numpy(), b) for t, b in zip((dY, dU, dV), (y, u, v)))


def test_scaler_feeds_the_encoder(ctx, oracle):
    """`-s 208x176` in front of the AMV encoder: scaled frames encode to the packets the oracle makes of the oracle's scaling"""
    y, u, v = synth_frames(4, 352, 288, seed=5, kind="sinus")
    sy, su, sv = ctx.scale_frames(y, u, v, 208, 176)
    wy, wu, wv = oracle.scale_frames(y, u, v, 208, 176)
    assert all(np.array_equal(a, b) for a, b in zip((sy, su, sv), (wy, wu, wv)))
    pk, off, sz, st = ctx.encode_frames(sy, su, sv)
    wpk, woff, wsz = oracle.encode_frames(wy, wu, wv, 208, 176)
    assert (st == 0).all() and np.array_equal(sz, wsz) and np.array_equal(pk[: int(sz.sum())], wpk[: int(wsz.sum())])


@pytest.mark.parametrize("case", sorted(set(k.split("/")[0] for k in GR.files if k.startswith("audio_"))))
def test_audio_resampler_matches_golden(ctx, case):
    """amv_audio_resample against streams the reference's audio_resample produced packet by packet"""
    rate, ch = map(int, case.split("_")[1:])
    assert np.array_equal(ctx.audio_resample(GR[case + "/pcm"], ch, rate, 22050), GR[case + "/out"])


@pytest.mark.parametrize("rate,ch,n", [(44100, 2, 50000), (48000, 1, 70000), (8000, 1, 9000), (22050, 2, 30000), (11025, 1, 20000),
                                        (32000, 2, 3000000), (44100, 1, 1000), (48000, 2, 100), (96000, 2, 40000), (16000, 1, 20),
                                        (96000, 1, 50), (44100, 2, 1), (705600, 1, 300000), (705600, 2, 2000), (11025, 2, 30000), (47999, 1, 40000), (22049, 1, 30000),
                                        (24000, 2, 50000)])
@pytest.mark.parametrize("kind", ["noise", "tones", "square"])
@pytest.mark.parametrize("form", [0, 1, 2])
def test_audio_resampler_identical(ctx, oracle, rate, ch, n, kind, form):
    """every kernel form (option resample_form: direct, tiles, phase rows)"""
    ctx.set_option("resample_form", form)
    try:
        _audio_resampler_identical(ctx, oracle, rate, ch, n, kind)
    finally:
        ctx.set_option("resample_form", 2)


def _audio_resampler_identical(ctx, oracle, rate, ch, n, kind):
    """host and device buffers; streams shorter than the filter (mirrored taps only), beyond 2^21 samples (64-bit positions),
    full-scale square waves (32-bit accumulator, saturation), a 32:1 reduction (640 taps: the direct kernel form)"""
    import torch
    pcm = synth_pcm(n * ch, seed=rate + n, kind=kind)
    want = oracle.audio_resample(pcm, ch, rate, 22050)
    assert ctx.audio_resample_count(n, rate, 22050) == len(want)
    got = ctx.audio_resample(pcm, ch, rate, 22050)
    assert np.array_equal(got, want)
    dev = torch.device("cuda", 0)
    d_in = torch.from_numpy(pcm).to(dev)
    d_out = torch.full((len(want) + 8,), 77, dtype=torch.int16, device=dev)
    torch.cuda.synchronize()
    k = ctx.audio_resample_raw(d_in, n, ch, rate, 22050, d_out, len(want), amv.MEM_DEVICE)
    ctx.sync()
    a = d_out.cpu().numpy()
    assert k == len(want) and np.array_equal(a[:k], want) and (a[k:] == 77).all()


@pytest.mark.parametrize("rate,ch", [(44100, 2), (48000, 1), (8000, 1), (96000, 2)])
def test_audio_resampler_packet_feed(ctx, oracle, rate, ch):
    """amv_audio_resample_from over a sliding window (what the audio_resample shim of glue/ffmpeg does): ragged packets,
    some shorter than the filter, give the stream's outputs in order, each exactly once"""
    rng = np.random.default_rng(rate)
    pcm = synth_pcm(60000 * ch, seed=rate, kind="noise")
    cuts = np.sort(rng.choice(np.arange(1, 60000), 40, replace=False))
    # the first packet covers the mirrored taps of the stream's head (the reference mirrors modulo the first call's size)
    cuts = np.concatenate([[0, 100, 105, 109], cuts[cuts > 109], [60000]])
    cuts = np.unique(cuts)
    packets = [pcm[a * ch:b * ch] for a, b in zip(cuts[:-1], cuts[1:])]
    outs = ctx.audio_resample_packets(packets, ch, rate, 22050)
    assert np.array_equal(np.concatenate(outs), oracle.audio_resample(pcm, ch, rate, 22050))


@pytest.mark.parametrize("rate,ch,n,world", [(44100, 2, 200000, 4), (8000, 1, 50000, 8), (48000, 1, 3000, 3), (96000, 2, 60, 2)])
@pytest.mark.parametrize("form", [1, 2])
def test_audio_resampler_sharded_stream(ctx, oracle, rate, ch, n, world, form):
    ctx.set_option("resample_form", form)
    try:
        _audio_resampler_sharded_stream(ctx, oracle, rate, ch, n, world)
    finally:
        ctx.set_option("resample_form", 2)


def _audio_resampler_sharded_stream(ctx, oracle, rate, ch, n, world):
    """sharding.resample_shard: one stream split by output range as `world` GPUs would take it (here one after the
    other on one GPU, device buffers): the pieces concatenate to the stream's output"""
    import torch
    pcm = synth_pcm(n * ch, seed=n, kind="noise")
    want = oracle.audio_resample(pcm, ch, rate, 22050)
    dev = torch.device("cuda", 0)
    d_in = torch.from_numpy(pcm).to(dev)
    pieces = []
    for r in range(world):
        k0, kc, base, nwin = amv.sharding.resample_shard(n, rate, 22050, r, world, ctx.lib)
        assert k0 == sum(len(p) for p in pieces)
        d_out = torch.full((kc + 4,), 77, dtype=torch.int16, device=dev)
        torch.cuda.synchronize()
        k = ctx.audio_resample_from_raw(d_in[base * ch:(base + nwin) * ch], base, nwin, ch, rate, 22050, k0, d_out, kc, amv.MEM_DEVICE) if kc else 0
        ctx.sync()
        a = d_out.cpu().numpy()
        assert k == kc and (a[kc:] == 77).all()
        pieces.append(a[:kc])
    assert np.array_equal(np.concatenate(pieces), want)


def test_audio_resampler_feeds_the_adpcm_encoder(ctx, oracle):
    """44.1 kHz stereo -> 22050 Hz mono -> ADPCM chunks, as do_audio_out chains them"""
    pcm = synth_pcm(2 * 44100, seed=9, kind="tones")
    mono = ctx.audio_resample(pcm, 2, 44100, 22050)
    assert np.array_equal(mono, oracle.audio_resample(pcm, 2, 44100, 22050))
    nchunks = len(mono) // 1378
    ns = np.full(nchunks, 1378, np.uint32)
    off = offsets_of(ns)
    out, ooff, osz, so, st = ctx.adpcm_encode_streams(mono[: nchunks * 1378], off, ns, np.array([0, nchunks], np.uint32))
    assert (st == 0).all()
    # the chunks decode back to something close to the resampled stream (the codec is lossy; the parity of each stage
    # is pinned by its own tests)
    back = ctx.adpcm_decode(out, ooff, osz)[0]
    assert len(back) == nchunks * 1378


def test_resample_argument_errors(ctx):
    y = np.zeros((1, 16, 16), np.uint8); c = np.zeros((1, 8, 8), np.uint8)
    with pytest.raises(amv.AmvError):
        ctx.scale_frames(y, c, c, 0, 16)
    with pytest.raises(amv.AmvError):
        ctx.scale_frames(y, c, c, 20000, 16)
    with pytest.raises(amv.AmvError):                     # no chroma samples to scale from: outside the reference's defined domain
        ctx.scale_frames(y[:, :1, :1], c[:, :1, :1], c[:, :1, :1], 8, 8)
    with pytest.raises(amv.AmvError):
        ctx.audio_resample(np.zeros(30, np.int16), 3, 44100, 22050)
    with pytest.raises(amv.AmvError):
        ctx.audio_resample_raw(np.zeros(1000, np.int16), 1000, 1, 44100, 22050, np.zeros(10, np.int16), 10, amv.MEM_HOST)
    assert len(ctx.audio_resample(np.zeros(0, np.int16), 1, 44100, 22050)) == 0
