"""BASELINE.json's configurations at their FULL sizes, device resident, through size-independent properties:
every unit of the big batch is a copy of one of a few hundred base units (drawn by a seeded index), the base
units are checked against the oracle on the CPU, and the big batch must reproduce them unit by unit -- so every
one of the 100 000 frames / 1 000 000 chunks / 1 000 000 frames is compared, bit for bit, with what the oracle
says, at the cost of a few hundred oracle units.

  config 2   100 000 frames 320x240 encoded (and decoded back)
  config 3   1 000 000 ADPCM chunks of 1378 samples encoded and decoded
  config 4   1 000 000 frames 208x176 decoded and re-encoded (four slices of 250 000, as four GPUs would split it)
"""
import numpy as np
import pytest

import amv_codec_tools_b200 as amv
from oracle_lib import Oracle, chroma_dims, offsets_of, synth_frames, synth_pcm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = amv.AmvCuda(device=0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


def _base_frames(w, h, nbase, seed):
    """sinus frames, a few flat ones and a few noise ones (long codes, FF bytes)"""
    y, u, v = synth_frames(nbase, w, h, seed=seed, kind="sinus")
    ny, nu, nv = synth_frames(8, w, h, seed=seed + 1, kind="noise")
    fy, fu, fv = synth_frames(8, w, h, seed=seed + 2, kind="flat")
    y[-16:-8], u[-16:-8], v[-16:-8] = ny, nu, nv
    y[-8:], u[-8:], v[-8:] = fy, fu, fv
    return y, u, v


def _encode_device(ctx, torch, Y, U, V, w, h, pkt_cap):
    n = Y.shape[0]
    dev = Y.device
    cw, ch = chroma_dims(w, h)
    out = torch.empty(n * pkt_cap, dtype=torch.uint8, device=dev)
    off = torch.zeros(n, dtype=torch.int64, device=dev)
    size = torch.zeros(n, dtype=torch.int32, device=dev)
    st = torch.zeros(n, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ctx.encode_frames_raw(Y, U, V, w, cw, w * h, cw * ch, n, w, h, None, out, out.numel(), pkt_cap, amv.LAYOUT_PACKED, off, size, st,
                          amv.MEM_DEVICE)
    ctx.sync()
    return out, off, size, st


def _decode_device(ctx, torch, pk, off, size, n, w, h):
    dev = pk.device
    cw, ch = chroma_dims(w, h)
    Y = torch.empty((n, h, w), dtype=torch.uint8, device=dev)
    U = torch.empty((n, ch, cw), dtype=torch.uint8, device=dev)
    V = torch.empty((n, ch, cw), dtype=torch.uint8, device=dev)
    st = torch.zeros(n, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ctx.decode_frames_raw(pk, pk.numel(), off, size, n, w, h, Y, U, V, w, cw, w * h, cw * ch, st, amv.MEM_DEVICE)
    ctx.sync()
    return Y, U, V, st


def _expect_packets(base_pk, base_off, base_sz, idx):
    """the byte string a packed batch must be: base packets in idx order"""
    views = [base_pk[int(o): int(o) + int(z)] for o, z in zip(base_off, base_sz)]
    sizes = base_sz[idx].astype(np.int64)
    offs = np.zeros(len(idx), np.int64)
    offs[1:] = np.cumsum(sizes)[:-1]
    return np.concatenate([views[j] for j in idx]), offs, sizes


def test_config2_100k_frames_320x240(ctx, oracle):
    import torch
    w, h, n, nbase = 320, 240, 100000, 160
    dev = torch.device("cuda", 0)
    by, bu, bv = _base_frames(w, h, nbase, 201)
    bpk, boff, bsz = oracle.encode_frames(by, bu, bv, w, h, 2)                      # the oracle's packets of the base frames
    wy, wu, wv, wst = oracle.decode_frames(bpk, boff, bsz, w, h)
    assert (wst == 0).all()
    idx = np.random.default_rng(202).integers(0, nbase, n)
    tidx = torch.from_numpy(idx).to(dev)
    Y, U, V = (torch.from_numpy(a).to(dev)[tidx] for a in (by, bu, bv))             # 11.5 GB of planes
    out, off, size, st = _encode_device(ctx, torch, Y, U, V, w, h, (int(bsz.max()) + 4096 + 4095) // 4096 * 4096)
    assert int(st.abs().sum().item()) == 0
    sz = size.cpu().numpy().astype(np.int64)
    assert np.array_equal(sz, bsz[idx].astype(np.int64))
    total = int(sz.sum())
    assert np.array_equal(off.cpu().numpy(), offsets_of(sz).astype(np.int64))
    want, _, _ = _expect_packets(bpk, boff, bsz, idx)
    assert len(want) == total
    assert torch.equal(out[:total], torch.from_numpy(want).to(dev)), "packets differ from the oracle's"
    del want
    # ... and back: every decoded frame equals the oracle's decode of its base packet
    del Y, U, V
    DY, DU, DV, dst = _decode_device(ctx, torch, out[:total], off, size, n, w, h)
    assert int(dst.abs().sum().item()) == 0
    for got_t, want_np in ((DY, wy), (DU, wu), (DV, wv)):
        assert torch.equal(got_t, torch.from_numpy(want_np).to(dev)[tidx])
    del DY, DU, DV, out
    torch.cuda.empty_cache()


def test_config3_1M_adpcm_chunks(ctx, oracle):
    import torch
    n, ns, nbase = 1000000, 1378, 1024
    dev = torch.device("cuda", 0)
    kinds = ["tones", "noise", "square", "silence"]
    base = np.concatenate([synth_pcm(ns * (nbase // 4), seed=300 + i, kind=k) for i, k in enumerate(kinds)]).reshape(nbase, ns)
    bstep = np.random.default_rng(301).integers(0, 89, nbase).astype(np.int16)      # start states
    bns = np.full(nbase, ns, np.uint32)
    benc, beoff, besz, bso = oracle.adpcm_encode(base.reshape(-1), offsets_of(bns), bns, bstep)
    bdec, _, _ = oracle.adpcm_decode(benc, beoff, besz)
    csz = int(besz[0])
    benc, bdec = benc.reshape(nbase, csz), bdec.reshape(nbase, -1)
    idx = np.random.default_rng(302).integers(0, nbase, n)
    tidx = torch.from_numpy(idx).to(dev)
    pcm = torch.from_numpy(base).to(dev)[tidx].reshape(-1)                           # 2.76 GB
    pcm_off = torch.arange(n, dtype=torch.int64, device=dev) * ns
    nsm = torch.full((n,), ns, dtype=torch.int32, device=dev)
    step_in = torch.from_numpy(bstep).to(dev)[tidx].contiguous()
    step_out = torch.zeros(n, dtype=torch.int16, device=dev)
    out = torch.empty(n * csz, dtype=torch.uint8, device=dev)
    out_off = torch.arange(n, dtype=torch.int64, device=dev) * csz
    st = torch.zeros(n, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ctx.adpcm_enc_chunks_raw(pcm, pcm.numel(), pcm_off, nsm, step_in, step_out, n, out, out.numel(), out_off, st, amv.MEM_DEVICE)
    ctx.sync()
    assert int(st.abs().sum().item()) == 0
    assert torch.equal(out.view(n, csz), torch.from_numpy(benc).to(dev)[tidx])
    assert torch.equal(step_out, torch.from_numpy(bso.astype(np.int16)).to(dev)[tidx])
    nd = bdec.shape[1]
    dec = torch.empty(n * nd, dtype=torch.int16, device=dev)
    dec_off = torch.arange(n, dtype=torch.int64, device=dev) * nd
    csizes = torch.full((n,), csz, dtype=torch.int32, device=dev)
    st.zero_()
    torch.cuda.synchronize()
    ctx.adpcm_dec_chunks_raw(out, out.numel(), out_off, csizes, n, dec, dec.numel(), dec_off, st, amv.MEM_DEVICE)
    ctx.sync()
    assert int(st.abs().sum().item()) == 0
    assert torch.equal(dec.view(n, nd), torch.from_numpy(bdec).to(dev)[tidx])


def test_config4_1M_frames_208x176_decode_encode(ctx, oracle):
    import torch
    w, h, n_total, nbase, slices = 208, 176, 1000000, 128, 4
    dev = torch.device("cuda", 0)
    by, bu, bv = _base_frames(w, h, nbase, 401)
    bpk, boff, bsz = oracle.encode_frames(by, bu, bv, w, h, 2)                      # the clip's packets
    dy, du, dv, dst = oracle.decode_frames(bpk, boff, bsz, w, h)
    assert (dst == 0).all()
    rpk, roff, rsz = oracle.encode_frames(dy, du, dv, w, h, 2)                      # what the round trip must write
    n = n_total // slices
    for s in range(slices):                                                          # contiguous frame ranges, one per "GPU"
        idx = np.random.default_rng(410 + s).integers(0, nbase, n)
        src, offs, sizes = _expect_packets(bpk, boff, bsz, idx)
        pk = torch.from_numpy(src).to(dev)
        off = torch.from_numpy(offs).to(dev)
        size = torch.from_numpy(sizes.astype(np.int32)).to(dev)
        Y, U, V, st = _decode_device(ctx, torch, pk, off, size, n, w, h)
        assert int(st.abs().sum().item()) == 0
        out, eoff, esize, est = _encode_device(ctx, torch, Y, U, V, w, h, (int(rsz.max()) + 4096 + 4095) // 4096 * 4096)
        assert int(est.abs().sum().item()) == 0
        esz = esize.cpu().numpy().astype(np.int64)
        assert np.array_equal(esz, rsz[idx].astype(np.int64))
        want, _, _ = _expect_packets(rpk, roff, rsz, idx)
        assert torch.equal(out[: int(esz.sum())], torch.from_numpy(want).to(dev)), "slice %d: re-encoded packets differ from the oracle's" % s
        del pk, Y, U, V, out
        torch.cuda.empty_cache()


@pytest.mark.parametrize("n,lanes_log2,what", [
    (37888, 0, "1 184 warps: CTAs of 8 warps (one per SM)"),
    (90000, 0, "2 813 warps: CTAs of 11 warps (two per SM, 22 warps on the busiest)"),
    (8192, 3, "2 048 warps: CTAs of 7 warps, with the synchronisation pass"),
    (200000, 0, "6 250 warps: several waves, CTAs of 8 warps"),
])
def test_decode_every_cta_size_of_the_lean_passes(ctx, oracle, n, lanes_log2, what):
    """The lean token pass and the lean synchronisation pass run in CTAs of 7, 8 or 11 warps, chosen per launch from the warp
    count (pick_vlc_warps, csrc/amv_dec.cu).  Batches of small frames sized to hit each choice: every decoded frame must
    equal the oracle's decode of the base packet it copies."""
    import torch
    w, h, nbase = 64, 48, 96
    dev = torch.device("cuda", 0)
    by, bu, bv = _base_frames(w, h, nbase, 401)
    bpk, boff, bsz = oracle.encode_frames(by, bu, bv, w, h, 2)
    wy, wu, wv, wst = oracle.decode_frames(bpk, boff, bsz, w, h)
    assert (wst == 0).all()
    idx = np.random.default_rng(402 + n).integers(0, nbase, n)
    tidx = torch.from_numpy(idx).to(dev)
    want, offs, sizes = _expect_packets(bpk, boff, bsz, idx)
    pk = torch.from_numpy(want).to(dev)
    off = torch.from_numpy(offs).to(dev)
    size = torch.from_numpy(sizes.astype(np.int32)).to(dev)
    ctx.set_option("decode_log2_lanes", lanes_log2)
    try:
        DY, DU, DV, dst = _decode_device(ctx, torch, pk, off, size, n, w, h)
    finally:
        ctx.set_option("decode_log2_lanes", -1)
    assert int(dst.abs().sum().item()) == 0, what
    for got_t, want_np in ((DY, wy), (DU, wu), (DV, wv)):
        assert torch.equal(got_t, torch.from_numpy(want_np).to(dev)[tidx]), what
    del DY, DU, DV, pk
    torch.cuda.empty_cache()
