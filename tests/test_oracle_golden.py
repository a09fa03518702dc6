"""The oracle against the committed golden vectors (tests/golden/amv_golden.npz,
made by tests/golden/make_golden.py from the unmodified reference).  Needs neither
/root/reference nor oracle/_ref: this is the pin that travels."""
import os

import numpy as np
import pytest

from oracle_lib import Oracle, offsets_of

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "amv_golden.npz"))
VIDEO_CASES = bytes(G["video_cases"]).decode().split("\n")


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


def test_tables(oracle):
    zz = oracle.zigzag()
    assert sorted(zz.tolist()) == list(range(64))
    assert zz[:10].tolist() == [0, 1, 8, 16, 9, 2, 3, 10, 17, 24] and zz[63] == 63
    for t, nsym in ((0, 12), (1, 12), (2, 162), (3, 162)):
        ln, cd = oracle.huff(t)
        assert (ln > 0).sum() == nsym
        # prefix-free and Kraft-complete except for the reserved all-ones code
        kraft = sum(2.0 ** -int(l) for l in ln if l)
        assert kraft < 1.0 and kraft + 2.0 ** -int(ln.max()) == 1.0
    ln, cd = oracle.huff(2)
    assert (ln[0x00], cd[0x00]) == (4, 0b1010) and (ln[0xF0], cd[0xF0]) == (11, 0b11111111001)
    ln, cd = oracle.huff(3)
    assert (ln[0x00], cd[0x00]) == (2, 0b00) and (ln[0xF0], cd[0xF0]) == (10, 0b1111111010)
    q = oracle.enc_qmat(2)      # qmat[0..2] = 65536 131072 131072
    assert q[:3].tolist() == [65536, 131072, 131072]


@pytest.mark.parametrize("case", VIDEO_CASES)
def test_encode_matches_golden(oracle, case):
    kind, dims, q = case.split("_")
    w, h = map(int, dims.split("x"))
    qscale = oracle.qscale_from_lambda(int(q[1:]))
    pk, off, sz = oracle.encode_frames(G[case + "/y"], G[case + "/u"], G[case + "/v"], w, h, qscale)
    assert np.array_equal(sz, G[case + "/sz"]) and np.array_equal(pk, G[case + "/pk"])


@pytest.mark.parametrize("case", VIDEO_CASES)
def test_decode_matches_golden(oracle, case):
    kind, dims, q = case.split("_")
    w, h = map(int, dims.split("x"))
    y, u, v, st, masks = oracle.decode_frames(G[case + "/pk"], G[case + "/off"], G[case + "/sz"], w, h, undef=True)
    assert (st == 0).all()
    for got, want, m in zip((y, u, v), (G[case + "/dy"], G[case + "/du"], G[case + "/dv"]), masks):
        assert np.array_equal(got[m == 0], want[m == 0])
        if kind != "noise":
            assert not m.any()


def test_decode_reference_fixture_head(oracle):
    w, h, fps, n = G["AMV1/dims"].tolist()
    y, u, v, st = oracle.decode_frames(G["AMV1/pk"], G["AMV1/off"], G["AMV1/sz"], w, h)
    assert (st == 0).all()
    assert np.array_equal(y, G["AMV1/dy"]) and np.array_equal(u, G["AMV1/du"]) and np.array_equal(v, G["AMV1/dv"])
    pcm, _, ast = oracle.adpcm_decode(G["AMV1/ak"], G["AMV1/aoff"], G["AMV1/asz"])
    assert (ast == 0).all() and np.array_equal(pcm, G["AMV1/pcm"])


@pytest.mark.parametrize("kind", ["tones", "noise", "square"])
def test_adpcm_matches_golden(oracle, kind):
    k = "adpcm_%s/" % kind
    out, off, sz, cons = G[k + "out"], G[k + "off"], G[k + "sz"], G[k + "cons"]
    step_in = np.array([int(out[int(o) + 2]) | (int(out[int(o) + 3]) << 8) for o in off], np.int16)
    eo, eoff, esz, step_out = oracle.adpcm_encode(G[k + "src"], offsets_of(cons), cons, step_in)
    assert np.array_equal(eo, out) and np.array_equal(esz, sz)
    dp, _, st = oracle.adpcm_decode(out, off, sz)
    assert (st == 0).all() and np.array_equal(dp, G[k + "dec"])


def test_error_statuses(oracle):
    # truncated packet -> overrun / bad code flagged, never a crash
    case = VIDEO_CASES[0]
    w, h = 160, 120
    pk = G[case + "/pk"][: int(G[case + "/sz"][0])]
    cut = np.concatenate([pk[:200], np.array([0xFF, 0xD9], np.uint8)])
    _, _, _, st = oracle.decode_frames(cut, np.array([0], np.uint64), np.array([len(cut)], np.uint32), w, h)
    assert st[0] != 0
    # ADPCM: short chunk and out-of-range step index are rejected
    bad = np.array([0, 0, 89, 0, 0, 0, 0, 0, 0x11], np.uint8)
    _, _, st = oracle.adpcm_decode(bad, np.array([0], np.uint64), np.array([9], np.uint32))
    assert st[0] < 0


# ------------------------------------------------------------------ amvlib flavour
GA = np.load(os.path.join(os.path.dirname(__file__), "golden", "amvlib_golden.npz"))
AMVLIB_CASES = bytes(GA["video_cases"]).decode().split("\n")


def test_amvlib_zigzag_typo(oracle):
    z = oracle.amvlib_zigzag()                       # raster -> zigzag index (AmvJpeg.c:131-141)
    assert z[3 * 8 + 4] == 37 and z[6 * 8 + 2] == 37 and 31 not in z.tolist()
    std = np.zeros(64, np.uint8)
    std[oracle.zigzag()] = np.arange(64)
    assert (z != std).sum() == 1


@pytest.mark.parametrize("case", AMVLIB_CASES)
def test_amvlib_decode_matches_golden(oracle, case):
    kind, dims, q = case.split("_")
    w, h = map(int, dims.split("x"))
    bgr, st, um = oracle.amvlib_decode_frames(GA[case + "/pk"], GA[case + "/off"], GA[case + "/sz"], w, h, undef=True)
    assert (st == 0).all()
    assert np.array_equal(bgr[um == 0], GA[case + "/bgr"][um == 0])
    if kind == "sinus":
        assert not um.any()


def test_amvlib_fixture_head(oracle):
    w, h, fps, n = GA["AMV1/dims"].tolist()
    bgr, st = oracle.amvlib_decode_frames(GA["AMV1/pk"], GA["AMV1/off"], GA["AMV1/sz"], w, h)
    assert (st == 0).all() and np.array_equal(bgr, GA["AMV1/bgr"])
    pcm, _, ns, ast = oracle.amvlib_audio_decode(GA["AMV1/ak"], GA["AMV1/aoff"], GA["AMV1/asz"])
    assert (ast == 0).all() and np.array_equal(ns, GA["AMV1/nsamp"]) and np.array_equal(pcm, GA["AMV1/pcm"])


# ------------------------------------------------------------------ SP5X
GS = np.load(os.path.join(os.path.dirname(__file__), "golden", "sp5x_golden.npz"))
SP5X_CASES = bytes(GS["cases"]).decode().split("\n")


@pytest.mark.parametrize("case", SP5X_CASES)
def test_sp5x_decode_matches_golden(oracle, case):
    kind, dims, q = case.split("_")
    w, h = map(int, dims.split("x"))
    y, u, v, st, masks = oracle.sp5x_decode_frames(GS[case + "/pk"], GS[case + "/off"], GS[case + "/sz"], w, h, undef=True)
    assert (st == 0).all()
    for got, want, m in zip((y, u, v), (GS[case + "/dy"], GS[case + "/du"], GS[case + "/dv"]), masks):
        assert np.array_equal(got[m == 0], want[m == 0])


# ------------------------------------------------------------------ plain MJPEG
GM = np.load(os.path.join(os.path.dirname(__file__), "golden", "mjpeg_golden.npz"))
MJPEG_CASES = bytes(GM["cases"]).decode().split("\n")


@pytest.mark.parametrize("case", MJPEG_CASES)
def test_mjpeg_decode_matches_golden(oracle, case):
    kind, dims, q = case.split("_")
    w, h = map(int, dims.split("x"))
    y, u, v, st, masks = oracle.mjpeg_decode_frames(GM[case + "/pk"], GM[case + "/off"], GM[case + "/sz"], w, h, undef=True)
    assert (st == 0).all()
    for got, want, m in zip((y, u, v), (GM[case + "/dy"], GM[case + "/du"], GM[case + "/dv"]), masks):
        assert np.array_equal(got[m == 0], want[m == 0])


def test_range_conversion_tables(oracle):
    """known answers of the four tables (colorspace.h:69-84 with SCALEBITS 10): ends, mid points, clamps"""
    v = np.arange(256, dtype=np.uint8).reshape(1, 16, 16)
    c = np.concatenate([np.arange(0, 256, 4), np.arange(3, 256, 4)]).astype(np.uint8).reshape(1, 8, 16)
    jy, ju, _ = oracle.convert_range(v, c, c, 0)
    cy, cu, _ = oracle.convert_range(v, c, c, 1)
    jy, cy = jy.reshape(-1), cy.reshape(-1)
    assert jy[16] == 0 and jy[235] == 255 and jy[0] == 0 and jy[255] == 255 and jy[126] == 128
    assert cy[0] == 16 and cy[255] == 235 and cy[128] == 126
    lut_ju = dict(zip(c.reshape(-1).tolist(), ju.reshape(-1).tolist()))
    lut_cu = dict(zip(c.reshape(-1).tolist(), cu.reshape(-1).tolist()))
    assert lut_ju[128] == 128 and lut_ju[16] == 1 and lut_ju[240] == 255 and lut_ju[0] == 0
    assert lut_cu[128] == 128 and lut_cu[0] == 16 and lut_cu[255] == 240


# ------------------------------------------------------------------ ADPCM -trellis N
GT = np.load(os.path.join(os.path.dirname(__file__), "golden", "adpcm_trellis_golden.npz"))
TRELLIS_CASES = sorted(set(k.split("/")[0] for k in GT.files))


@pytest.mark.parametrize("case", TRELLIS_CASES)
def test_adpcm_trellis_matches_golden(oracle, case):
    trellis = int(case[1])
    out, off, sz, cons = (GT[case + "/" + k] for k in ("out", "off", "sz", "cons"))
    step_in = np.array([int(out[int(o) + 2]) | (int(out[int(o) + 3]) << 8) for o in off], np.int16)
    eo, _, esz, _ = oracle.adpcm_encode_trellis(GT[case + "/src"], offsets_of(cons), cons, step_in, trellis)
    assert np.array_equal(esz, sz) and np.array_equal(eo, out)


# ------------------------------------------------------------------ picture scaler / audio resampler
GR = np.load(os.path.join(os.path.dirname(__file__), "golden", "resample_golden.npz"))
SCALE_CASES = sorted(set(k.split("/")[0] for k in GR.files if k.startswith("scale_")))
AUDIO_CASES = sorted(set(k.split("/")[0] for k in GR.files if k.startswith("audio_")))


@pytest.mark.parametrize("case", SCALE_CASES)
def test_scaler_matches_golden(oracle, case):
    """img_resample of the reference on small pictures (down, up, odd sizes, one axis only)"""
    ow, oh = map(int, case.split("_")[2].split("x"))
    got = oracle.scale_frames(GR[case + "/y"], GR[case + "/u"], GR[case + "/v"], ow, oh, fill=7)
    for a, nm in zip(got, ("oy", "ou", "ov")):
        assert np.array_equal(a, GR[case + "/" + nm])


@pytest.mark.parametrize("case", AUDIO_CASES)
def test_audio_resampler_matches_golden(oracle, case):
    """audio_resample of the reference fed in packets against the oracle's one-shot closed form"""
    rate, ch = map(int, case.split("_")[1:])
    got = oracle.audio_resample(GR[case + "/pcm"], ch, rate, 22050)
    assert np.array_equal(got, GR[case + "/out"])
    assert np.array_equal(oracle.resample_bank(rate, 22050)[[0, 1, 511, 512, 1023]], GR[case + "/bank_rows"])


def test_host_filter_banks_match_golden_and_oracle(oracle):
    """libamvcuda's host-side filter design (amv_scale_banks / amv_audio_resample_bank; no device involved)"""
    import amv_codec_tools_b200 as amv
    lib = amv.load_library()
    for case in AUDIO_CASES:
        rate = int(case.split("_")[1])
        bank = amv.audio_resample_bank(rate, 22050, lib=lib)
        assert np.array_equal(bank, oracle.resample_bank(rate, 22050))
        assert np.array_equal(bank[[0, 1, 511, 512, 1023]], GR[case + "/bank_rows"])
    for case in SCALE_CASES:
        iw, ih = map(int, case.split("_")[1].split("x"))
        ow, oh = map(int, case.split("_")[2].split("x"))
        hb, vb, hi, vi = amv.scale_banks(iw, ih, ow, oh, lib=lib)
        ohb, ovb = oracle.scale_banks(iw, ih, ow, oh)
        assert np.array_equal(hb, ohb) and np.array_equal(vb, ovb)
        assert hi == iw * 65536 // ow and vi == ih * 65536 // oh
    # how many samples the reference returns over a stream: a closed form in the ABI, counted by the oracle
    for rate, n in ((44100, 5000), (48000, 777), (8000, 100), (22050, 1378), (32000, 31), (96000, 50)):
        pcm = np.zeros(n, np.int16)
        assert lib.amv_audio_resample_count(n, rate, 22050) == len(oracle.audio_resample(pcm, 1, rate, 22050))
