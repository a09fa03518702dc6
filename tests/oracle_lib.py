"""ctypes front-ends for the CPU checkers (TEST INFRASTRUCTURE ONLY).

* ``Oracle``  -> oracle/_build/libamvoracle.so  (our C restatement, oracle/amv_oracle.c)
* ``Ref``     -> oracle/_ref/libamvref.so       (the unmodified reference codecs, built in
                 place by oracle/build_ref.sh; may be absent on a box without the prebuilt file)

Plus the shared synthetic input generators and an AMV
container walker used to pull packets out of the in-tree fixture.
"""
import ctypes as C
import os
import struct
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libamvoracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libamvref.so")
AMVLIB_REF_SO = os.path.join(ORACLE_DIR, "_ref", "libamvlibref.so")
FIXTURE_AMV = "/root/reference/C-AMVDecoder/bin/AMV1.amv"


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def build_oracle():
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
        os.path.join(ORACLE_DIR, "amv_oracle.c")
    ):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "_build/libamvoracle.so"])
    return ORACLE_SO


def chroma_dims(w, h):
    return (w + 1) // 2, (h + 1) // 2


def offsets_of(sizes):
    sizes = np.asarray(sizes, dtype=np.uint64)
    off = np.zeros(len(sizes), dtype=np.uint64)
    if len(sizes) > 1:
        off[1:] = np.cumsum(sizes)[:-1]
    return off


def amvlib_line_bytes(w):
    """WIDTHBYTES(w*24) of AmvJpeg.c:420,1526: bytes per bitmap row"""
    return (w * 24 + 31) // 32 * 4


class _AmvlibOracleMixin:
    """amvlib flavour of the oracle"""

    def amvlib_zigzag(self):
        z = np.zeros(64, np.uint8)
        self.lib.amvo_amvlib_zigzag(_p(z))
        return z

    def amvlib_idct(self, blocks):
        b = np.ascontiguousarray(blocks, dtype=np.int32).reshape(-1, 64).copy()
        for i in range(b.shape[0]):
            self.lib.amvo_amvlib_idct(_p(b[i]))
        return b

    def amvlib_decode_frames(self, pkts, off, size, w, h, line_bytes=None, undef=False):
        """-> bgr (n, h, line_bytes), status[, undef mask]: the mask is 1 where an IDCT output left the
        reference's clamp table (-512..511), i.e. where amvlib itself reads foreign memory."""
        n = len(size)
        lb = line_bytes or amvlib_line_bytes(w)
        bgr = np.zeros((n, h, lb), np.uint8)
        um = np.zeros((n, h, lb), np.uint8) if undef else None
        st = np.zeros(n, np.int32)
        self.lib.amvo_amvlib_decode_frames(_p(np.ascontiguousarray(pkts, np.uint8)), _p(np.ascontiguousarray(off, np.uint64)),
                                           _p(np.ascontiguousarray(size, np.uint32)), n, w, h, _p(bgr), lb,
                                           C.c_uint64(h * lb), _p(st), _p(um))
        return (bgr, st, um) if undef else (bgr, st)

    def amvlib_frame_coefs(self, pkt, w, h):
        mbw, mbh = (w + 15) // 16, (h + 15) // 16
        coef = np.zeros((mbw * mbh * 6, 64), np.int32)
        lb = amvlib_line_bytes(w)
        bgr = np.zeros((h, lb), np.uint8)
        pk = np.ascontiguousarray(pkt, np.uint8)
        st = self.lib.amvo_amvlib_decode_frame(_p(pk), len(pk), w, h, _p(bgr), lb, _p(coef))
        return coef, bgr, st

    def amvlib_audio_decode(self, chunks, off, size):
        n = len(size)
        ns = (np.maximum(np.asarray(size, np.int64) - 8, 0) + 3) // 4 * 8
        poff = offsets_of(ns)
        pcm = np.zeros(int(ns.sum()), np.int16)
        st = np.zeros(n, np.int32)
        ck = np.ascontiguousarray(chunks, np.uint8)
        for i in range(n):
            r = self.lib.amvo_amvlib_audio_decode_chunk(_p(ck[int(off[i]):]), int(size[i]), _p(pcm[int(poff[i]):]))
            st[i] = 0 if r >= 0 else r
        return pcm, poff, ns.astype(np.uint32), st


class Oracle(_AmvlibOracleMixin):
    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        L.amvo_unstuff.restype = C.c_size_t

    # -- stages
    def zigzag(self):
        z = np.zeros(64, np.uint8)
        self.lib.amvo_get_zigzag(_p(z))
        return z

    def huff(self, t):
        ln = np.zeros(256, np.uint8)
        cd = np.zeros(256, np.uint16)
        self.lib.amvo_get_huff(t, _p(ln), _p(cd))
        return ln, cd

    def fdct(self, blocks):
        b = np.ascontiguousarray(blocks, dtype=np.int16).reshape(-1, 64).copy()
        for i in range(b.shape[0]):
            self.lib.amvo_fdct_islow(_p(b[i]))
        return b

    def idct_put(self, blocks):
        b = np.ascontiguousarray(blocks, dtype=np.int16).reshape(-1, 64)
        out = np.zeros((b.shape[0], 64), np.uint8)
        for i in range(b.shape[0]):
            self.lib.amvo_idct_put(_p(b[i]), _p(out[i]))
        return out

    def enc_qmat(self, qscale):
        q = np.zeros(64, np.int32)
        self.lib.amvo_enc_qmat(int(qscale), _p(q))
        return q

    def qscale_from_lambda(self, lam, qmin=2, qmax=31):
        return self.lib.amvo_qscale_from_lambda(int(lam), qmin, qmax)

    # -- video
    def encode_frames(self, y, u, v, w, h, qscale=2, cap=None):
        n = y.shape[0]
        cap = cap or n * (w * h * 3 + 4096)
        out = np.zeros(cap, np.uint8)
        off = np.zeros(n, np.uint64)
        size = np.zeros(n, np.uint32)
        r = self.lib.amvo_encode_frames(_p(y), _p(u), _p(v), n, w, h, int(qscale), _p(out), _p(off), _p(size),
                                        C.c_uint64(cap))
        if r != n:
            raise RuntimeError("oracle encode failed: %d" % r)
        return out[: int(size.sum())].copy(), off, size

    def decode_frames(self, pkts, off, size, w, h, undef=False):
        """-> y, u, v, status[, (uy, uu, uv)]: the undef planes are 1 where the pre-clamp value
        falls outside ff_cropTbl's domain, i.e. where the reference itself is undefined."""
        n = len(size)
        cw, ch = chroma_dims(w, h)
        y = np.zeros((n, h, w), np.uint8)
        u = np.zeros((n, ch, cw), np.uint8)
        v = np.zeros((n, ch, cw), np.uint8)
        m = [np.zeros_like(a) for a in (y, u, v)] if undef else [None] * 3
        st = np.zeros(n, np.int32)
        pk = np.ascontiguousarray(pkts, np.uint8)
        self.lib.amvo_decode_frames(_p(pk), _p(np.ascontiguousarray(off, np.uint64)),
                                    _p(np.ascontiguousarray(size, np.uint32)), n, w, h, _p(y), _p(u), _p(v), _p(st),
                                    _p(m[0]), _p(m[1]), _p(m[2]))
        return (y, u, v, st, tuple(m)) if undef else (y, u, v, st)

    def convert_range(self, y, u, v, direction):
        y, u, v = (np.ascontiguousarray(a, np.uint8) for a in (y, u, v))
        oy, ou, ov = np.zeros_like(y), np.zeros_like(u), np.zeros_like(v)
        self.lib.amvo_convert_range(_p(y), _p(u), _p(v), C.c_size_t(y.size), C.c_size_t(u.size), int(direction), _p(oy), _p(ou), _p(ov))
        return oy, ou, ov

    def scale_frames(self, y, u, v, ow, oh, fill=0):
        """img_resample on tight planes [n,ih,iw] / [n,ich,icw]; bytes the reference leaves alone keep `fill`"""
        y, u, v = (np.ascontiguousarray(a, np.uint8) for a in (y, u, v))
        n, ih, iw = y.shape
        ocw, och = chroma_dims(ow, oh)
        oy, ou, ov = np.full((n, oh, ow), fill, np.uint8), np.full((n, och, ocw), fill, np.uint8), np.full((n, och, ocw), fill, np.uint8)
        if self.lib.amvo_scale_frames(_p(y), _p(u), _p(v), n, iw, ih, ow, oh, _p(oy), _p(ou), _p(ov)) != n:
            raise RuntimeError("oracle scaler failed")
        return oy, ou, ov

    def sws_scale(self, y, u, v, ow, oh, jpeg_in, jpeg_out):
        """the fork's sws_scale as a composition of the restated stages (imgresample.c:617-682): img_convert to
        YUV420P in front of the scaler for a YUVJ420P source, img_convert to YUVJ420P behind it (even sizes)"""
        if jpeg_in:
            y, u, v = self.convert_range(y, u, v, 1)
        y, u, v = self.scale_frames(y, u, v, ow, oh)
        if jpeg_out:
            y, u, v = self.convert_range(y, u, v, 0)
        return y, u, v

    def scale_banks(self, iw, ih, ow, oh):
        hb, vb = np.zeros((16, 4), np.int16), np.zeros((16, 4), np.int16)
        self.lib.amvo_scale_banks(iw, ih, ow, oh, _p(hb), _p(vb))
        return hb, vb

    def resample_bank(self, in_rate, out_rate=22050):
        n = self.lib.amvo_resample_filter_length(in_rate, out_rate)
        bank = np.zeros((1024, n), np.int16)
        self.lib.amvo_resample_bank(in_rate, out_rate, _p(bank))
        return bank

    def audio_resample(self, pcm, in_channels, in_rate, out_rate=22050):
        pcm = np.ascontiguousarray(pcm, np.int16).reshape(-1)
        n_in = pcm.size // in_channels
        cap = int(n_in * out_rate / in_rate) + 64
        out = np.zeros(cap, np.int16)
        self.lib.amvo_audio_resample.restype = C.c_int64
        k = self.lib.amvo_audio_resample(_p(pcm), C.c_int64(n_in), in_channels, in_rate, out_rate, _p(out), C.c_int64(cap))
        if k < 0:
            raise RuntimeError("oracle resampler failed")
        return out[:k].copy()

    def sp5x_decode_frames(self, pkts, off, size, w, h, undef=False):
        n = len(size)
        cw, ch = chroma_dims(w, h)
        y = np.zeros((n, h, w), np.uint8)
        u = np.zeros((n, ch, cw), np.uint8)
        v = np.zeros((n, ch, cw), np.uint8)
        m = [np.zeros_like(a) for a in (y, u, v)] if undef else [None] * 3
        st = np.zeros(n, np.int32)
        self.lib.amvo_sp5x_decode_frames(_p(np.ascontiguousarray(pkts, np.uint8)), _p(np.ascontiguousarray(off, np.uint64)),
                                         _p(np.ascontiguousarray(size, np.uint32)), n, w, h, _p(y), _p(u), _p(v), _p(st),
                                         _p(m[0]), _p(m[1]), _p(m[2]))
        return (y, u, v, st, tuple(m)) if undef else (y, u, v, st)

    def mjpeg_header(self, pkt, chroma=False):
        """-> (w, h, scan_start[, cw, ch]) of a baseline JPEG the path supports, or None"""
        pk = np.ascontiguousarray(pkt, np.uint8)
        w, h, ss, cw, ch = C.c_int(0), C.c_int(0), C.c_uint32(0), C.c_int(0), C.c_int(0)
        if self.lib.amvo_mjpeg_header(_p(pk), len(pk), C.byref(w), C.byref(h), C.byref(ss), C.byref(cw), C.byref(ch)):
            return None
        return (w.value, h.value, ss.value, cw.value, ch.value) if chroma else (w.value, h.value, ss.value)

    def mjpeg_decode_frames(self, pkts, off, size, w, h, undef=False):
        """chroma planes come at the size the frames' sampling gives (first frame's header)"""
        n = len(size)
        pk0 = np.ascontiguousarray(pkts, np.uint8)[int(off[0]): int(off[0]) + int(size[0])]
        hd = self.mjpeg_header(pk0, chroma=True)
        cw, ch = (hd[3], hd[4]) if hd else chroma_dims(w, h)
        y = np.zeros((n, h, w), np.uint8)
        u = np.zeros((n, ch, cw), np.uint8)
        v = np.zeros((n, ch, cw), np.uint8)
        m = [np.zeros_like(a) for a in (y, u, v)] if undef else [None] * 3
        st = np.zeros(n, np.int32)
        self.lib.amvo_mjpeg_decode_frames(_p(np.ascontiguousarray(pkts, np.uint8)), _p(np.ascontiguousarray(off, np.uint64)),
                                          _p(np.ascontiguousarray(size, np.uint32)), n, w, h, cw, ch, _p(y), _p(u), _p(v), _p(st),
                                          _p(m[0]), _p(m[1]), _p(m[2]))
        return (y, u, v, st, tuple(m)) if undef else (y, u, v, st)

    def decode_frame_coefs(self, pkt, w, h):
        """Dequantised coefficients of every block in bitstream order (raster inside a block)."""
        mbw, mbh = (w + 15) // 16, (h + 15) // 16
        cw, ch = chroma_dims(w, h)
        coef = np.zeros((mbw * mbh * 6, 64), np.int16)
        y = np.zeros((h, w), np.uint8)
        u = np.zeros((ch, cw), np.uint8)
        v = np.zeros((ch, cw), np.uint8)
        pk = np.ascontiguousarray(pkt, np.uint8)
        st = self.lib.amvo_decode_frame(_p(pk), len(pk), w, h, _p(y), _p(u), _p(v), w, cw, _p(coef))
        return coef, st

    def unstuff(self, pkt):
        pk = np.ascontiguousarray(pkt, np.uint8)
        dst = np.zeros(len(pk) + 16, np.uint8)
        fl = C.c_int(0)
        n = self.lib.amvo_unstuff(_p(pk), len(pk), _p(dst), C.byref(fl))
        return dst[:n].copy(), fl.value

    # -- audio
    def adpcm_decode(self, chunks, off, size):
        n = len(size)
        ns = np.maximum(np.asarray(size, np.int64) - 8, 0) * 2
        poff = offsets_of(ns)
        pcm = np.zeros(int(ns.sum()), np.int16)
        st = np.zeros(n, np.int32)
        self.lib.amvo_adpcm_decode_chunks(_p(np.ascontiguousarray(chunks, np.uint8)),
                                          _p(np.ascontiguousarray(off, np.uint64)),
                                          _p(np.ascontiguousarray(size, np.uint32)), n, _p(pcm), _p(poff), _p(st))
        return pcm, poff, st

    def adpcm_encode_trellis(self, pcm, pcm_off, nsamples, step_in, trellis):
        """chunks encoded independently with the -trellis N beam search (adpcm.c:287-443)"""
        n = len(nsamples)
        pcm = np.ascontiguousarray(pcm, np.int16)
        osz = 8 + np.asarray(nsamples, np.uint64) // 2
        ooff = offsets_of(osz)
        out = np.zeros(int(osz.sum()), np.uint8)
        step_out = np.zeros(n, np.int16)
        for i in range(n):
            so = C.c_int(0)
            a, m = int(pcm_off[i]), int(nsamples[i])
            r = self.lib.amvo_adpcm_encode_chunk_trellis(_p(pcm[a:a + m].copy()), m, int(step_in[i]), int(trellis),
                                                         _p(out[int(ooff[i]):]), C.byref(so))
            if r != int(osz[i]):
                raise RuntimeError("oracle trellis encode failed: %d" % r)
            step_out[i] = so.value
        return out, ooff, osz.astype(np.uint32), step_out

    def adpcm_encode(self, pcm, pcm_off, nsamples, step_in):
        n = len(nsamples)
        nsamples = np.ascontiguousarray(nsamples, np.uint32)
        osz = 8 + nsamples.astype(np.uint64) // 2
        ooff = offsets_of(osz)
        out = np.zeros(int(osz.sum()), np.uint8)
        step_out = np.zeros(n, np.int16)
        r = self.lib.amvo_adpcm_encode_chunks(_p(np.ascontiguousarray(pcm, np.int16)),
                                              _p(np.ascontiguousarray(pcm_off, np.uint64)), _p(nsamples),
                                              _p(np.ascontiguousarray(step_in, np.int16)), _p(step_out), n,
                                              _p(out), _p(ooff))
        if r != n:
            raise RuntimeError("oracle adpcm encode failed: %d" % r)
        return out, ooff, osz.astype(np.uint32), step_out


class AmvlibRef:
    """The unmodified reference amvlib (C-AMVDecoder/amvlib), compiled in place by oracle/build_ref.sh."""

    @staticmethod
    def available():
        return os.path.exists(AMVLIB_REF_SO)

    def __init__(self):
        self.lib = C.CDLL(AMVLIB_REF_SO)

    def video_decode(self, pkts, off, size, w, h):
        n = len(size)
        lb = self.lib.amvlibref_line_bytes(w)
        bgr = np.zeros((n, h, lb), np.uint8)
        ret = np.zeros(n, np.int32)
        r = self.lib.amvlibref_video_decode(_p(np.ascontiguousarray(pkts, np.uint8)), _p(np.ascontiguousarray(off, np.uint64)),
                                            _p(np.ascontiguousarray(size, np.uint32)), n, w, h, _p(bgr), _p(ret))
        if r != n:
            raise RuntimeError("amvlib reference decode failed: %d" % r)
        return bgr, ret

    def audio_decode(self, chunks, off, size):
        n = len(size)
        ns = (np.maximum(np.asarray(size, np.int64) - 8, 0) + 3) // 4 * 8
        poff = offsets_of(ns)
        pcm = np.zeros(int(ns.sum()) + 16, np.int16)
        nsamp = np.zeros(n, np.uint32)
        ret = np.zeros(n, np.int32)
        r = self.lib.amvlibref_audio_decode(_p(np.ascontiguousarray(chunks, np.uint8)), _p(np.ascontiguousarray(off, np.uint64)),
                                            _p(np.ascontiguousarray(size, np.uint32)), n, _p(pcm), _p(poff), _p(nsamp), _p(ret))
        if r != n:
            raise RuntimeError("amvlib reference audio decode failed: %d" % r)
        return pcm[: int(ns.sum())], poff, nsamp, ret


class Ref:
    """The unmodified reference codecs (libavcodec 51.47.1 of the AMVmuxer fork)."""

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def __init__(self):
        self.lib = C.CDLL(REF_SO)
        self.lib.amvref_version.restype = C.c_char_p

    def version(self):
        return self.lib.amvref_version().decode()

    def encode_frames(self, y, u, v, w, h, quality=0, cap=None):
        n = y.shape[0]
        cap = cap or n * (w * h * 3 + 4096)
        out = np.zeros(cap, np.uint8)
        off = np.zeros(n, np.uint64)
        size = np.zeros(n, np.uint32)
        r = self.lib.amvref_encode_frames(_p(y), _p(u), _p(v), n, w, h, int(quality), _p(out), _p(off), _p(size),
                                          C.c_uint64(cap))
        if r != n:
            raise RuntimeError("reference encode failed: %d" % r)
        return out[: int(size.sum())].copy(), off, size

    def mjpeg_encode_frames(self, y, u, v, w, h, quality=0, cap=None):
        """the reference's plain MJPEG encoder (full JPEG frames, tables in the stream)"""
        n = y.shape[0]
        cap = cap or n * (w * h * 3 + 4096)
        out = np.zeros(cap, np.uint8)
        off = np.zeros(n, np.uint64)
        size = np.zeros(n, np.uint32)
        r = self.lib.amvref_mjpeg_encode_frames(_p(y), _p(u), _p(v), n, w, h, int(u.shape[1]), int(quality), _p(out), _p(off),
                                                _p(size), C.c_uint64(cap))
        if r != n:
            raise RuntimeError("reference mjpeg encode failed: %d" % r)
        return out[: int(size.sum())].copy(), off, size

    def decode_frames(self, pkts, off, size, w, h, sp5x=False, mjpeg=False, chroma=None):
        """chroma = (cw, ch): chroma plane size of 4:2:2 / 4:4:4 JPEG frames (default: 4:2:0)"""
        n = len(size)
        cw, ch = chroma if chroma else chroma_dims(w, h)
        y = np.zeros((n, h, w), np.uint8)
        u = np.zeros((n, ch, cw), np.uint8)
        v = np.zeros((n, ch, cw), np.uint8)
        got = np.zeros(n, np.int32)
        rb = np.zeros(n, np.int32)
        fn = self.lib.amvref_sp5x_decode_frames if sp5x else self.lib.amvref_decode_frames
        if mjpeg:
            fn = self.lib.amvref_mjpeg_decode_frames
        r = fn(_p(np.ascontiguousarray(pkts, np.uint8)),
                                          _p(np.ascontiguousarray(off, np.uint64)),
                                          _p(np.ascontiguousarray(size, np.uint32)), n, w, h, _p(y), _p(u), _p(v),
                                          _p(got), _p(rb))
        if r != n:
            raise RuntimeError("reference decode failed: %d" % r)
        return y, u, v, got, rb

    def adpcm_decode(self, chunks, off, size):
        n = len(size)
        ns = np.maximum(np.asarray(size, np.int64) - 8, 0) * 2
        poff = offsets_of(ns)
        pcm = np.zeros(int(ns.sum()), np.int16)
        nsamp = np.zeros(n, np.uint32)
        r = self.lib.amvref_adpcm_decode(_p(np.ascontiguousarray(chunks, np.uint8)),
                                         _p(np.ascontiguousarray(off, np.uint64)),
                                         _p(np.ascontiguousarray(size, np.uint32)), n, _p(pcm), _p(poff), _p(nsamp))
        if r != n:
            raise RuntimeError("reference adpcm decode failed: %d" % r)
        return pcm, poff, nsamp

    def adpcm_encode_stream(self, pcm, frame_size, max_chunks=1 << 20, trellis=0):
        pcm = np.ascontiguousarray(pcm, np.int16)
        cap = len(pcm) + 16 * (len(pcm) // max(frame_size, 1) + 4) + 65536
        out = np.zeros(cap, np.uint8)
        mc = min(max_chunks, len(pcm) // max(frame_size, 1) + 2)
        off = np.zeros(mc, np.uint64)
        size = np.zeros(mc, np.uint32)
        cons = np.zeros(mc, np.uint32)
        k = self.lib.amvref_adpcm_encode_stream_trellis(_p(pcm), C.c_uint64(len(pcm)), int(frame_size), int(trellis), _p(out),
                                                        _p(off), _p(size), _p(cons), mc, C.c_uint64(cap))
        if k < 0:
            raise RuntimeError("reference adpcm encode failed: %d" % k)
        return out[: int(size[:k].sum())].copy(), off[:k], size[:k], cons[:k]

    def fdct(self, blocks):
        b = np.ascontiguousarray(blocks, dtype=np.int16).reshape(-1, 64).copy()
        self.lib.amvref_fdct_islow(_p(b), b.shape[0])
        return b

    def convert_range(self, y, u, v, direction):
        """img_convert between PIX_FMT_YUV420P and PIX_FMT_YUVJ420P (tight planes [n,h,w] / [n,ch,cw])"""
        y, u, v = (np.ascontiguousarray(a, np.uint8) for a in (y, u, v))
        n, h, w = y.shape
        oy, ou, ov = np.zeros_like(y), np.zeros_like(u), np.zeros_like(v)
        r = self.lib.amvref_convert_range(_p(y), _p(u), _p(v), n, w, h, int(direction), _p(oy), _p(ou), _p(ov))
        if r != n:
            raise RuntimeError("reference img_convert failed: %d" % r)
        return oy, ou, ov

    def scale_frames(self, y, u, v, ow, oh, fill=0):
        """img_resample_init + img_resample (what sws_scale of the fork calls) on tight planes"""
        y, u, v = (np.ascontiguousarray(a, np.uint8) for a in (y, u, v))
        n, ih, iw = y.shape
        ocw, och = chroma_dims(ow, oh)
        oy, ou, ov = np.full((n, oh, ow), fill, np.uint8), np.full((n, och, ocw), fill, np.uint8), np.full((n, och, ocw), fill, np.uint8)
        if self.lib.amvref_img_resample(_p(y), _p(u), _p(v), n, iw, ih, ow, oh, _p(oy), _p(ou), _p(ov)) != n:
            raise RuntimeError("reference img_resample failed")
        return oy, ou, ov

    def sws_scale(self, y, u, v, ow, oh, jpeg_in, jpeg_out):
        """sws_getContext + sws_scale of the fork (what ffmpeg.c calls for -s): YUV420P / YUVJ420P on either side"""
        y, u, v = (np.ascontiguousarray(a, np.uint8) for a in (y, u, v))
        n, ih, iw = y.shape
        ocw, och = chroma_dims(ow, oh)
        oy, ou, ov = np.zeros((n, oh, ow), np.uint8), np.zeros((n, och, ocw), np.uint8), np.zeros((n, och, ocw), np.uint8)
        if self.lib.amvref_sws_scale(_p(y), _p(u), _p(v), n, iw, ih, ow, oh, int(jpeg_in), int(jpeg_out), _p(oy), _p(ou), _p(ov)) != n:
            raise RuntimeError("reference sws_scale failed")
        return oy, ou, ov

    def audio_resample(self, pcm, in_channels, in_rate, out_rate=22050, chunk=4096):
        """audio_resample_init(1, in_channels, out_rate, in_rate) + audio_resample fed `chunk` samples per call"""
        pcm = np.ascontiguousarray(pcm, np.int16).reshape(-1)
        n_in = pcm.size // in_channels
        cap = int(n_in * out_rate / in_rate) + 64
        out = np.zeros(cap, np.int16)
        self.lib.amvref_audio_resample.restype = C.c_int64
        k = self.lib.amvref_audio_resample(_p(pcm), C.c_int64(n_in), in_channels, in_rate, out_rate, int(chunk), _p(out), C.c_int64(cap))
        if k < 0:
            raise RuntimeError("reference audio_resample failed")
        return out[:k].copy()

    def resample_bank(self, in_rate, out_rate=22050):
        """filter bank of av_resample_init(out_rate, in_rate, 16, 10, 0, 0.8): [1024, filter_length]"""
        buf = np.zeros(1025 * 8192, np.int16)
        ln = C.c_int(0)
        if self.lib.amvref_resample_bank(out_rate, in_rate, _p(buf), buf.size, C.byref(ln)) < 0:
            raise RuntimeError("reference av_resample_init failed")
        return buf[: 1024 * ln.value].reshape(1024, ln.value).copy()

    # -- container (libavformat amv_muxer / avi_demuxer, driven like ffmpeg.c does)
    def mux(self, w, h, fps, sample_rate, vpk, voff, vsz, apk, aoff, asz):
        n = len(vsz)
        cap = int(np.sum(vsz, dtype=np.uint64) + np.sum(asz, dtype=np.uint64)) + 16 * n + 4096
        out = np.zeros(cap, np.uint8)
        self.lib.amvref_mux.restype = C.c_int64
        r = self.lib.amvref_mux(w, h, fps, sample_rate, n, _p(np.ascontiguousarray(vpk, np.uint8)),
                                _p(np.ascontiguousarray(voff, np.uint64)), _p(np.ascontiguousarray(vsz, np.uint32)),
                                _p(np.ascontiguousarray(apk, np.uint8)), _p(np.ascontiguousarray(aoff, np.uint64)),
                                _p(np.ascontiguousarray(asz, np.uint32)), _p(out), C.c_uint64(cap))
        if r < 0:
            raise RuntimeError("reference mux failed: %d" % r)
        return out[:r].tobytes()

    def demux(self, data):
        """-> (info[w,h,fps,rate,nv,na], [video packets], [audio chunks]) as the reference demuxer reads them"""
        fb = np.frombuffer(data, np.uint8).copy()
        cap_units, cap_bytes = len(fb) // 8 + 16, len(fb) + 64
        info = np.zeros(6, np.int32)
        vd, ad = np.zeros(cap_bytes, np.uint8), np.zeros(cap_bytes, np.uint8)
        vs, as_ = np.zeros(cap_units, np.uint32), np.zeros(cap_units, np.uint32)
        r = self.lib.amvref_demux(_p(fb), C.c_uint64(len(fb)), _p(info), _p(vd), _p(vs), _p(ad), _p(as_), cap_units,
                                  C.c_uint64(cap_bytes))
        if r < 0:
            raise RuntimeError("reference demux failed: %d" % r)
        vo = np.concatenate([[0], np.cumsum(vs[: info[4]], dtype=np.int64)])
        ao = np.concatenate([[0], np.cumsum(as_[: info[5]], dtype=np.int64)])
        return info, [vd[vo[i]:vo[i + 1]].tobytes() for i in range(info[4])], [ad[ao[i]:ao[i + 1]].tobytes() for i in range(info[5])]

    def idct_put(self, blocks):
        b = np.ascontiguousarray(blocks, dtype=np.int16).reshape(-1, 64)
        out = np.zeros((b.shape[0], 64), np.uint8)
        self.lib.amvref_simple_idct_put(_p(b), b.shape[0], _p(out))
        return out


# ----------------------------------------------------------------- generators

def synth_frames(n, w, h, seed=1, t0=0, kind="sinus"):
    """YUVJ420P test content. kind: sinus | noise | flat | edges"""
    cw, ch = chroma_dims(w, h)
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return (rng.integers(0, 256, (n, h, w), dtype=np.uint8),
                rng.integers(0, 256, (n, ch, cw), dtype=np.uint8),
                rng.integers(0, 256, (n, ch, cw), dtype=np.uint8))
    if kind == "flat":
        lv = rng.integers(0, 256, (n, 3))
        y = np.broadcast_to(lv[:, 0, None, None], (n, h, w)).astype(np.uint8).copy()
        u = np.broadcast_to(lv[:, 1, None, None], (n, ch, cw)).astype(np.uint8).copy()
        v = np.broadcast_to(lv[:, 2, None, None], (n, ch, cw)).astype(np.uint8).copy()
        if n > 1:  # one frame with a few isolated impulses: single-coefficient rows/columns
            y[1, ::13, ::7] = 255
        return y, u, v
    if kind == "edges":
        y = np.zeros((n, h, w), np.uint8)
        u = np.full((n, ch, cw), 128, np.uint8)
        v = np.full((n, ch, cw), 128, np.uint8)
        for i in range(n):
            p = 1 + (i % 7)
            yy, xx = np.mgrid[0:h, 0:w]
            y[i] = (((xx // p) + (yy // p)) & 1) * 255
            u[i, :, ::2] = 0 if i & 1 else 255
            v[i, ::2, :] = 255 if i & 1 else 0
        return y, u, v
    t = (np.arange(n) + t0)[:, None, None].astype(np.float64)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    Y = 128 + 60 * np.sin((xx[None] + 3 * t) / 17.0) + 50 * np.cos((yy[None] - 2 * t) / 11.0)
    Y = Y + rng.normal(0.0, 6.0, size=Y.shape)
    cy, cx = np.mgrid[0:ch, 0:cw].astype(np.float64)
    U = 128 + 40 * np.sin((cx[None] + t) / 23.0) + 0 * cy[None]
    V = 128 + 40 * np.cos((cy[None] + t) / 19.0) + 0 * cx[None]
    f = lambda a: np.clip(np.rint(a), 0, 255).astype(np.uint8)
    return f(Y), f(U), f(V)


def synth_pcm(nsamples, seed=1, kind="tones"):
    t = np.arange(nsamples, dtype=np.float64)
    if kind == "noise":
        return np.random.default_rng(seed).integers(-32768, 32768, nsamples).astype(np.int16)
    if kind == "square":
        return np.where((np.arange(nsamples) // 37) & 1, 32767, -32768).astype(np.int16)
    if kind == "silence":
        return np.zeros(nsamples, np.int16)
    s = 8000 * np.sin(2 * np.pi * 440 * t / 22050) + 2000 * np.sin(2 * np.pi * 1234 * t / 22050)
    return np.rint(s).astype(np.int16)


# ------------------------------------------------------------ AMV container

def walk_amv(data):
    """Chunk walker in the spirit of AMVmuxer/compare_amv.c:44-94: returns
    (width, height, fps, [video packets], [audio chunks])."""
    i = data.find(b"amvh")
    hdr = struct.unpack_from("<14I", data, i + 8)
    w, h, fps = hdr[8], hdr[9], hdr[10]
    p = data.find(b"movi") + 4
    vids, auds = [], []
    while p + 8 <= len(data):
        tag = data[p:p + 4]
        if tag not in (b"00dc", b"01wb"):
            break
        sz = struct.unpack_from("<I", data, p + 4)[0]
        (vids if tag == b"00dc" else auds).append(data[p + 8:p + 8 + sz])
        p += 8 + sz
    return w, h, fps, vids, auds


def sp5x_from_amv(oracle, pkts, off, size, header=None):
    """SP5X packets carrying the scans of the given AMV packets: 14 header bytes (the decoder skips them,
    sp5xdec.c:78) + the un-stuffed scan with literal FF bytes; no encoder for SP5X exists in the reference."""
    hdr = bytes(header if header is not None else range(0xF1, 0xFF)) 
    assert len(hdr) == 14
    out = []
    for o, s in zip(off, size):
        scan, fl = oracle.unstuff(np.asarray(pkts[int(o):int(o) + int(s)]))
        assert fl == 0 and scan[-1] == 0xFF
        out.append(hdr + scan[:-1].tobytes())          # the trailing FF is the appended EOI's, not payload
    return pack(out)


def mjpeg_with_dqt(pkts, off, size, seed):
    """the same JPEG frames with the 64 quantisers of their (first) DQT segment replaced by seeded values 1..60:
    same scan, other dequantisation -- every frame gets the same table, so the headers stay identical"""
    out = np.array(pkts, np.uint8, copy=True)
    q = np.random.default_rng(seed).integers(1, 61, 64).astype(np.uint8)
    for o, s in zip(off, size):
        d = out[int(o): int(o) + int(s)]
        j = bytes(d[:1024]).find(b"\xff\xdb")
        assert j > 0 and d[j + 4] == 0
        d[j + 5: j + 5 + 64] = q
    return out


def jpeg_encode_simple(oracle, y, u, v, sampling, q=(16, 24), restart=0):
    """A minimal baseline JPEG writer (test input only; what it means is defined by the reference DECODER):
    sampling = ((hY, vY), (hC, vC)); flat quantisers q[0] (component 0) / q[1] (components 1, 2), standard Huffman
    tables, float DCT; restart = MCUs per restart interval (DRI + RSTn markers), 0 for none.
    y is [h, w]; u, v are the chroma planes at the size the sampling implies."""
    from scipy.fft import dctn
    (hy, vy), (hc, vc) = sampling
    h, w = y.shape
    mbw, mbh = -(-w // (8 * hy)), -(-h // (8 * vy))
    zz = oracle.zigzag()
    huff = [oracle.huff(t) for t in range(4)]
    bits = []

    def put(val, n):
        if n:
            bits.append((int(val) & ((1 << n) - 1), n))

    def category(vv):
        a = abs(int(vv))
        return a.bit_length()

    def pad(p, ph, pw):
        return np.pad(p, ((0, ph - p.shape[0]), (0, pw - p.shape[1])), mode="edge")

    planes = [pad(y, mbh * 8 * vy, mbw * 8 * hy), pad(u, mbh * 8 * vc, mbw * 8 * hc), pad(v, mbh * 8 * vc, mbw * 8 * hc)]
    fac = [(hy, vy), (hc, vc), (hc, vc)]
    pred = [0, 0, 0]
    nmcu, rst = 0, 0
    for my in range(mbh):
        for mx in range(mbw):
            if restart and nmcu and nmcu % restart == 0:
                bits.append(("rst", rst))            # pad to a byte with ones, FF D0+n, predictors start over
                rst = (rst + 1) & 7
                pred = [0, 0, 0]
            nmcu += 1
            for c in range(3):
                hh, vv = fac[c]
                for b in range(hh * vv):
                    bx, by = (hh * mx + b % hh) * 8, (vv * my + b // hh) * 8
                    blk = planes[c][by:by + 8, bx:bx + 8].astype(np.float64) - 128.0
                    co = np.rint(dctn(blk, norm="ortho") / (q[0] if c == 0 else q[1])).astype(int).reshape(-1)[zz]
                    tq = 0 if c == 0 else 1
                    d = int(co[0]) - pred[c]
                    pred[c] = int(co[0])
                    n = category(d)
                    put(huff[tq][1][n], int(huff[tq][0][n]))
                    put(d if d >= 0 else d - 1, n)
                    run = 0
                    last = max([k for k in range(1, 64) if co[k]] or [0])
                    for k in range(1, last + 1):
                        if co[k] == 0:
                            run += 1
                            continue
                        while run >= 16:
                            put(huff[2 + tq][1][0xF0], int(huff[2 + tq][0][0xF0]))
                            run -= 16
                        n = category(co[k])
                        sym = (run << 4) | n
                        put(huff[2 + tq][1][sym], int(huff[2 + tq][0][sym]))
                        put(co[k] if co[k] >= 0 else co[k] - 1, n)
                        run = 0
                    if last < 63:
                        put(huff[2 + tq][1][0], int(huff[2 + tq][0][0]))
    acc, nacc, scan = 0, 0, bytearray()
    for val, n in bits:
        if val == "rst":
            if nacc:
                byte = ((acc << (8 - nacc)) | ((1 << (8 - nacc)) - 1)) & 0xFF
                scan.append(byte)
                if byte == 0xFF:
                    scan.append(0)
            acc, nacc = 0, 0
            scan += bytes([0xFF, 0xD0 + n])
            continue
        acc = (acc << n) | val
        nacc += n
        while nacc >= 8:
            byte = (acc >> (nacc - 8)) & 0xFF
            scan.append(byte)
            if byte == 0xFF:
                scan.append(0)
            nacc -= 8
    if nacc:
        byte = ((acc << (8 - nacc)) | ((1 << (8 - nacc)) - 1)) & 0xFF
        scan.append(byte)
        if byte == 0xFF:
            scan.append(0)
    out = bytearray(b"\xff\xd8")
    for tid, qq in ((0, q[0]), (1, q[1])):
        out += b"\xff\xdb" + (67).to_bytes(2, "big") + bytes([tid]) + bytes([qq] * 64)
    dht = bytearray()
    for t, (cls, tid) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        ln, cd = huff[t]
        syms = sorted([sm for sm in range(256) if ln[sm]], key=lambda sm: (int(ln[sm]), int(cd[sm])))
        counts = [sum(1 for sm in syms if ln[sm] == L) for L in range(1, 17)]
        dht += bytes([(cls << 4) | tid]) + bytes(counts) + bytes(syms)
    out += b"\xff\xc4" + (2 + len(dht)).to_bytes(2, "big") + dht
    out += b"\xff\xc0" + (17).to_bytes(2, "big") + bytes([8]) + h.to_bytes(2, "big") + w.to_bytes(2, "big") + bytes([3])
    out += bytes([1, (hy << 4) | vy, 0, 2, (hc << 4) | vc, 1, 3, (hc << 4) | vc, 1])
    if restart:
        out += b"\xff\xdd" + (4).to_bytes(2, "big") + int(restart).to_bytes(2, "big")
    out += b"\xff\xda" + (12).to_bytes(2, "big") + bytes([3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0])
    out += scan + b"\xff\xd9"
    return np.frombuffer(bytes(out), np.uint8)


def resample_chroma(a, w, h, sampling):
    """4:2:0 chroma planes [n, ch, cw] -> planes at the size `sampling` implies (nearest neighbour; test input only)"""
    (hy, vy), (hc, vc) = sampling
    cw, ch = -(-w * hc // hy), -(-h * vc // vy)
    full = np.repeat(np.repeat(a, 2, axis=1), 2, axis=2)[:, :h, :w]
    return full[:, ::(vy // vc), ::(hy // hc)][:, :ch, :cw].copy()


def pack(chunks):
    size = np.array([len(c) for c in chunks], np.uint32)
    return np.frombuffer(b"".join(chunks), np.uint8).copy(), offsets_of(size), size
