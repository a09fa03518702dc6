"""amv-codec-tools_b200 -- B200-native AMV codec path (libamvcuda) and its Python host binding.

The product is the C-ABI shared library ``lib/libamvcuda.so`` (include/amvcuda.h) built from the
hand-written sm_100a kernels under ``csrc/``.  This module is a thin ctypes front-end over that
ABI -- the same calls the reference-side AVCodec / amvlib shims make -- used by
tests/, bench.py and __graft_entry__.py.  PyTorch appears only as plumbing (device buffers, streams).

There is NO CPU path: if the library is missing or no sm_100 device is present, construction of
:class:`AmvCuda` raises.

Import name: the directory name is not a Python identifier; ``import amv_codec_tools_b200`` (the
alias module at the repository root) loads this package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import sharding  # noqa: F401  (host-side multi-GPU partitioning)

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "lib", "libamvcuda.so")
HEADER_PATH = os.path.join(os.path.dirname(PKG_DIR), "include", "amvcuda.h")

MEM_HOST, MEM_DEVICE = 0, 1
SCALE_IN_JPEG_RANGE, SCALE_OUT_JPEG_RANGE = 1, 2
LAYOUT_PACKED, LAYOUT_SLOTS = 0, 1

ST_SHORT, ST_BADCODE, ST_COEFIDX, ST_MARKER, ST_OVERRUN, ST_RANGE, ST_NOSPACE, ST_HEADER = (1 << i for i in range(8))

EXPORTS = [
    "amv_create", "amv_destroy", "amv_set_stream", "amv_sync", "amv_strerror", "amv_last_error", "amv_version",
    "amv_launch_count", "amv_host_alloc", "amv_host_free", "amv_set_option", "amv_get_stat",
    "amv_qscale_from_quality", "amv_decode_frames", "amv_encode_frames", "amv_adpcm_dec_chunks",
    "amv_adpcm_enc_chunks", "amv_adpcm_enc_streams", "amv_decode_frames_bgr24",
    "amv_file_index", "amv_file_mux", "amv_decode_frames_sp5x",
    "amv_convert_range", "amv_mjpeg_configure", "amv_decode_frames_mjpeg",
    "amv_scale_frames", "amv_audio_resample", "amv_audio_resample_count", "amv_scale_banks", "amv_audio_resample_bank", "amv_audio_resample_from", "amv_audio_resample_first_tap", "amv_scale_frames_ex",
]


class AmvError(RuntimeError):
    pass


def build(verbose=False):
    """Compile libamvcuda.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", PKG_DIR, "-j4"], capture_output=True, text=True)
    if out.returncode != 0:
        raise AmvError("building libamvcuda failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)
    return LIB_PATH


class FileInfo(C.Structure):
    """amv_file_info of include/amvcuda.h"""
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("fps", C.c_int), ("sample_rate", C.c_int), ("channels", C.c_int),
                ("us_per_frame", C.c_uint32), ("nb_frames_header", C.c_uint32), ("duration_s", C.c_uint32),
                ("nvideo", C.c_uint32), ("naudio", C.c_uint32), ("movi_offset", C.c_uint64),
                ("has_end_marker", C.c_int), ("truncated", C.c_int)]


class MuxParams(C.Structure):
    """amv_mux_params of include/amvcuda.h"""
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("tb_num", C.c_int), ("tb_den", C.c_int), ("sample_rate", C.c_int),
                ("video_bit_rate", C.c_int), ("audio_bit_rate", C.c_int)]


class _Params(C.Structure):
    _fields_ = [("device", C.c_int), ("stream", C.c_void_p), ("flags", C.c_uint32)]


def load_library(path=LIB_PATH):
    if not os.path.exists(path):
        raise AmvError("libamvcuda.so not built (%s): run amv_codec_tools_b200.build() -- there is no fallback" % path)
    lib = C.CDLL(path)
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
    lib.amv_create.argtypes = [C.POINTER(_Params), C.POINTER(vp)]
    lib.amv_destroy.argtypes = [vp]
    lib.amv_destroy.restype = None
    lib.amv_set_stream.argtypes = [vp, vp]
    lib.amv_sync.argtypes = [vp]
    lib.amv_strerror.argtypes = [i32]
    lib.amv_strerror.restype = C.c_char_p
    lib.amv_last_error.argtypes = [vp]
    lib.amv_last_error.restype = C.c_char_p
    lib.amv_launch_count.argtypes = [vp]
    lib.amv_launch_count.restype = u64
    lib.amv_host_alloc.argtypes = [C.c_size_t]
    lib.amv_host_alloc.restype = vp
    lib.amv_host_free.argtypes = [vp]
    lib.amv_host_free.restype = None
    lib.amv_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    lib.amv_get_stat.argtypes = [vp, C.c_char_p]
    lib.amv_get_stat.restype = C.c_int64
    lib.amv_qscale_from_quality.argtypes = [i32, i32, i32]
    lib.amv_decode_frames.argtypes = [vp, vp, u64, vp, vp, i32, i32, i32, vp, vp, vp, i32, i32, u64, u64, vp, i32]
    lib.amv_decode_frames_sp5x.argtypes = lib.amv_decode_frames.argtypes
    lib.amv_decode_frames_mjpeg.argtypes = lib.amv_decode_frames.argtypes
    lib.amv_mjpeg_configure.argtypes = [vp, vp, C.c_uint32, vp, vp]
    lib.amv_convert_range.argtypes = [vp, vp, vp, vp, i32, i32, u64, u64, i32, i32, i32, i32, vp, vp, vp, i32, i32, u64, u64, i32]
    lib.amv_scale_frames.argtypes = [vp, vp, vp, vp, i32, i32, u64, u64, i32, i32, i32, vp, vp, vp, i32, i32, u64, u64, i32, i32, i32]
    lib.amv_scale_frames_ex.argtypes = [vp, vp, vp, vp, i32, i32, u64, u64, i32, i32, i32, vp, vp, vp, i32, i32, u64, u64, i32, i32, i32, i32]
    lib.amv_audio_resample.argtypes = [vp, vp, u64, i32, i32, i32, vp, u64, vp, i32]
    lib.amv_audio_resample_from.argtypes = [vp, vp, u64, u64, i32, i32, i32, u64, vp, u64, vp, i32]
    lib.amv_audio_resample_first_tap.argtypes = [u64, i32, i32]
    lib.amv_audio_resample_first_tap.restype = C.c_int64
    lib.amv_audio_resample_count.argtypes = [u64, i32, i32]
    lib.amv_audio_resample_count.restype = u64
    lib.amv_scale_banks.argtypes = [i32, i32, i32, i32, vp, vp, vp, vp]
    lib.amv_audio_resample_bank.argtypes = [i32, i32, vp, u64]
    lib.amv_decode_frames_bgr24.argtypes = [vp, vp, u64, vp, vp, i32, i32, i32, vp, i32, u64, vp, i32]
    lib.amv_file_index.argtypes = [vp, u64, C.POINTER(FileInfo), vp, vp, vp, vp, u32]
    lib.amv_file_mux.argtypes = [C.POINTER(MuxParams), i32, vp, vp, vp, vp, vp, vp, vp, u64]
    lib.amv_file_mux.restype = C.c_int64
    lib.amv_encode_frames.argtypes = [vp, vp, vp, vp, i32, i32, u64, u64, i32, i32, i32, vp, vp, u64, u32, i32, vp, vp,
                                      vp, i32]
    lib.amv_adpcm_dec_chunks.argtypes = [vp, vp, u64, vp, vp, i32, vp, u64, vp, vp, i32]
    lib.amv_adpcm_enc_chunks.argtypes = [vp, vp, u64, vp, vp, vp, vp, i32, vp, u64, vp, vp, i32]
    lib.amv_adpcm_enc_streams.argtypes = [vp, vp, u64, vp, vp, vp, i32, i32, vp, vp, vp, u64, vp, vp, i32]
    return lib


def _ptr(a):
    """Raw address of a numpy array (host) or a torch tensor (device or pinned host); None -> NULL."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise AmvError("array must be C-contiguous")
        return a.ctypes.data
    return a.data_ptr()          # torch tensor


def chroma_dims(w, h):
    return (w + 1) // 2, (h + 1) // 2


def amvlib_line_bytes(w):
    """bytes per bitmap row the reference amvlib uses: WIDTHBYTES(w*24) (amvlib/AmvJpeg.c:420,1526)"""
    return (w * 24 + 31) // 32 * 4


def offsets_of(sizes):
    sizes = np.asarray(sizes, dtype=np.uint64)
    off = np.zeros(len(sizes), dtype=np.uint64)
    if len(sizes) > 1:
        off[1:] = np.cumsum(sizes)[:-1]
    return off


def scale_banks(iw, ih, ow, oh, lib=None):
    """(h_bank[16,4], v_bank[16,4], h_incr, v_incr) the scaler runs on (host-side filter design, no device)"""
    lib = lib or load_library()
    hb, vb = np.zeros((16, 4), np.int16), np.zeros((16, 4), np.int16)
    hi, vi = C.c_int32(0), C.c_int32(0)
    r = lib.amv_scale_banks(iw, ih, ow, oh, hb.ctypes.data, vb.ctypes.data, C.addressof(hi), C.addressof(vi))
    if r < 0:
        raise AmvError("amv_scale_banks: %d" % r)
    return hb, vb, hi.value, vi.value


def audio_resample_bank(in_rate, out_rate=22050, lib=None):
    """the [1024, filter_length] polyphase bank of the audio resampler (host-side filter design, no device)"""
    lib = lib or load_library()
    n = lib.amv_audio_resample_bank(in_rate, out_rate, None, 0)
    if n < 0:
        raise AmvError("amv_audio_resample_bank: %d" % n)
    bank = np.zeros((1024, n), np.int16)
    lib.amv_audio_resample_bank(in_rate, out_rate, bank.ctypes.data, bank.size)
    return bank


def file_index(data, lib=None):
    """Index an AMV file held in memory (bytes / numpy uint8).  Host-only, needs no device.
    -> (FileInfo, v_off, v_size, a_off, a_size): offsets point into `data`."""
    lib = lib or load_library()
    buf = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, np.uint8)
    info = FileInfo()
    r = lib.amv_file_index(buf.ctypes.data, buf.nbytes, C.byref(info), None, None, None, None, 0)
    if r != 0:
        raise AmvError("amv_file_index: not an AMV file (%d)" % r)
    cap = max(info.nvideo, info.naudio, 1)
    v_off, a_off = np.zeros(cap, np.uint64), np.zeros(cap, np.uint64)
    v_size, a_size = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32)
    lib.amv_file_index(buf.ctypes.data, buf.nbytes, C.byref(info), v_off.ctypes.data, v_size.ctypes.data, a_off.ctypes.data,
                       a_size.ctypes.data, cap)
    return info, v_off[: info.nvideo], v_size[: info.nvideo], a_off[: info.naudio], a_size[: info.naudio]


def file_mux(width, height, fps, sample_rate, vpk, v_off, v_size, apk, a_off, a_size, video_bit_rate=0, audio_bit_rate=0,
             lib=None):
    """Write an AMV file (bytes) from n video packets and n audio chunks.  Host-only."""
    lib = lib or load_library()
    n = len(v_size)
    assert len(a_size) == n
    mp = MuxParams(width, height, 1, fps, sample_rate, video_bit_rate, audio_bit_rate)
    vpk, apk = np.ascontiguousarray(vpk, np.uint8), np.ascontiguousarray(apk, np.uint8)
    v_off, a_off = np.ascontiguousarray(v_off, np.uint64), np.ascontiguousarray(a_off, np.uint64)
    v_size, a_size = np.ascontiguousarray(v_size, np.uint32), np.ascontiguousarray(a_size, np.uint32)
    args = (vpk.ctypes.data, v_off.ctypes.data, v_size.ctypes.data, apk.ctypes.data, a_off.ctypes.data, a_size.ctypes.data)
    need = -lib.amv_file_mux(C.byref(mp), n, *args, None, 0)
    if need <= 0:
        raise AmvError("amv_file_mux: bad arguments (%d)" % -need)
    out = np.zeros(need, np.uint8)
    got = lib.amv_file_mux(C.byref(mp), n, *args, out.ctypes.data, out.nbytes)
    if got != need:
        raise AmvError("amv_file_mux failed (%d)" % got)
    return out.tobytes()


class AmvCuda:
    """One libamvcuda context (one CUDA stream).  Mirrors include/amvcuda.h one to one."""

    def __init__(self, device=-1, stream=None, lib_path=LIB_PATH):
        self.lib = load_library(lib_path)
        prm = _Params(device, stream, 0)
        ctx = C.c_void_p()
        r = self.lib.amv_create(C.byref(prm), C.byref(ctx))
        if r != 0:
            raise AmvError("amv_create failed: %s" % self.lib.amv_strerror(r).decode())
        self.ctx = ctx

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.amv_destroy(self.ctx)
            self.ctx = None

    __del__ = close

    def _ck(self, r):
        if r != 0:
            raise AmvError("%s (%s)" % (self.lib.amv_strerror(r).decode(), self.lib.amv_last_error(self.ctx).decode()))

    def set_stream(self, stream_ptr):
        self._ck(self.lib.amv_set_stream(self.ctx, stream_ptr))

    def use_torch_stream(self):
        import torch
        self.set_stream(torch.cuda.current_stream().cuda_stream)

    def sync(self):
        self._ck(self.lib.amv_sync(self.ctx))

    def set_option(self, key, value):
        self._ck(self.lib.amv_set_option(self.ctx, key.encode(), int(value)))

    def get_stat(self, key):
        return int(self.lib.amv_get_stat(self.ctx, key.encode()))

    def launch_count(self):
        return int(self.lib.amv_launch_count(self.ctx))

    def qscale_from_quality(self, quality, qmin=2, qmax=31):
        return self.lib.amv_qscale_from_quality(int(quality), qmin, qmax)

    # ---------------------------------------------------------------- raw ABI calls (any memory kind)
    def decode_frames_raw(self, pkts, pkts_bytes, pkt_off, pkt_size, n, w, h, y, u, v, ls_y, ls_c, fs_y, fs_c, status, mem,
                          sp5x=False, mjpeg=False):
        fn = self.lib.amv_decode_frames_sp5x if sp5x else self.lib.amv_decode_frames
        if mjpeg:
            fn = self.lib.amv_decode_frames_mjpeg
        self._ck(fn(self.ctx, _ptr(pkts), pkts_bytes, _ptr(pkt_off), _ptr(pkt_size), n, w, h,
                                            _ptr(y), _ptr(u), _ptr(v), ls_y, ls_c, fs_y, fs_c, _ptr(status), mem))

    def convert_range_raw(self, y, u, v, ls_y, ls_c, fs_y, fs_c, n, w, h, direction, oy, ou, ov, ols_y, ols_c, ofs_y, ofs_c, mem):
        self._ck(self.lib.amv_convert_range(self.ctx, _ptr(y), _ptr(u), _ptr(v), ls_y, ls_c, fs_y, fs_c, n, w, h, direction,
                                            _ptr(oy), _ptr(ou), _ptr(ov), ols_y, ols_c, ofs_y, ofs_c, mem))

    def convert_range(self, y, u, v, direction):
        """numpy planes [n,h,w] / [n,ch,cw]; direction 0: yuv420p -> yuvj420p, 1: yuvj420p -> yuv420p"""
        y, u, v = (np.ascontiguousarray(a, np.uint8) for a in (y, u, v))
        n, h, w = y.shape
        cw, ch = chroma_dims(w, h)
        oy, ou, ov = np.zeros_like(y), np.zeros_like(u), np.zeros_like(v)
        self.convert_range_raw(y, u, v, w, cw, w * h, cw * ch, n, w, h, direction, oy, ou, ov, w, cw, w * h, cw * ch, MEM_HOST)
        return oy, ou, ov

    def scale_frames_raw(self, y, u, v, ls_y, ls_c, fs_y, fs_c, n, iw, ih, oy, ou, ov, ols_y, ols_c, ofs_y, ofs_c, ow, oh, mem,
                         flags=0):
        self._ck(self.lib.amv_scale_frames_ex(self.ctx, _ptr(y), _ptr(u), _ptr(v), ls_y, ls_c, fs_y, fs_c, n, iw, ih,
                                              _ptr(oy), _ptr(ou), _ptr(ov), ols_y, ols_c, ofs_y, ofs_c, ow, oh, flags, mem))

    def scale_frames(self, y, u, v, ow, oh, fill=0, flags=0):
        """numpy planes [n,ih,iw] / [n,ich,icw] -> [n,oh,ow] / [n,och,ocw] as the reference's img_resample scales them
        (chroma at sizes >> 1; output bytes the reference does not write keep `fill`)"""
        y, u, v = (np.ascontiguousarray(a, np.uint8) for a in (y, u, v))
        n, ih, iw = y.shape
        icw, ich = chroma_dims(iw, ih)
        ocw, och = chroma_dims(ow, oh)
        oy = np.full((n, oh, ow), fill, np.uint8)
        ou = np.full((n, och, ocw), fill, np.uint8)
        ov = np.full((n, och, ocw), fill, np.uint8)
        self.scale_frames_raw(y, u, v, iw, icw, iw * ih, icw * ich, n, iw, ih, oy, ou, ov, ow, ocw, ow * oh, ocw * och, ow, oh,
                              MEM_HOST, flags)
        return oy, ou, ov

    def audio_resample_count(self, n_in, in_rate, out_rate=22050):
        return int(self.lib.amv_audio_resample_count(int(n_in), int(in_rate), int(out_rate)))

    def audio_resample_raw(self, pcm, n_in, in_channels, in_rate, out_rate, out, out_cap, mem):
        k = C.c_uint64(0)
        self._ck(self.lib.amv_audio_resample(self.ctx, _ptr(pcm), int(n_in), int(in_channels), int(in_rate), int(out_rate),
                                             _ptr(out), int(out_cap), C.addressof(k), mem))
        return int(k.value)

    def audio_resample_from_raw(self, pcm, in_base, n_in, in_channels, in_rate, out_rate, k_start, out, out_cap, mem):
        k = C.c_uint64(0)
        self._ck(self.lib.amv_audio_resample_from(self.ctx, _ptr(pcm), int(in_base), int(n_in), int(in_channels), int(in_rate),
                                                  int(out_rate), int(k_start), _ptr(out), int(out_cap), C.addressof(k), mem))
        return int(k.value)

    def audio_resample_packets(self, packets, in_channels, in_rate, out_rate=22050):
        """feed a stream packet by packet, keeping only the tail the next outputs still need (what the reference's
        audio_resample carries from call to call); returns the list of per-packet outputs"""
        outs, k_next, base = [], 0, 0
        buf = np.zeros(0, np.int16)
        for pk in packets:
            buf = np.concatenate([buf, np.ascontiguousarray(pk, np.int16).reshape(-1)])
            n_buf = buf.size // in_channels
            cap = max(self.audio_resample_count(base + n_buf, in_rate, out_rate) - k_next, 0)
            out = np.zeros(max(cap, 1), np.int16)
            k = self.audio_resample_from_raw(buf, base, n_buf, in_channels, in_rate, out_rate, k_next, out, cap, MEM_HOST) if n_buf else 0
            outs.append(out[:k])
            k_next += k
            keep_from = max(int(self.lib.amv_audio_resample_first_tap(k_next, in_rate, out_rate)), 0)
            if keep_from > base:
                drop = min(keep_from - base, n_buf)
                buf = buf[drop * in_channels:]
                base += drop
        return outs

    def audio_resample(self, pcm, in_channels, in_rate, out_rate=22050):
        """interleaved int16 samples -> mono int16 at out_rate, as the reference's audio_resample produces over the stream"""
        pcm = np.ascontiguousarray(pcm, np.int16).reshape(-1)
        n_in = pcm.size // in_channels
        cap = self.audio_resample_count(n_in, in_rate, out_rate)
        out = np.zeros(max(cap, 1), np.int16)
        k = self.audio_resample_raw(pcm, n_in, in_channels, in_rate, out_rate, out, cap, MEM_HOST)
        return out[:k]

    def decode_frames_bgr24_raw(self, pkts, pkts_bytes, pkt_off, pkt_size, n, w, h, bgr, line_bytes, frame_stride, status, mem):
        self._ck(self.lib.amv_decode_frames_bgr24(self.ctx, _ptr(pkts), pkts_bytes, _ptr(pkt_off), _ptr(pkt_size), n, w, h,
                                                  _ptr(bgr), line_bytes, frame_stride, _ptr(status), mem))

    def encode_frames_raw(self, y, u, v, ls_y, ls_c, fs_y, fs_c, n, w, h, qscale, out, out_cap, pkt_cap, layout, out_off,
                          out_size, status, mem):
        self._ck(self.lib.amv_encode_frames(self.ctx, _ptr(y), _ptr(u), _ptr(v), ls_y, ls_c, fs_y, fs_c, n, w, h,
                                            _ptr(qscale), _ptr(out), out_cap, pkt_cap, layout, _ptr(out_off),
                                            _ptr(out_size), _ptr(status), mem))

    def adpcm_dec_chunks_raw(self, chunks, chunks_bytes, off, size, n, pcm, pcm_samples, pcm_off, status, mem):
        self._ck(self.lib.amv_adpcm_dec_chunks(self.ctx, _ptr(chunks), chunks_bytes, _ptr(off), _ptr(size), n, _ptr(pcm),
                                               pcm_samples, _ptr(pcm_off), _ptr(status), mem))

    def adpcm_enc_chunks_raw(self, pcm, pcm_samples, pcm_off, nsamples, step_in, step_out, n, out, out_bytes, out_off,
                             status, mem):
        self._ck(self.lib.amv_adpcm_enc_chunks(self.ctx, _ptr(pcm), pcm_samples, _ptr(pcm_off), _ptr(nsamples),
                                               _ptr(step_in), _ptr(step_out), n, _ptr(out), out_bytes, _ptr(out_off),
                                               _ptr(status), mem))

    def adpcm_enc_streams_raw(self, pcm, pcm_samples, pcm_off, nsamples, first_chunk, nstreams, nchunks, step_in, step_out,
                              out, out_bytes, out_off, status, mem):
        self._ck(self.lib.amv_adpcm_enc_streams(self.ctx, _ptr(pcm), pcm_samples, _ptr(pcm_off), _ptr(nsamples),
                                                _ptr(first_chunk), nstreams, nchunks, _ptr(step_in), _ptr(step_out),
                                                _ptr(out), out_bytes, _ptr(out_off), _ptr(status), mem))

    # ---------------------------------------------------------------- host (numpy) convenience wrappers
    def mjpeg_configure(self, jpeg):
        """reads the tables / frame header of one sample JPEG frame (host bytes). -> (w, h)"""
        buf = np.ascontiguousarray(np.frombuffer(bytes(jpeg), np.uint8) if not isinstance(jpeg, np.ndarray) else jpeg, np.uint8)
        w, h = C.c_int(0), C.c_int(0)
        self._ck(self.lib.amv_mjpeg_configure(self.ctx, _ptr(buf), buf.nbytes, C.addressof(w), C.addressof(h)))
        return w.value, h.value

    def decode_frames(self, pkts, pkt_off, pkt_size, w, h, sp5x=False, mjpeg=False):
        """numpy in / numpy out through AMV_MEM_HOST. -> (y[n,h,w], u[n,ch,cw], v[n,ch,cw], status[n]);
        sp5x=True decodes SP5X packets (amv_decode_frames_sp5x); mjpeg=True decodes plain JPEG frames
        (amv_decode_frames_mjpeg, after mjpeg_configure)"""
        pkts = np.ascontiguousarray(pkts, np.uint8)
        pkt_off = np.ascontiguousarray(pkt_off, np.uint64)
        pkt_size = np.ascontiguousarray(pkt_size, np.uint32)
        n = len(pkt_size)
        cw, ch = chroma_dims(w, h)
        if mjpeg:       # 4:2:2 / 4:4:4 frames have larger chroma planes: the configured header says
            cw, ch = self.get_stat("mjpeg_chroma_width"), self.get_stat("mjpeg_chroma_height")
        y = np.zeros((n, h, w), np.uint8)
        u = np.zeros((n, ch, cw), np.uint8)
        v = np.zeros((n, ch, cw), np.uint8)
        st = np.zeros(n, np.int32)
        self.decode_frames_raw(pkts, pkts.nbytes, pkt_off, pkt_size, n, w, h, y, u, v, w, cw, w * h, cw * ch, st, MEM_HOST,
                               sp5x=sp5x, mjpeg=mjpeg)
        return y, u, v, st

    def decode_frames_bgr24(self, pkts, pkt_off, pkt_size, w, h, line_bytes=None):
        """amvlib flavour (AmvVideoDecode): numpy in / numpy out -> (bgr[n,h,line_bytes] bottom-up rows, status[n]);
        line_bytes defaults to the reference's ((w*24+31)//32)*4"""
        pkts = np.ascontiguousarray(pkts, np.uint8)
        pkt_off = np.ascontiguousarray(pkt_off, np.uint64)
        pkt_size = np.ascontiguousarray(pkt_size, np.uint32)
        n = len(pkt_size)
        lb = int(line_bytes or amvlib_line_bytes(w))
        bgr = np.zeros((n, h, lb), np.uint8)
        st = np.zeros(n, np.int32)
        self.decode_frames_bgr24_raw(pkts, pkts.nbytes, pkt_off, pkt_size, n, w, h, bgr, lb, h * lb, st, MEM_HOST)
        return bgr, st

    def encode_frames(self, y, u, v, qscale=None, pkt_cap=None, layout=LAYOUT_PACKED):
        """numpy planes [n,h,w] / [n,ch,cw] -> (packets, off[n], size[n], status[n])"""
        y = np.ascontiguousarray(y, np.uint8)
        u = np.ascontiguousarray(u, np.uint8)
        v = np.ascontiguousarray(v, np.uint8)
        n, h, w = y.shape
        cw, ch = chroma_dims(w, h)
        assert u.shape == (n, ch, cw) and v.shape == (n, ch, cw)
        pkt_cap = int(pkt_cap or (w * h * 3 + 4096))
        q = None if qscale is None else np.ascontiguousarray(np.broadcast_to(np.asarray(qscale, np.int32), (n,)))
        out = np.zeros(n * pkt_cap, np.uint8)
        off = np.zeros(n, np.uint64)
        size = np.zeros(n, np.uint32)
        st = np.zeros(n, np.int32)
        self.encode_frames_raw(y, u, v, w, cw, w * h, cw * ch, n, w, h, q, out, out.nbytes, pkt_cap, layout, off, size, st,
                               MEM_HOST)
        if layout == LAYOUT_PACKED:
            out = out[: int(size.astype(np.uint64).sum())].copy()
        return out, off, size, st

    def adpcm_decode(self, chunks, off, size):
        chunks = np.ascontiguousarray(chunks, np.uint8)
        off = np.ascontiguousarray(off, np.uint64)
        size = np.ascontiguousarray(size, np.uint32)
        n = len(size)
        ns = np.maximum(size.astype(np.int64) - 8, 0) * 2
        poff = offsets_of(ns)
        pcm = np.zeros(max(int(ns.sum()), 1), np.int16)
        st = np.zeros(n, np.int32)
        self.adpcm_dec_chunks_raw(chunks, chunks.nbytes, off, size, n, pcm, int(ns.sum()), poff, st, MEM_HOST)
        return pcm[: int(ns.sum())], poff, st

    def adpcm_encode(self, pcm, pcm_off, nsamples, step_in=None):
        pcm = np.ascontiguousarray(pcm, np.int16)
        pcm_off = np.ascontiguousarray(pcm_off, np.uint64)
        nsamples = np.ascontiguousarray(nsamples, np.uint32)
        n = len(nsamples)
        osz = 8 + nsamples.astype(np.uint64) // 2
        ooff = offsets_of(osz)
        out = np.zeros(int(osz.sum()), np.uint8)
        si = None if step_in is None else np.ascontiguousarray(step_in, np.int16)
        so = np.zeros(n, np.int16)
        st = np.zeros(n, np.int32)
        self.adpcm_enc_chunks_raw(pcm, len(pcm), pcm_off, nsamples, si, so, n, out, out.nbytes, ooff, st, MEM_HOST)
        return out, ooff, osz.astype(np.uint32), so, st

    def adpcm_encode_streams(self, pcm, pcm_off, nsamples, first_chunk, step_in=None):
        pcm = np.ascontiguousarray(pcm, np.int16)
        pcm_off = np.ascontiguousarray(pcm_off, np.uint64)
        nsamples = np.ascontiguousarray(nsamples, np.uint32)
        first_chunk = np.ascontiguousarray(first_chunk, np.uint32)
        nchunks, nstreams = len(nsamples), len(first_chunk) - 1
        osz = 8 + nsamples.astype(np.uint64) // 2
        ooff = offsets_of(osz)
        out = np.zeros(int(osz.sum()), np.uint8)
        si = None if step_in is None else np.ascontiguousarray(step_in, np.int16)
        so = np.zeros(nstreams, np.int16)
        st = np.zeros(nchunks, np.int32)
        self.adpcm_enc_streams_raw(pcm, len(pcm), pcm_off, nsamples, first_chunk, nstreams, nchunks, si, so, out, out.nbytes,
                                   ooff, st, MEM_HOST)
        return out, ooff, osz.astype(np.uint32), so, st
