/*
 * TEST INFRASTRUCTURE ONLY -- never linked into, or called by, the product.
 *
 * Thin batch driver around the UNMODIFIED reference codecs (AMVmuxer FFmpeg
 * fork) compiled in place by oracle/build_ref.sh into oracle/_ref/libamvref.so.
 * It goes through the reference's own public plug-in API
 * (avcodec_open / avcodec_encode_video / avcodec_decode_video /
 * avcodec_encode_audio / avcodec_decode_audio2: libavcodec/utils.c:835-976)
 * with the reference's own AVCodec instances
 * (amv_decoder sp5xdec.c:203, amv_encoder mjpegenc.c:485,
 *  adpcm_ima_amv_{en,de}coder adpcm.c:1535) -- one frame / chunk per call,
 * exactly like ffmpeg.c:1083/814/1062/522 does.
 *
 * Used by tests/ (to pin oracle/amv_oracle.c and to generate tests/golden/)
 * and by bench.py's cpu_baseline / --impl reference arm.
 */
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include "avcodec.h"

extern AVCodec amv_decoder, amv_encoder;
extern AVCodec adpcm_ima_amv_decoder, adpcm_ima_amv_encoder;

static int g_inited;
static void ref_init(void)
{
    if (!g_inited) {
        avcodec_init();
        av_log_set_level(AV_LOG_QUIET);
        g_inited = 1;
    }
}

const char *amvref_version(void) { return "amv-codec-tools AMVmuxer libavcodec " AV_STRINGIFY(LIBAVCODEC_VERSION) " generic-C"; }

/* ---- video encode: n frames, planes tightly packed (Y w*h, Cb/Cr cw*ch) ----
 * quality = AVFrame.quality (lambda; 0 => reference default qscale 2).
 * Packets are written back to back into out[0..cap); off[i]/size[i] locate them.
 * Returns number of frames encoded, or a negative error. */
static int encode_frames_with(AVCodec *codec, int pix_fmt, const uint8_t *y, const uint8_t *u, const uint8_t *v,
                              int n, int w, int h, int quality,
                              uint8_t *out, uint64_t *off, uint32_t *size, uint64_t cap);

int amvref_encode_frames(const uint8_t *y, const uint8_t *u, const uint8_t *v,
                         int n, int w, int h, int quality,
                         uint8_t *out, uint64_t *off, uint32_t *size, uint64_t cap)
{
    return encode_frames_with(&amv_encoder, PIX_FMT_YUVJ420P, y, u, v, n, w, h, quality, out, off, size, cap);
}

/* the plain MJPEG codecs of the same source files (mjpegenc.c:474-483, mjpegdec.c:1361-1372): full JPEG
 * frames with their tables in the stream, top-down pictures */
extern AVCodec mjpeg_encoder, mjpeg_decoder;
/* chroma_rows == h: the planes are YUVJ422P (chroma ceil(w/2) x h), else YUVJ420P */
int amvref_mjpeg_encode_frames(const uint8_t *y, const uint8_t *u, const uint8_t *v,
                               int n, int w, int h, int chroma_rows, int quality,
                               uint8_t *out, uint64_t *off, uint32_t *size, uint64_t cap)
{
    return encode_frames_with(&mjpeg_encoder, chroma_rows == h ? PIX_FMT_YUVJ422P : PIX_FMT_YUVJ420P, y, u, v, n, w, h, quality,
                              out, off, size, cap);
}

static int encode_frames_with(AVCodec *codec, int pix_fmt, const uint8_t *y, const uint8_t *u, const uint8_t *v,
                              int n, int w, int h, int quality,
                              uint8_t *out, uint64_t *off, uint32_t *size, uint64_t cap)
{
    ref_init();
    AVCodecContext *c = avcodec_alloc_context();
    AVFrame *pic = avcodec_alloc_frame();
    int cw = (w + 1) >> 1, ch = pix_fmt == PIX_FMT_YUVJ422P ? h : (h + 1) >> 1, i, ret = 0;
    int bufsz = w * h * 6 + 262144;
    uint8_t *buf = av_malloc(bufsz);
    uint64_t pos = 0;
    c->width = w; c->height = h;
    c->time_base.num = 1; c->time_base.den = 16;
    c->pix_fmt = pix_fmt;
    if (avcodec_open(c, codec) < 0) { ret = -2; goto done; }
    for (i = 0; i < n; i++) {
        /* amv_encode_picture mutates data[]/linesize[] (mjpegenc.c:467-470): refill every frame */
        pic->data[0] = (uint8_t *)y + (size_t)i * w * h;
        pic->data[1] = (uint8_t *)u + (size_t)i * cw * ch;
        pic->data[2] = (uint8_t *)v + (size_t)i * cw * ch;
        pic->linesize[0] = w; pic->linesize[1] = cw; pic->linesize[2] = cw;
        pic->quality = quality;
        pic->pts = i;
        int sz = avcodec_encode_video(c, buf, bufsz, pic);
        if (sz < 0) { ret = -3; goto close; }
        if (pos + sz > cap) { ret = -4; goto close; }
        memcpy(out + pos, buf, sz);
        off[i] = pos; size[i] = sz; pos += sz;
    }
    ret = n;
close:
    avcodec_close(c);
done:
    av_free(buf); av_free(pic); av_free(c);
    return ret;
}

/* ---- video decode: n packets -> tightly packed planes ----
 * got[i] receives got_picture; ret_bytes[i] the decoder's return value (may be NULL). */
extern AVCodec sp5x_decoder;
static int decode_frames_with(AVCodec *codec, const uint8_t *pkts, const uint64_t *off, const uint32_t *size,
                              int n, int w, int h,
                              uint8_t *y, uint8_t *u, uint8_t *v, int *got, int *ret_bytes);

int amvref_decode_frames(const uint8_t *pkts, const uint64_t *off, const uint32_t *size,
                         int n, int w, int h,
                         uint8_t *y, uint8_t *u, uint8_t *v, int *got, int *ret_bytes)
{
    return decode_frames_with(&amv_decoder, pkts, off, size, n, w, h, y, u, v, got, ret_bytes);
}

/* same driver on the sibling codec of the same source file: sp5x_decoder (sp5xdec.c:190-201) */
int amvref_sp5x_decode_frames(const uint8_t *pkts, const uint64_t *off, const uint32_t *size,
                              int n, int w, int h,
                              uint8_t *y, uint8_t *u, uint8_t *v, int *got, int *ret_bytes)
{
    return decode_frames_with(&sp5x_decoder, pkts, off, size, n, w, h, y, u, v, got, ret_bytes);
}

int amvref_mjpeg_decode_frames(const uint8_t *pkts, const uint64_t *off, const uint32_t *size,
                               int n, int w, int h,
                               uint8_t *y, uint8_t *u, uint8_t *v, int *got, int *ret_bytes)
{
    return decode_frames_with(&mjpeg_decoder, pkts, off, size, n, w, h, y, u, v, got, ret_bytes);
}

static int decode_frames_with(AVCodec *codec, const uint8_t *pkts, const uint64_t *off, const uint32_t *size,
                              int n, int w, int h,
                              uint8_t *y, uint8_t *u, uint8_t *v, int *got, int *ret_bytes)
{
    ref_init();
    AVCodecContext *c = avcodec_alloc_context();
    AVFrame *pic = avcodec_alloc_frame();
    int cw = (w + 1) >> 1, ch = (h + 1) >> 1, i, r, ret = 0;
    uint32_t maxsz = 0;
    for (i = 0; i < n; i++) if (size[i] > maxsz) maxsz = size[i];
    uint8_t *buf = av_mallocz(maxsz + FF_INPUT_BUFFER_PADDING_SIZE + 16);
    c->width = w; c->height = h;          /* container supplies dims: avidec.c:429-434 */
    c->coded_width = w; c->coded_height = h;
    if (avcodec_open(c, codec) < 0) { ret = -2; goto done; }
    for (i = 0; i < n; i++) {
        int g = 0;
        memcpy(buf, pkts + off[i], size[i]);
        memset(buf + size[i], 0, FF_INPUT_BUFFER_PADDING_SIZE);
        r = avcodec_decode_video(c, pic, &g, buf, size[i]);
        if (got) got[i] = g;
        if (ret_bytes) ret_bytes[i] = r;
        if (r < 0 || !g) continue;
        /* chroma plane size by what the decoder says it decoded (mjpegdec.c:283-311); planes are packed at that size */
        if (c->pix_fmt == PIX_FMT_YUVJ422P) { ch = h; }
        else if (c->pix_fmt == PIX_FMT_YUVJ444P) { ch = h; cw = w; }
        for (r = 0; r < h; r++)
            memcpy(y + (size_t)i * w * h + (size_t)r * w, pic->data[0] + r * pic->linesize[0], w);
        for (r = 0; r < ch; r++) {
            memcpy(u + (size_t)i * cw * ch + (size_t)r * cw, pic->data[1] + r * pic->linesize[1], cw);
            memcpy(v + (size_t)i * cw * ch + (size_t)r * cw, pic->data[2] + r * pic->linesize[2], cw);
        }
    }
    ret = n;
    avcodec_close(c);
done:
    av_free(buf); av_free(pic); av_free(c);
    return ret;
}

/* ---- ADPCM decode: n chunks -> pcm; nsamp[i] = samples produced ---- */
int amvref_adpcm_decode(const uint8_t *chunks, const uint64_t *off, const uint32_t *size,
                        int n, int16_t *pcm, const uint64_t *pcm_off, uint32_t *nsamp)
{
    ref_init();
    AVCodecContext *c = avcodec_alloc_context();
    int i, ret = 0;
    int16_t *tmp = av_malloc(AVCODEC_MAX_AUDIO_FRAME_SIZE * 2);
    uint8_t *buf = av_mallocz(65536 + 16);
    c->channels = 1; c->sample_rate = 22050;
    if (avcodec_open(c, &adpcm_ima_amv_decoder) < 0) { ret = -2; goto done; }
    for (i = 0; i < n; i++) {
        int bytes = AVCODEC_MAX_AUDIO_FRAME_SIZE * 2;
        if (size[i] > 65536) { ret = -5; break; }
        memcpy(buf, chunks + off[i], size[i]);
        int r = avcodec_decode_audio2(c, tmp, &bytes, buf, size[i]);
        if (r < 0) { nsamp[i] = 0; continue; }
        /* adpcm_decode_frame sets *data_size = (uint8_t*)samples - (uint8_t*)data */
        nsamp[i] = bytes / 2;
        memcpy(pcm + pcm_off[i], tmp, bytes);
    }
    if (!ret) ret = n;
    avcodec_close(c);
done:
    av_free(tmp); av_free(buf); av_free(c);
    return ret;
}

/* ---- ADPCM encode of ONE continuous stream, chunk by chunk (state chained,
 * adpcm.c:461-496).  frame_size = samples per call as the AMV muxer sets it
 * (amvenc.c:276-281).  Every call consumes 2n samples where 2n is what the
 * encoder wrote into the chunk header (bytes 4..7); the harness advances by
 * that (ffmpeg.c itself advances by frame_size -- SURVEY §9.10).
 * Returns number of chunks written. consumed[i] = samples consumed by chunk i. */
int amvref_adpcm_encode_stream_trellis(const int16_t *pcm, uint64_t total_samples, int frame_size, int trellis,
                                       uint8_t *out, uint64_t *off, uint32_t *size,
                                       uint32_t *consumed, int max_chunks, uint64_t cap);
int amvref_adpcm_encode_stream(const int16_t *pcm, uint64_t total_samples, int frame_size,
                               uint8_t *out, uint64_t *off, uint32_t *size,
                               uint32_t *consumed, int max_chunks, uint64_t cap)
{
    return amvref_adpcm_encode_stream_trellis(pcm, total_samples, frame_size, 0, out, off, size, consumed, max_chunks, cap);
}

/* trellis > 0: the `-trellis N` path of the same encoder (adpcm_compress_trellis, adpcm.c:287-443) */
int amvref_adpcm_encode_stream_trellis(const int16_t *pcm, uint64_t total_samples, int frame_size, int trellis,
                                       uint8_t *out, uint64_t *off, uint32_t *size,
                                       uint32_t *consumed, int max_chunks, uint64_t cap)
{
    ref_init();
    AVCodecContext *c = avcodec_alloc_context();
    int k = 0;
    uint64_t pos = 0, opos = 0;
    uint8_t *buf = av_malloc(FF_MIN_BUFFER_SIZE + 4 * frame_size + 65536);
    c->channels = 1; c->sample_rate = 22050; c->frame_size = frame_size;
    c->trellis = trellis;
    if (avcodec_open(c, &adpcm_ima_amv_encoder) < 0) { k = -2; goto done; }
    c->frame_size = frame_size;
    while (k < max_chunks && pos + 2 * (uint64_t)frame_size + 2 <= total_samples) {
        int sz = avcodec_encode_audio(c, buf, FF_MIN_BUFFER_SIZE + 4 * frame_size + 65536, pcm + pos);
        if (sz < 8) { k = -3; break; }
        if (opos + sz > cap) { k = -4; break; }
        uint32_t two_n = buf[4] | (buf[5] << 8) | (buf[6] << 16) | ((uint32_t)buf[7] << 24);
        memcpy(out + opos, buf, sz);
        off[k] = opos; size[k] = sz; consumed[k] = two_n;
        opos += sz; pos += two_n; k++;
    }
    avcodec_close(c);
done:
    av_free(buf); av_free(c);
    return k;
}

/* ---- single-stage entry points, for pinning the oracle stage by stage ---- */
void ff_jpeg_fdct_islow(int16_t *data);
void simple_idct_put(uint8_t *dest, int line_size, int16_t *block);
void amvref_fdct_islow(int16_t *blocks, int nblocks)
{
    int i; for (i = 0; i < nblocks; i++) ff_jpeg_fdct_islow(blocks + 64 * i);
}
void amvref_simple_idct_put(const int16_t *blocks, int nblocks, uint8_t *dest /* 64 B per block */)
{
    int i; int16_t tmp[64];
    ref_init();                       /* fills ff_cropTbl (dsputil.c:3812-3820) */
    for (i = 0; i < nblocks; i++) {
        memcpy(tmp, blocks + 64 * i, sizeof(tmp));
        simple_idct_put(dest + 64 * i, 8, tmp);
    }
}

/* ---- pre/post stage next to the codec: the range conversion ffmpeg.c inserts through img_convert
 * when the source is yuv420p (CCIR range) and the AMV encoder wants yuvj420p (ffmpeg.c:do_video_out ->
 * img_convert, imgconvert.c:2379-2510 -> img_apply_table with y/c_ccir_to_jpeg, :1216-1260; colorspace.h:69-84).
 * dir 0: YUV420P -> YUVJ420P, dir 1: YUVJ420P -> YUV420P.  Tight planes. */
int amvref_convert_range(const uint8_t *y, const uint8_t *u, const uint8_t *v, int n, int w, int h, int dir,
                         uint8_t *oy, uint8_t *ou, uint8_t *ov)
{
    ref_init();
    int cw = (w + 1) >> 1, ch = (h + 1) >> 1, i;
    for (i = 0; i < n; i++) {
        AVPicture src, dst;
        src.data[0] = (uint8_t *)y + (size_t)i * w * h; src.data[1] = (uint8_t *)u + (size_t)i * cw * ch;
        src.data[2] = (uint8_t *)v + (size_t)i * cw * ch; src.data[3] = NULL;
        src.linesize[0] = w; src.linesize[1] = cw; src.linesize[2] = cw; src.linesize[3] = 0;
        dst.data[0] = oy + (size_t)i * w * h; dst.data[1] = ou + (size_t)i * cw * ch;
        dst.data[2] = ov + (size_t)i * cw * ch; dst.data[3] = NULL;
        dst.linesize[0] = w; dst.linesize[1] = cw; dst.linesize[2] = cw; dst.linesize[3] = 0;
        if (img_convert(&dst, dir ? PIX_FMT_YUV420P : PIX_FMT_YUVJ420P, &src, dir ? PIX_FMT_YUVJ420P : PIX_FMT_YUV420P, w, h) < 0)
            return -1;
    }
    return n;
}

/* ---- pre stages next to the codec: the picture scaler ffmpeg.c reaches through sws_getContext /
 * sws_scale (imgresample.c:515-690 -> img_resample_init / img_resample, :433-507) for `-s WxH`, and the audio
 * resampler of do_audio_out (ffmpeg.c:501-505 -> audio_resample, resample.c:131-235 -> av_resample, resample2.c).
 * Tight planes; chroma planes are stored with the rounded-up size the rest of the harness uses, the reference
 * itself touches only (w>>1) x (h>>1) of them. */
int amvref_img_resample(const uint8_t *y, const uint8_t *u, const uint8_t *v, int n, int iw, int ih, int ow, int oh,
                        uint8_t *oy, uint8_t *ou, uint8_t *ov)
{
    ref_init();
    int icw = (iw + 1) >> 1, ich = (ih + 1) >> 1, ocw = (ow + 1) >> 1, och = (oh + 1) >> 1, i;
    ImgReSampleContext *s = img_resample_init(ow, oh, iw, ih);
    if (!s) return -1;
    for (i = 0; i < n; i++) {
        AVPicture src, dst;
        src.data[0] = (uint8_t *)y + (size_t)i * iw * ih; src.data[1] = (uint8_t *)u + (size_t)i * icw * ich;
        src.data[2] = (uint8_t *)v + (size_t)i * icw * ich; src.data[3] = NULL;
        src.linesize[0] = iw; src.linesize[1] = icw; src.linesize[2] = icw; src.linesize[3] = 0;
        dst.data[0] = oy + (size_t)i * ow * oh; dst.data[1] = ou + (size_t)i * ocw * och;
        dst.data[2] = ov + (size_t)i * ocw * och; dst.data[3] = NULL;
        dst.linesize[0] = ow; dst.linesize[1] = ocw; dst.linesize[2] = ocw; dst.linesize[3] = 0;
        img_resample(s, &dst, &src);
    }
    img_resample_close(s);
    return n;
}

/* in: interleaved int16, n_in samples per channel, fed to audio_resample `chunk` samples per call (as ffmpeg.c
 * feeds it one decoded packet at a time); output mono.  Returns the samples written, -1 on error. */
int64_t amvref_audio_resample(const int16_t *in, int64_t n_in, int in_ch, int in_rate, int out_rate, int chunk,
                              int16_t *out, int64_t out_cap)
{
    ref_init();
    ReSampleContext *s = audio_resample_init(1, in_ch, out_rate, in_rate);
    if (!s) return -1;
    int64_t done = 0, pos = 0;
    int lenmax = (int)(4.0 * chunk * ((double)out_rate / in_rate + 1.0)) + 64;
    short *tmp = av_malloc(sizeof(short) * lenmax);
    while (pos < n_in) {
        int nb = (int)(n_in - pos < chunk ? n_in - pos : chunk);
        int k = audio_resample(s, tmp, (short *)in + pos * in_ch, nb);
        if (k < 0 || done + k > out_cap) { done = -1; break; }
        memcpy(out + done, tmp, sizeof(short) * k);
        done += k; pos += nb;
    }
    av_free(tmp);
    audio_resample_close(s);
    return done;
}

/* the polyphase bank av_resample_init builds (resample2.c:185-206 -> av_build_filter :93-141): `len` receives
 * filter_length, bank receives filter_length * ((1<<phase_shift) + 1) coefficients */
struct AmvrefResampleCtxView { int16_t *filter_bank; int filter_length; };
int amvref_resample_bank(int out_rate, int in_rate, int16_t *bank, int cap, int *len)
{
    struct AVResampleContext *c = av_resample_init(out_rate, in_rate, 16, 10, 0, 0.8);
    struct AmvrefResampleCtxView *vw = (struct AmvrefResampleCtxView *)c;
    int total = vw->filter_length * 1025;
    *len = vw->filter_length;
    if (total > cap) { av_resample_close(c); return -1; }
    memcpy(bank, vw->filter_bank, sizeof(int16_t) * total);
    av_resample_close(c);
    return total;
}

/* the same through the fork's sws_getContext / sws_scale (imgresample.c:515-690), the entry ffmpeg.c uses: with
 * jpeg_in / jpeg_out the source / destination format is PIX_FMT_YUVJ420P instead of PIX_FMT_YUV420P, which makes
 * sws_scale wrap the scaler in img_convert steps.  Tight planes, even sizes. */
#include "swscale.h"
int amvref_sws_scale(const uint8_t *y, const uint8_t *u, const uint8_t *v, int n, int iw, int ih, int ow, int oh,
                     int jpeg_in, int jpeg_out, uint8_t *oy, uint8_t *ou, uint8_t *ov)
{
    ref_init();
    int icw = (iw + 1) >> 1, ich = (ih + 1) >> 1, ocw = (ow + 1) >> 1, och = (oh + 1) >> 1, i;
    struct SwsContext *c = sws_getContext(iw, ih, jpeg_in ? PIX_FMT_YUVJ420P : PIX_FMT_YUV420P, ow, oh,
                                          jpeg_out ? PIX_FMT_YUVJ420P : PIX_FMT_YUV420P, SWS_BICUBIC, NULL, NULL, NULL);
    if (!c) return -1;
    for (i = 0; i < n; i++) {
        uint8_t *src[4] = { (uint8_t *)y + (size_t)i * iw * ih, (uint8_t *)u + (size_t)i * icw * ich, (uint8_t *)v + (size_t)i * icw * ich, NULL };
        uint8_t *dst[4] = { oy + (size_t)i * ow * oh, ou + (size_t)i * ocw * och, ov + (size_t)i * ocw * och, NULL };
        int sst[4] = { iw, icw, icw, 0 }, dstr[4] = { ow, ocw, ocw, 0 };
        if (sws_scale(c, src, sst, 0, ih, dst, dstr) < 0) { sws_freeContext(c); return -2; }
    }
    sws_freeContext(c);
    return n;
}
