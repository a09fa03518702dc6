/*
 * amvcuda_codecs.c -- reference-side binding: AVCodec instances that drop libamvcuda into the
 * AMVmuxer FFmpeg fork (libavcodec 51.47.1) in place of its CPU codecs.
 *
 * This file is compiled INSIDE the reference tree's include path (it needs avcodec.h); it is
 * the stub a maintainer adds next to sp5xdec.c / mjpegenc.c / adpcm.c.  It holds no codec
 * arithmetic: every callback translates AVCodecContext / AVFrame / packet fields and calls the
 * C ABI of include/amvcuda.h with n = 1.
 *
 *   amvcuda_amv_decoder            replaces amv_decoder            (sp5xdec.c:203-212)
 *   amvcuda_sp5x_decoder           replaces sp5x_decoder           (sp5xdec.c:190-201; same callback, other framing)
 *   amvcuda_mjpeg_decoder          replaces mjpeg_decoder          (mjpegdec.c:1356-1367; baseline frames, 4:2:0 / 4:2:2 / 4:4:4)
 *   amvcuda_amv_encoder            replaces amv_encoder            (mjpegenc.c:485-494)
 *   amvcuda_adpcm_ima_amv_decoder  replaces adpcm_ima_amv_decoder  (adpcm.c:1535)
 *   amvcuda_adpcm_ima_amv_encoder  replaces adpcm_ima_amv_encoder  (adpcm.c:1535)
 *
 *   amvcuda_amv_encoder_delay / amvcuda_amv_decoder_delay: the same two codecs with a look-ahead queue
 *                                  (CODEC_CAP_DELAY, avcodec.h:446): the library is a BATCH codec and a call per frame
 *                                  pays a fixed host cost; these queue AMVCUDA_LOOKAHEAD (default 64) frames / packets,
 *                                  hand them to libamvcuda in one call and return results one per call, the encoder
 *                                  answering 0 bytes and the decoder got_picture = 0 while the queue fills
 *                                  (utils.c:910,938 call a delayed codec with NULL / no data to drain it, ffmpeg.c:1088
 *                                  and its end-of-stream flush handle both).  Registered instead of the plain ones when
 *                                  the environment sets AMVCUDA_LOOKAHEAD.
 *
 * avcodec_find_decoder/encoder return the FIRST registered match (utils.c:1023-1056), so
 * amvcuda_register_codecs() must run before avcodec_register_all() -- or the four
 * REGISTER_ENCDEC lines of allcodecs.c:64,255 are pointed at these symbols.
 */
#include <stdlib.h>
#include <string.h>
#include "avcodec.h"
#include "amvcuda.h"

/* ------------------------------------------------------------------------------ video decode */
typedef struct AmvCudaVideoDec {
    amv_ctx *h;
    AVFrame picture;
    int configured;                      /* plain MJPEG: amv_mjpeg_configure has seen a frame */
} AmvCudaVideoDec;

static int amvcuda_dec_init(AVCodecContext *avctx)
{
    AmvCudaVideoDec *c = avctx->priv_data;
    avctx->pix_fmt = PIX_FMT_YUVJ420P;                     /* mjpegdec.c:311-312 for 0x221111 */
    return amv_create(NULL, &c->h) == AMV_OK ? 0 : -1;     /* no device => the codec cannot open: no CPU path */
}

static int amvcuda_dec_close(AVCodecContext *avctx)
{
    AmvCudaVideoDec *c = avctx->priv_data;
    if (c->picture.data[0]) avctx->release_buffer(avctx, &c->picture);
    amv_destroy(c->h);
    c->h = NULL;
    return 0;
}

/* same contract as sp5x_decode_frame (sp5xdec.c:33-93) + the EOI branch of
 * ff_mjpeg_decode_frame (mjpegdec.c:1271-1297): decoder-owned picture from get_buffer, released
 * on the next call (:327-328); key frame, I type, quality = FF_QP2LAMBDA * max qscale. */
static int amvcuda_dec_frame(AVCodecContext *avctx, void *data, int *data_size, uint8_t *buf, int buf_size)
{
    AmvCudaVideoDec *c = avctx->priv_data;
    AVFrame *out = data;
    const int w = avctx->width, h = avctx->height;
    uint64_t off = 0;
    uint32_t size = (uint32_t)buf_size;
    int32_t status = 0;

    if (!w || !h) return -1;                               /* sp5xdec.c:45-46 */
    avctx->pix_fmt = PIX_FMT_YUVJ420P;
    if (c->picture.data[0]) avctx->release_buffer(avctx, &c->picture);
    c->picture.reference = 0;
    if (avctx->get_buffer(avctx, &c->picture) < 0) return -1;
    c->picture.pict_type = FF_I_TYPE;
    c->picture.key_frame = 1;
    /* one callback serves both codecs, like sp5x_decode_frame: CODEC_ID_AMV takes the flipped, stuffed framing,
     * CODEC_ID_SP5X the 14-byte header + literal bytes (sp5xdec.c:75-84); avcodec_open set codec_id (utils.c:862) */
    if ((avctx->codec_id == CODEC_ID_SP5X ? amv_decode_frames_sp5x : amv_decode_frames)(
            c->h, buf, (uint64_t)buf_size, &off, &size, 1, w, h,
            c->picture.data[0], c->picture.data[1], c->picture.data[2],
            c->picture.linesize[0], c->picture.linesize[1],
            (uint64_t)c->picture.linesize[0] * h, (uint64_t)c->picture.linesize[1] * ((h + 1) / 2),
            &status, AMV_MEM_HOST) != AMV_OK)
        return -1;
    /* scan errors are swallowed by the reference too (mjpegdec.c:1300): the picture is returned */
    *out = c->picture;
    /* qscale[i] = max(q[1], q[8]) >> 1 of the two fixed tables (mjpegdec.c:137-139): 10>>1 and 19>>1 */
    out->quality = 9 * FF_QP2LAMBDA;
    *data_size = sizeof(AVFrame);
    return buf_size;
}

AVCodec amvcuda_amv_decoder = {
    "amv", CODEC_TYPE_VIDEO, CODEC_ID_AMV, sizeof(AmvCudaVideoDec),
    amvcuda_dec_init, NULL, amvcuda_dec_close, amvcuda_dec_frame,
};

AVCodec amvcuda_sp5x_decoder = {
    "sp5x", CODEC_TYPE_VIDEO, CODEC_ID_SP5X, sizeof(AmvCudaVideoDec),
    amvcuda_dec_init, NULL, amvcuda_dec_close, amvcuda_dec_frame,
};

/* plain MJPEG: ff_mjpeg_decode_frame (mjpegdec.c:1106-1340) for baseline 4:2:0 frames.  The frame's own marker
 * segments say how to decode it: amv_mjpeg_configure reads them from the first packet and again whenever a packet
 * does not fit the current configuration (AMV_ST_HEADER); the picture size comes from SOF0 as in
 * ff_mjpeg_decode_sof (:215-251, avcodec_set_dimensions).  quality = FF_QP2LAMBDA * max over the DQT tables of
 * max(q[1], q[8]) >> 1 (:137-139, :1283-1286). */
static int amvcuda_mjpeg_frame(AVCodecContext *avctx, void *data, int *data_size, uint8_t *buf, int buf_size)
{
    AmvCudaVideoDec *c = avctx->priv_data;
    AVFrame *out = data;
    uint64_t off = 0;
    uint32_t size = (uint32_t)buf_size;
    int32_t status = 0;
    int w = avctx->width, h = avctx->height, attempt, i, qmax = 0;

    for (attempt = 0; attempt < 2; attempt++) {
        if (attempt == 1 || !c->configured) {
            if (amv_mjpeg_configure(c->h, buf, size, &w, &h) != AMV_OK) return -1;      /* not a frame this path covers */
            c->configured = 1;
            if (w != avctx->width || h != avctx->height) avcodec_set_dimensions(avctx, w, h);
        }
        {   /* pix_fmt from the sampling, as ff_mjpeg_decode_sof does (mjpegdec.c:283-311) */
            const int cw = (int)amv_get_stat(c->h, "mjpeg_chroma_width"), ch = (int)amv_get_stat(c->h, "mjpeg_chroma_height");
            avctx->pix_fmt = ch < h ? PIX_FMT_YUVJ420P : (cw < w ? PIX_FMT_YUVJ422P : PIX_FMT_YUVJ444P);
            if (c->picture.data[0]) avctx->release_buffer(avctx, &c->picture);
            c->picture.reference = 0;
            if (avctx->get_buffer(avctx, &c->picture) < 0) return -1;
            c->picture.pict_type = FF_I_TYPE;
            c->picture.key_frame = 1;
            if (amv_decode_frames_mjpeg(c->h, buf, (uint64_t)buf_size, &off, &size, 1, w, h,
                                        c->picture.data[0], c->picture.data[1], c->picture.data[2],
                                        c->picture.linesize[0], c->picture.linesize[1],
                                        (uint64_t)c->picture.linesize[0] * h, (uint64_t)c->picture.linesize[1] * ch,
                                        &status, AMV_MEM_HOST) != AMV_OK)
                return -1;
        }
        if (!(status & AMV_ST_HEADER)) break;
        if (attempt == 1) return -1;
    }
    for (i = 2; i + 69 <= buf_size && i < 4096; i++)                  /* the DQT segments in front of the scan */
        if (buf[i] == 0xff && buf[i + 1] == 0xdb) {
            int len = (buf[i + 2] << 8) | buf[i + 3], k;
            for (k = i + 4; k + 65 <= i + 2 + len && k + 65 <= buf_size; k += 65) {
                int q = (buf[k + 2] > buf[k + 9] ? buf[k + 2] : buf[k + 9]) >> 1;
                if (q > qmax) qmax = q;
            }
            i += 1 + len;
        } else if (buf[i] == 0xff && buf[i + 1] == 0xda) break;
    *out = c->picture;
    out->quality = qmax * FF_QP2LAMBDA;
    *data_size = sizeof(AVFrame);
    return buf_size;
}

AVCodec amvcuda_mjpeg_decoder = {
    "mjpeg", CODEC_TYPE_VIDEO, CODEC_ID_MJPEG, sizeof(AmvCudaVideoDec),
    amvcuda_dec_init, NULL, amvcuda_dec_close, amvcuda_mjpeg_frame,
};

/* ------------------------------------------------------------------------------ video encode */
typedef struct AmvCudaVideoEnc {
    amv_ctx *h;
    AVFrame coded;
} AmvCudaVideoEnc;

static int amvcuda_enc_init(AVCodecContext *avctx)
{
    AmvCudaVideoEnc *c = avctx->priv_data;
    /* what MPV_encode_init accepts for CODEC_ID_AMV (mpegvideo_enc.c:249-257) minus 4:2:2, which the
     * reference's own decoder cannot read; options outside the contract are refused */
    if (avctx->pix_fmt != PIX_FMT_YUVJ420P && avctx->pix_fmt != PIX_FMT_YUV420P) return -1;
    if (avctx->thread_count > 1 || avctx->trellis || avctx->intra_dc_precision || (avctx->flags & CODEC_FLAG_GRAY)) return -1;
    if (!avctx->time_base.num || !avctx->time_base.den) return -1;    /* mpegvideo_enc.c:461-464 */
    avctx->coded_frame = &c->coded;
    avctx->delay = 0;
    return amv_create(NULL, &c->h) == AMV_OK ? 0 : -1;
}

static int amvcuda_enc_close(AVCodecContext *avctx)
{
    AmvCudaVideoEnc *c = avctx->priv_data;
    amv_destroy(c->h);
    c->h = NULL;
    return 0;
}

/* amv_encode_picture (mjpegenc.c:454-472): one frame in, one packet out, delay 0.
 * The reference flips pic->data/linesize in place; callers never rely on that, we leave pic alone. */
static int amvcuda_enc_frame(AVCodecContext *avctx, uint8_t *buf, int buf_size, void *data)
{
    AmvCudaVideoEnc *c = avctx->priv_data;
    AVFrame *pic = data;
    const int w = avctx->width, h = avctx->height;
    int32_t qscale, status = 0;
    uint64_t off = 0;
    uint32_t size = 0;

    if (avctx->flags & CODEC_FLAG_EMU_EDGE) return -1;                       /* mjpegenc.c:463-464 */
    qscale = amv_qscale_from_quality(pic->quality, avctx->qmin, avctx->qmax);  /* update_qscale */
    if (amv_encode_frames(c->h, pic->data[0], pic->data[1], pic->data[2], pic->linesize[0], pic->linesize[1],
                          (uint64_t)pic->linesize[0] * h, (uint64_t)pic->linesize[1] * ((h + 1) / 2),
                          1, w, h, &qscale, buf, (uint64_t)buf_size, (uint32_t)buf_size, AMV_LAYOUT_SLOTS,
                          &off, &size, &status, AMV_MEM_HOST) != AMV_OK)
        return -1;
    if (status) return -1;                                  /* "encoded frame too large" (mpegvideo_enc.c:2077-2080) */
    c->coded.key_frame = 1;
    c->coded.pict_type = FF_I_TYPE;
    c->coded.quality = pic->quality;
    c->coded.pts = pic->pts;
    return (int)size;
}

static const enum PixelFormat amvcuda_pix_fmts[] = { PIX_FMT_YUVJ420P, -1 };

AVCodec amvcuda_amv_encoder = {
    "amv", CODEC_TYPE_VIDEO, CODEC_ID_AMV, sizeof(AmvCudaVideoEnc),
    amvcuda_enc_init, amvcuda_enc_frame, amvcuda_enc_close, NULL,
    .pix_fmts = amvcuda_pix_fmts,
};


/* ------------------------------------------------------------------- look-ahead (delayed) video codecs */
static int amvcuda_lookahead(void)
{
    const char *e = getenv("AMVCUDA_LOOKAHEAD");
    int k = e ? atoi(e) : 64;
    return k < 1 ? 1 : (k > 4096 ? 4096 : k);
}

typedef struct AmvCudaEncQueue {
    amv_ctx *h;
    AVFrame coded;
    int depth, nq, w, h_, cw, ch;        /* queue depth, frames waiting */
    uint8_t *y, *u, *v;                  /* depth frames, tight planes (pinned) */
    int32_t *qscale;
    int64_t *pts, *out_pts;
    int *quality, *out_quality;
    uint8_t *pk;                         /* the batch's packets (pinned), packed */
    uint64_t pk_cap, *off;
    uint32_t *size;
    int32_t *status;
    int nout, iout;                      /* packets of the last batch, next one to hand out */
} AmvCudaEncQueue;

static int amvcuda_encq_init(AVCodecContext *avctx)
{
    AmvCudaEncQueue *c = avctx->priv_data;
    const int w = avctx->width, h = avctx->height, k = amvcuda_lookahead();
    if (avctx->pix_fmt != PIX_FMT_YUVJ420P && avctx->pix_fmt != PIX_FMT_YUV420P) return -1;
    if (avctx->thread_count > 1 || avctx->trellis || avctx->intra_dc_precision || (avctx->flags & CODEC_FLAG_GRAY)) return -1;
    if (!avctx->time_base.num || !avctx->time_base.den) return -1;
    if (amv_create(NULL, &c->h) != AMV_OK) return -1;
    c->depth = k; c->w = w; c->h_ = h; c->cw = (w + 1) / 2; c->ch = (h + 1) / 2;
    c->pk_cap = (uint64_t)k * (3000ull * ((w + 15) / 16) * ((h + 15) / 16) + 1024);      /* MAX_MB_BYTES per macroblock */
    c->y = amv_host_alloc((size_t)k * w * h);
    c->u = amv_host_alloc((size_t)k * c->cw * c->ch);
    c->v = amv_host_alloc((size_t)k * c->cw * c->ch);
    c->pk = amv_host_alloc(c->pk_cap);
    c->qscale = av_malloc(k * sizeof(*c->qscale)); c->status = av_malloc(k * sizeof(*c->status));
    c->pts = av_malloc(k * sizeof(*c->pts)); c->out_pts = av_malloc(k * sizeof(*c->out_pts));
    c->quality = av_malloc(k * sizeof(int)); c->out_quality = av_malloc(k * sizeof(int));
    c->off = av_malloc(k * sizeof(*c->off)); c->size = av_malloc(k * sizeof(*c->size));
    if (!c->y || !c->u || !c->v || !c->pk || !c->qscale || !c->status || !c->pts || !c->out_pts || !c->off || !c->size) return -1;
    avctx->coded_frame = &c->coded;
    avctx->delay = k - 1;
    return 0;
}

static int amvcuda_encq_close(AVCodecContext *avctx)
{
    AmvCudaEncQueue *c = avctx->priv_data;
    amv_host_free(c->y); amv_host_free(c->u); amv_host_free(c->v); amv_host_free(c->pk);
    av_free(c->qscale); av_free(c->status); av_free(c->pts); av_free(c->out_pts); av_free(c->quality); av_free(c->out_quality);
    av_free(c->off); av_free(c->size);
    amv_destroy(c->h);
    c->h = NULL;
    return 0;
}

static int amvcuda_encq_run(AmvCudaEncQueue *c)
{
    const int w = c->w, h = c->h_;
    int i;
    if (amv_encode_frames(c->h, c->y, c->u, c->v, w, c->cw, (uint64_t)w * h, (uint64_t)c->cw * c->ch, c->nq, w, h, c->qscale,
                          c->pk, c->pk_cap, (uint32_t)(c->pk_cap / c->depth), AMV_LAYOUT_PACKED, c->off, c->size, c->status,
                          AMV_MEM_HOST) != AMV_OK)
        return -1;
    for (i = 0; i < c->nq; i++) { c->out_pts[i] = c->pts[i]; c->out_quality[i] = c->quality[i]; }
    c->nout = c->nq; c->iout = 0; c->nq = 0;
    return 0;
}

/* one frame in (or NULL to drain), at most one packet out; 0 = nothing yet */
static int amvcuda_encq_frame(AVCodecContext *avctx, uint8_t *buf, int buf_size, void *data)
{
    AmvCudaEncQueue *c = avctx->priv_data;
    AVFrame *pic = data;
    const int w = c->w, h = c->h_;
    int r, i;
    if (avctx->flags & CODEC_FLAG_EMU_EDGE) return -1;
    if (pic) {
        uint8_t *dy = c->y + (size_t)c->nq * w * h, *du = c->u + (size_t)c->nq * c->cw * c->ch, *dv = c->v + (size_t)c->nq * c->cw * c->ch;
        for (r = 0; r < h; r++) memcpy(dy + r * w, pic->data[0] + r * pic->linesize[0], w);
        for (r = 0; r < c->ch; r++) {
            memcpy(du + r * c->cw, pic->data[1] + r * pic->linesize[1], c->cw);
            memcpy(dv + r * c->cw, pic->data[2] + r * pic->linesize[2], c->cw);
        }
        c->qscale[c->nq] = amv_qscale_from_quality(pic->quality, avctx->qmin, avctx->qmax);
        c->pts[c->nq] = pic->pts; c->quality[c->nq] = pic->quality;
        c->nq++;
    }
    /* the previous batch is handed out one packet per call, so it is empty exactly when the queue is full again */
    if (c->iout >= c->nout && c->nq > 0 && (c->nq == c->depth || !pic)) {
        if (amvcuda_encq_run(c) < 0) return -1;
    }
    if (c->iout >= c->nout) return 0;
    i = c->iout++;
    if (c->status[i] || (int)c->size[i] > buf_size) return -1;       /* "encoded frame too large" (mpegvideo_enc.c:2077-2080) */
    memcpy(buf, c->pk + c->off[i], c->size[i]);
    c->coded.key_frame = 1;
    c->coded.pict_type = FF_I_TYPE;
    c->coded.quality = c->out_quality[i];
    c->coded.pts = c->out_pts[i];
    return (int)c->size[i];
}

AVCodec amvcuda_amv_encoder_delay = {
    "amv", CODEC_TYPE_VIDEO, CODEC_ID_AMV, sizeof(AmvCudaEncQueue),
    amvcuda_encq_init, amvcuda_encq_frame, amvcuda_encq_close, NULL, CODEC_CAP_DELAY,
    .pix_fmts = amvcuda_pix_fmts,
};

typedef struct AmvCudaDecQueue {
    amv_ctx *h;
    AVFrame picture;
    int depth, nq, w, h_, cw, ch;
    uint8_t *pk;                         /* queued packets (pinned), back to back */
    uint64_t pk_cap, pk_used, *off;
    uint32_t *size;
    int32_t *status;
    uint8_t *y[2], *u[2], *v[2];         /* two batches of decoded planes: one being handed out, one being filled next */
    int cur, nout, iout;
} AmvCudaDecQueue;

static int amvcuda_decq_init(AVCodecContext *avctx)
{
    AmvCudaDecQueue *c = avctx->priv_data;
    avctx->pix_fmt = PIX_FMT_YUVJ420P;
    c->depth = amvcuda_lookahead();
    c->off = av_malloc(c->depth * sizeof(*c->off)); c->size = av_malloc(c->depth * sizeof(*c->size));
    c->status = av_malloc(c->depth * sizeof(*c->status));
    if (!c->off || !c->size || !c->status) return -1;
    return amv_create(NULL, &c->h) == AMV_OK ? 0 : -1;
}

static int amvcuda_decq_close(AVCodecContext *avctx)
{
    AmvCudaDecQueue *c = avctx->priv_data;
    int i;
    for (i = 0; i < 2; i++) { amv_host_free(c->y[i]); amv_host_free(c->u[i]); amv_host_free(c->v[i]); }
    amv_host_free(c->pk);
    av_free(c->off); av_free(c->size); av_free(c->status);
    amv_destroy(c->h);
    c->h = NULL;
    return 0;
}

static int amvcuda_decq_run(AmvCudaDecQueue *c)
{
    const int w = c->w, h = c->h_, b = c->cur ^ 1;
    if (amv_decode_frames(c->h, c->pk, c->pk_used, c->off, c->size, c->nq, w, h, c->y[b], c->u[b], c->v[b], w, c->cw,
                          (uint64_t)w * h, (uint64_t)c->cw * c->ch, c->status, AMV_MEM_HOST) != AMV_OK)
        return -1;
    c->cur = b; c->nout = c->nq; c->iout = 0; c->nq = 0; c->pk_used = 0;
    return 0;
}

/* one packet in (or buf_size 0 to drain), at most one picture out.  The pictures are decoder-owned (planes of the batch,
 * valid until the batch after next is decoded, i.e. for at least `depth` further calls). */
static int amvcuda_decq_frame(AVCodecContext *avctx, void *data, int *data_size, uint8_t *buf, int buf_size)
{
    AmvCudaDecQueue *c = avctx->priv_data;
    AVFrame *out = data;
    const int w = avctx->width, h = avctx->height;
    int i;
    if (!w || !h) return -1;
    if (!c->pk) {        /* first call: the dimensions are known now (avidec.c:429-434 sets them from the container) */
        c->w = w; c->h_ = h; c->cw = (w + 1) / 2; c->ch = (h + 1) / 2;
        c->pk_cap = (uint64_t)c->depth * (3000ull * ((w + 15) / 16) * ((h + 15) / 16) + 1024);
        c->pk = amv_host_alloc(c->pk_cap);
        for (i = 0; i < 2; i++) {
            c->y[i] = amv_host_alloc((size_t)c->depth * w * h);
            c->u[i] = amv_host_alloc((size_t)c->depth * c->cw * c->ch);
            c->v[i] = amv_host_alloc((size_t)c->depth * c->cw * c->ch);
            if (!c->y[i] || !c->u[i] || !c->v[i]) return -1;
        }
        if (!c->pk) return -1;
    } else if (w != c->w || h != c->h_) return -1;
    avctx->pix_fmt = PIX_FMT_YUVJ420P;
    if (buf_size > 0) {
        if (c->pk_used + (uint64_t)buf_size > c->pk_cap) return -1;
        memcpy(c->pk + c->pk_used, buf, buf_size);
        c->off[c->nq] = c->pk_used; c->size[c->nq] = (uint32_t)buf_size;
        c->pk_used += (uint64_t)buf_size; c->nq++;
    }
    if (c->iout >= c->nout && c->nq > 0 && (c->nq == c->depth || buf_size == 0)) {
        if (amvcuda_decq_run(c) < 0) return -1;
    }
    *data_size = 0;
    if (c->iout >= c->nout) return buf_size;
    i = c->iout++;
    memset(&c->picture, 0, sizeof(c->picture));
    c->picture.data[0] = c->y[c->cur] + (size_t)i * w * h;
    c->picture.data[1] = c->u[c->cur] + (size_t)i * c->cw * c->ch;
    c->picture.data[2] = c->v[c->cur] + (size_t)i * c->cw * c->ch;
    c->picture.linesize[0] = w; c->picture.linesize[1] = c->cw; c->picture.linesize[2] = c->cw;
    c->picture.pict_type = FF_I_TYPE;
    c->picture.key_frame = 1;
    c->picture.quality = 9 * FF_QP2LAMBDA;
    *out = c->picture;
    *data_size = sizeof(AVFrame);
    return buf_size;
}

AVCodec amvcuda_amv_decoder_delay = {
    "amv", CODEC_TYPE_VIDEO, CODEC_ID_AMV, sizeof(AmvCudaDecQueue),
    amvcuda_decq_init, NULL, amvcuda_decq_close, amvcuda_decq_frame, CODEC_CAP_DELAY,
};

/* ------------------------------------------------------------------------------------- audio */
typedef struct AmvCudaAudio {
    amv_ctx *h;
    AVFrame coded;
    int16_t step_index;          /* carried across calls like ADPCMChannelStatus.step_index (adpcm.c:466) */
    int extra_amv_samples;       /* adpcm.c:148-149 */
    int samples_written;
} AmvCudaAudio;

static int amvcuda_adpcm_init(AVCodecContext *avctx)
{
    AmvCudaAudio *c = avctx->priv_data;
    if (avctx->codec->encode) {                                           /* adpcm.c:190-199 */
        if (avctx->channels != 1 || avctx->sample_rate != 22050 || avctx->trellis < 0 || avctx->trellis > 5) return -1;
        avctx->coded_frame = &c->coded;
        c->coded.key_frame = 1;
    } else if (avctx->channels > 2) return -1;
    if (amv_create(NULL, &c->h) != AMV_OK) return -1;
    /* -trellis N selects adpcm_compress_trellis in the reference (adpcm.c:481-488); same switch here */
    if (avctx->codec->encode && avctx->trellis > 0 && amv_set_option(c->h, "adpcm_trellis", avctx->trellis) != AMV_OK) return -1;
    return 0;
}

static int amvcuda_adpcm_close(AVCodecContext *avctx)
{
    AmvCudaAudio *c = avctx->priv_data;
    amv_destroy(c->h);
    c->h = NULL;
    return 0;
}

/* adpcm_decode_frame, AMV case (adpcm.c:894-935,1268-1292) */
static int amvcuda_adpcm_dec_frame(AVCodecContext *avctx, void *data, int *data_size, uint8_t *buf, int buf_size)
{
    AmvCudaAudio *c = avctx->priv_data;
    uint64_t off = 0, pcm_off = 0;
    uint32_t size = (uint32_t)buf_size;
    int32_t status = 0;
    if (!buf_size) return 0;
    if (*data_size / 4 < buf_size + 8) return -1;                          /* adpcm.c:924 */
    *data_size = 0;
    if (buf_size < 8) return buf_size;
    if (amv_adpcm_dec_chunks(c->h, buf, (uint64_t)buf_size, &off, &size, 1, (int16_t *)data,
                             2ull * (buf_size - 8), &pcm_off, &status, AMV_MEM_HOST) != AMV_OK || status)
        return -1;
    *data_size = 4 * (buf_size - 8);
    return buf_size;
}

/* adpcm_encode_frame, AMV case (adpcm.c:461-496): the chunk takes 2n samples where n follows the
 * reference's frame_size / odd-sample / second-boundary bookkeeping; the step index is chained. */
static int amvcuda_adpcm_enc_frame(AVCodecContext *avctx, unsigned char *frame, int buf_size, void *data)
{
    AmvCudaAudio *c = avctx->priv_data;
    uint64_t pcm_off = 0, out_off = 0;
    uint32_t nsamples;
    int16_t step_out = 0;
    int32_t status = 0;
    int n, i;
    avctx->coded_frame->pts = c->samples_written;
    n = avctx->frame_size >> 1;
    c->extra_amv_samples += avctx->frame_size & 1;
    n += c->extra_amv_samples >> 1;
    c->extra_amv_samples &= 1;
    i = (c->samples_written + 2 * n) % avctx->sample_rate;
    if (i && i + avctx->frame_size > avctx->sample_rate) n += (avctx->sample_rate - i) >> 1;
    nsamples = 2u * (uint32_t)n;
    if (buf_size < 8 + n) return -1;
    if (amv_adpcm_enc_chunks(c->h, (const int16_t *)data, nsamples, &pcm_off, &nsamples, &c->step_index, &step_out, 1,
                             frame, (uint64_t)(8 + n), &out_off, &status, AMV_MEM_HOST) != AMV_OK || status)
        return -1;
    c->step_index = step_out;
    c->samples_written += (int)nsamples;
    return 8 + n;
}

AVCodec amvcuda_adpcm_ima_amv_decoder = {
    "adpcm_ima_amv", CODEC_TYPE_AUDIO, CODEC_ID_ADPCM_IMA_AMV, sizeof(AmvCudaAudio),
    amvcuda_adpcm_init, NULL, amvcuda_adpcm_close, amvcuda_adpcm_dec_frame,
};
AVCodec amvcuda_adpcm_ima_amv_encoder = {
    "adpcm_ima_amv", CODEC_TYPE_AUDIO, CODEC_ID_ADPCM_IMA_AMV, sizeof(AmvCudaAudio),
    amvcuda_adpcm_init, amvcuda_adpcm_enc_frame, amvcuda_adpcm_close, NULL,
};

/* Call before avcodec_register_all(): first match wins in avcodec_find_{en,de}coder. */
void amvcuda_register_codecs(void)
{
    const int delayed = getenv("AMVCUDA_LOOKAHEAD") != NULL;        /* opt-in: the batch (look-ahead) video codecs */
    register_avcodec(delayed ? &amvcuda_amv_encoder_delay : &amvcuda_amv_encoder);
    register_avcodec(delayed ? &amvcuda_amv_decoder_delay : &amvcuda_amv_decoder);
    register_avcodec(&amvcuda_sp5x_decoder);
    register_avcodec(&amvcuda_mjpeg_decoder);
    register_avcodec(&amvcuda_adpcm_ima_amv_encoder);
    register_avcodec(&amvcuda_adpcm_ima_amv_decoder);
}
