/*
 * amvcuda_resample.c -- reference-side binding for the two resampling stages ffmpeg.c runs in front of the AMV
 * encoders.  Compiled inside the reference tree's include path like amvcuda_codecs.c; no
 * arithmetic here, only the translation of the reference's call conventions onto include/amvcuda.h.
 *
 *   amvcuda_sws_getContext / amvcuda_sws_scale / amvcuda_sws_freeContext
 *       same signatures as the fork's libavcodec emulation of libswscale (imgresample.c:515-690), which is what
 *       ffmpeg.c calls for `-s WxH` (ffmpeg.c:757,1684-1691,2074).  Planar 4:2:0 in and out, PIX_FMT_YUV420P or
 *       PIX_FMT_YUVJ420P on either side (the AMV encoder takes YUVJ420P), with the reference's img_convert steps
 *       around the scaler; other formats return NULL, and the caller keeps the reference's chain for those.
 *   amvcuda_audio_resample_init / amvcuda_audio_resample / amvcuda_audio_resample_close
 *       same signatures as resample.c:93-129,131-235,237-243 for 1 output channel (what adpcm_ima_amv takes),
 *       1 or 2 input channels.  Like the reference, the context carries the unconsumed tail of the input from
 *       call to call and limits one call's output to lenout = 4 * nb_samples * ratio + 16 samples.
 *
 * To switch ffmpeg.c over a maintainer adds, after the includes of ffmpeg.c:
 *   #define sws_getContext amvcuda_sws_getContext     (and sws_scale, sws_freeContext, audio_resample_init, ...)
 */
#include <stdlib.h>
#include <string.h>
#include "avcodec.h"
#include "swscale.h"
#include "amvcuda.h"

/* ------------------------------------------------------------------------------ picture scaler */
typedef struct AmvCudaSws {
    amv_ctx *h;
    int iw, ih, ow, oh;
    int src_fmt, dst_fmt;
} AmvCudaSws;

struct SwsContext *amvcuda_sws_getContext(int srcW, int srcH, int srcFormat, int dstW, int dstH, int dstFormat,
                                          int flags, SwsFilter *srcFilter, SwsFilter *dstFilter, double *param)
{
    AmvCudaSws *s;
    if ((srcFormat != PIX_FMT_YUV420P && srcFormat != PIX_FMT_YUVJ420P) || (dstFormat != PIX_FMT_YUV420P && dstFormat != PIX_FMT_YUVJ420P))
        return NULL;
    if (srcW <= 0 || srcH <= 0 || dstW <= 0 || dstH <= 0) return NULL;            /* img_resample_full_init :447-448 */
    s = av_mallocz(sizeof(*s));
    if (!s) return NULL;
    if (amv_create(NULL, &s->h) != AMV_OK) { av_free(s); return NULL; }              /* no device: no CPU path */
    s->iw = srcW; s->ih = srcH; s->ow = dstW; s->oh = dstH;
    s->src_fmt = srcFormat; s->dst_fmt = dstFormat;
    return (struct SwsContext *)s;
}

void amvcuda_sws_freeContext(struct SwsContext *ctx)
{
    AmvCudaSws *s = (AmvCudaSws *)ctx;
    if (!s) return;
    amv_destroy(s->h);
    av_free(s);
}

int amvcuda_sws_scale(struct SwsContext *ctx, uint8_t *src[], int srcStride[], int srcSliceY, int srcSliceH,
                      uint8_t *dst[], int dstStride[])
{
    AmvCudaSws *s = (AmvCudaSws *)ctx;
    int p, r;
    if (s->iw == s->ow && s->ih == s->oh && s->src_fmt != s->dst_fmt)      /* imgresample.c:671-682: img_convert only */
        return amv_convert_range(s->h, src[0], src[1], src[2], srcStride[0], srcStride[1], (uint64_t)srcStride[0] * s->ih,
                                 (uint64_t)srcStride[1] * ((s->ih + 1) / 2), 1, s->iw, s->ih, s->dst_fmt == PIX_FMT_YUVJ420P ? 0 : 1,
                                 dst[0], dst[1], dst[2], dstStride[0], dstStride[1], (uint64_t)dstStride[0] * s->oh,
                                 (uint64_t)dstStride[1] * ((s->oh + 1) / 2), AMV_MEM_HOST) == AMV_OK ? 0 : -1;
    if (s->iw == s->ow && s->ih == s->oh) {                  /* imgresample.c:683-686: a plain copy, no arithmetic */
        for (p = 0; p < 3; p++) {
            const int w = p ? (s->ow + 1) >> 1 : s->ow, h = p ? (s->oh + 1) >> 1 : s->oh;   /* av_picture_copy rounds up */
            for (r = 0; r < h; r++) memcpy(dst[p] + (size_t)r * dstStride[p], src[p] + (size_t)r * srcStride[p], w);
        }
        return 0;
    }
    if (srcStride[1] != srcStride[2] || dstStride[1] != dstStride[2]) return -1;
    /* the scaler works on YUV420P; a YUVJ420P side goes through img_convert first / afterwards (imgresample.c:617-682) */
    return amv_scale_frames_ex(s->h, src[0], src[1], src[2], srcStride[0], srcStride[1],
                            (uint64_t)srcStride[0] * s->ih, (uint64_t)srcStride[1] * ((s->ih + 1) / 2), 1, s->iw, s->ih,
                            dst[0], dst[1], dst[2], dstStride[0], dstStride[1],
                            (uint64_t)dstStride[0] * s->oh, (uint64_t)dstStride[1] * ((s->oh + 1) / 2), s->ow, s->oh,
                            (s->src_fmt == PIX_FMT_YUVJ420P ? AMV_SCALE_IN_JPEG_RANGE : 0) |
                            (s->dst_fmt == PIX_FMT_YUVJ420P ? AMV_SCALE_OUT_JPEG_RANGE : 0),
                            AMV_MEM_HOST) == AMV_OK ? 0 : -1;
}

/* ------------------------------------------------------------------------------ audio resampler */
typedef struct AmvCudaResample {
    amv_ctx *h;
    int in_ch, in_rate, out_rate;
    float ratio;                 /* resample.c:111 */
    short *buf;                  /* the stream's samples [base, base + n_buf), interleaved */
    int64_t base, n_buf, cap;
    int64_t k_next;              /* next output of the stream */
} AmvCudaResample;

ReSampleContext *amvcuda_audio_resample_init(int output_channels, int input_channels, int output_rate, int input_rate)
{
    AmvCudaResample *s;
    if (input_channels > 2 || input_channels < 1 || output_channels != 1) return NULL;   /* resample.c:98-102; AMV audio is mono */
    s = av_mallocz(sizeof(*s));
    if (!s) return NULL;
    if (amv_create(NULL, &s->h) != AMV_OK) { av_free(s); return NULL; }
    s->in_ch = input_channels; s->in_rate = input_rate; s->out_rate = output_rate;
    s->ratio = (float)output_rate / (float)input_rate;
    return (ReSampleContext *)s;
}

void amvcuda_audio_resample_close(ReSampleContext *ctx)
{
    AmvCudaResample *s = (AmvCudaResample *)ctx;
    if (!s) return;
    amv_destroy(s->h);
    av_free(s->buf); av_free(s);
}

int amvcuda_audio_resample(ReSampleContext *ctx, short *output, short *input, int nb_samples)
{
    AmvCudaResample *s = (AmvCudaResample *)ctx;
    const int lenout = (int)(4 * nb_samples * s->ratio) + 16;                      /* resample.c:154 */
    uint64_t avail, got = 0;
    int64_t keep;
    if (s->n_buf + nb_samples > s->cap) {
        s->cap = 2 * (s->n_buf + nb_samples) + 64;
        s->buf = av_realloc(s->buf, sizeof(short) * s->cap * s->in_ch);
    }
    memcpy(s->buf + s->n_buf * s->in_ch, input, sizeof(short) * (size_t)nb_samples * s->in_ch);
    s->n_buf += nb_samples;
    avail = amv_audio_resample_count((uint64_t)(s->base + s->n_buf), s->in_rate, s->out_rate);
    avail = avail > (uint64_t)s->k_next ? avail - (uint64_t)s->k_next : 0;
    /* av_resample stops at dst_size = lenout (resample.c:154,202); the rest comes with the next call */
    if (avail && amv_audio_resample_from(s->h, s->buf, (uint64_t)s->base, (uint64_t)s->n_buf, s->in_ch, s->in_rate, s->out_rate,
                                         (uint64_t)s->k_next, output, (uint64_t)lenout, &got, AMV_MEM_HOST) != AMV_OK)
        return -1;
    s->k_next += (int64_t)got;
    keep = amv_audio_resample_first_tap((uint64_t)s->k_next, s->in_rate, s->out_rate);   /* resample2.c:303: consumed */
    if (keep > s->base) {
        int64_t drop = keep - s->base;
        if (drop > s->n_buf) drop = s->n_buf;
        memmove(s->buf, s->buf + drop * s->in_ch, sizeof(short) * (size_t)(s->n_buf - drop) * s->in_ch);
        s->base += drop; s->n_buf -= drop;
    }
    return (int)got;
}
