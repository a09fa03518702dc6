/*
 * dropin_check.c -- exercises the AVCodec drop-in exactly the way ffmpeg.c drives a codec
 * (avcodec_find_* -> avcodec_open -> avcodec_{en,de}code_* per frame/chunk, ffmpeg.c:1062,1083,522,814)
 * once with the libamvcuda shims registered first and once with the reference's own codecs, and
 * compares packets, planes, chunks and PCM byte for byte.  Built against the reference's headers
 * and objects by glue/build_dropin.sh (only where the reference tree is mounted); the binary
 * travels to the GPU box.  Exit code 0 = identical.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include "avcodec.h"
#include "swscale.h"

struct SwsContext *amvcuda_sws_getContext(int srcW, int srcH, int srcFormat, int dstW, int dstH, int dstFormat, int flags,
                                          SwsFilter *srcFilter, SwsFilter *dstFilter, double *param);
void amvcuda_sws_freeContext(struct SwsContext *ctx);
int amvcuda_sws_scale(struct SwsContext *ctx, uint8_t *src[], int srcStride[], int srcSliceY, int srcSliceH, uint8_t *dst[], int dstStride[]);
ReSampleContext *amvcuda_audio_resample_init(int output_channels, int input_channels, int output_rate, int input_rate);
int amvcuda_audio_resample(ReSampleContext *s, short *output, short *input, int nb_samples);
void amvcuda_audio_resample_close(ReSampleContext *s);

extern AVCodec amv_decoder, amv_encoder, adpcm_ima_amv_decoder, adpcm_ima_amv_encoder, sp5x_decoder, mjpeg_decoder, mjpeg_encoder;
extern AVCodec amvcuda_amv_decoder, amvcuda_amv_encoder, amvcuda_adpcm_ima_amv_decoder, amvcuda_adpcm_ima_amv_encoder,
               amvcuda_sp5x_decoder, amvcuda_mjpeg_decoder, amvcuda_amv_encoder_delay, amvcuda_amv_decoder_delay;
void amvcuda_register_codecs(void);

static int g_pix_fmt = PIX_FMT_YUVJ420P;      /* input / output layout of encode_all / decode_all */
static unsigned rng_state = 12345;
static unsigned rnd(void) { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

static void make_frame(uint8_t *y, uint8_t *u, uint8_t *v, int w, int h, int t)
{
    int cw = (w + 1) / 2, ch = (h + 1) / 2, x, r;
    for (r = 0; r < h; r++) for (x = 0; x < w; x++) {
        double s = 128 + 60 * sin((x + 3 * t) / 17.0) + 50 * cos((r - 2 * t) / 11.0) + (int)(rnd() % 25) - 12;
        y[r * w + x] = s < 0 ? 0 : (s > 255 ? 255 : (uint8_t)s);
    }
    for (r = 0; r < ch; r++) for (x = 0; x < cw; x++) {
        u[r * cw + x] = (uint8_t)(128 + 40 * sin((x + t) / 23.0));
        v[r * cw + x] = (uint8_t)(128 + 40 * cos((r + t) / 19.0));
    }
}

typedef struct { uint8_t *pk; int size; } Packet;

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

/* The per-frame loop of ffmpeg.c for codecs that may delay their output (CODEC_CAP_DELAY): a call may return nothing
 * (ffmpeg.c:1088 `if (!got_picture) goto discard_packet`; the encoder side writes a packet only `if (ret > 0)`), and at
 * the end of the stream the codec is called with no input until it has nothing left.  `reps` passes over the n frames are
 * timed (codec calls only, open / close outside); the outputs of the first pass are kept. */
static int encode_timed(AVCodec *codec, int w, int h, int n, int quality, uint8_t **ys, uint8_t **us, uint8_t **vs, Packet *out,
                        int reps, double *fps)
{
    AVCodecContext *c = avcodec_alloc_context();
    AVFrame *pic = avcodec_alloc_frame();
    int i, r, got = 0, total = 0, bufsz = w * h * 6 + 262144, cw = (w + 1) / 2;
    uint8_t *buf = av_malloc(bufsz);
    double t0;
    c->width = w; c->height = h; c->time_base.num = 1; c->time_base.den = 16; c->pix_fmt = g_pix_fmt;
    if (avcodec_open(c, codec) < 0) return -1;
    t0 = now_s();
    for (r = 0; r < reps; r++)
        for (i = 0; i <= n; i++) {
            int sz;
            if (i < n) {
                pic->data[0] = ys[i]; pic->data[1] = us[i]; pic->data[2] = vs[i];
                pic->linesize[0] = w; pic->linesize[1] = cw; pic->linesize[2] = cw;
                pic->quality = quality; pic->pts = (int64_t)r * n + i;
                sz = avcodec_encode_video(c, buf, bufsz, pic);
            } else if (r == reps - 1 && (codec->capabilities & CODEC_CAP_DELAY)) {
                sz = avcodec_encode_video(c, buf, bufsz, NULL);            /* drain */
                if (sz > 0) i--;                                           /* ... until it answers 0 */
            } else break;
            if (sz < 0) return -2;
            if (sz == 0) continue;
            if (c->coded_frame->pts != total) return -4;                   /* packets come out in frame order with their own pts */
            if (got < n) {
                out[got].pk = malloc(sz + FF_INPUT_BUFFER_PADDING_SIZE); memset(out[got].pk, 0, sz + FF_INPUT_BUFFER_PADDING_SIZE);
                memcpy(out[got].pk, buf, sz); out[got].size = sz; got++;
            }
            total++;
        }
    *fps = total / (now_s() - t0);
    avcodec_close(c); av_free(c); av_free(pic); av_free(buf);
    return total == reps * n ? 0 : -5;
}

static int decode_timed(AVCodec *codec, int w, int h, int n, Packet *pk, uint8_t *planes, int reps, double *fps)
{
    AVCodecContext *c = avcodec_alloc_context();
    AVFrame *pic = avcodec_alloc_frame();
    int i, r, rr, outn = 0, total = 0, cw = (w + 1) / 2, ch = (h + 1) / 2;
    double t0;
    c->width = w; c->height = h; c->coded_width = w; c->coded_height = h;
    if (avcodec_open(c, codec) < 0) return -1;
    t0 = now_s();
    for (r = 0; r < reps; r++)
        for (i = 0; i <= n; i++) {
            int got = 0, ret;
            if (i < n) ret = avcodec_decode_video(c, pic, &got, pk[i].pk, pk[i].size);
            else if (r == reps - 1 && (codec->capabilities & CODEC_CAP_DELAY)) {
                ret = avcodec_decode_video(c, pic, &got, NULL, 0);        /* drain (utils.c:938) */
                if (got) i--;
            } else break;
            if (ret < 0) return -2;
            if (!got) continue;
            if (outn < n) {
                uint8_t *d = planes + (size_t)outn * (w * h + 2 * cw * ch);
                for (rr = 0; rr < h; rr++) memcpy(d + rr * w, pic->data[0] + rr * pic->linesize[0], w);
                for (rr = 0; rr < ch; rr++) {
                    memcpy(d + w * h + rr * cw, pic->data[1] + rr * pic->linesize[1], cw);
                    memcpy(d + w * h + cw * ch + rr * cw, pic->data[2] + rr * pic->linesize[2], cw);
                }
                outn++;
            }
            total++;
        }
    *fps = total / (now_s() - t0);
    avcodec_close(c); av_free(c); av_free(pic);
    return total == reps * n ? 0 : -5;
}

static int encode_all(AVCodec *codec, int w, int h, int n, int quality, uint8_t **ys, uint8_t **us, uint8_t **vs, Packet *out)
{
    AVCodecContext *c = avcodec_alloc_context();
    AVFrame *pic = avcodec_alloc_frame();
    int i, bufsz = w * h * 6 + 262144, cw = (w + 1) / 2;
    uint8_t *buf = av_malloc(bufsz);
    c->width = w; c->height = h; c->time_base.num = 1; c->time_base.den = 16; c->pix_fmt = g_pix_fmt;
    if (avcodec_open(c, codec) < 0) return -1;
    for (i = 0; i < n; i++) {
        pic->data[0] = ys[i]; pic->data[1] = us[i]; pic->data[2] = vs[i];      /* 4:2:2: the caller passes full-height chroma */
        pic->linesize[0] = w; pic->linesize[1] = cw; pic->linesize[2] = cw;
        pic->quality = quality; pic->pts = i;
        int sz = avcodec_encode_video(c, buf, bufsz, pic);
        if (sz < 0) return -2;
        out[i].pk = malloc(sz + FF_INPUT_BUFFER_PADDING_SIZE); memset(out[i].pk, 0, sz + FF_INPUT_BUFFER_PADDING_SIZE);
        memcpy(out[i].pk, buf, sz); out[i].size = sz;
        if (!c->coded_frame->key_frame) return -3;
    }
    avcodec_close(c); av_free(c); av_free(pic); av_free(buf);
    return 0;
}

static int decode_all(AVCodec *codec, int w, int h, int n, Packet *pk, uint8_t *planes /* n * w*h*3/2 tight */)
{
    AVCodecContext *c = avcodec_alloc_context();
    AVFrame *pic = avcodec_alloc_frame();
    int i, r, cw = (w + 1) / 2, ch = g_pix_fmt == PIX_FMT_YUVJ422P ? h : (h + 1) / 2;
    c->width = w; c->height = h; c->coded_width = w; c->coded_height = h;
    if (avcodec_open(c, codec) < 0) return -1;
    for (i = 0; i < n; i++) {
        int got = 0;
        int ret = avcodec_decode_video(c, pic, &got, pk[i].pk, pk[i].size);
        if (ret < 0 || !got) return -2;
        if (!pic->key_frame || pic->pict_type != FF_I_TYPE || c->pix_fmt != g_pix_fmt) return -3;
        uint8_t *d = planes + (size_t)i * (w * h + 2 * cw * ch);
        for (r = 0; r < h; r++) memcpy(d + r * w, pic->data[0] + r * pic->linesize[0], w);
        for (r = 0; r < ch; r++) {
            memcpy(d + w * h + r * cw, pic->data[1] + r * pic->linesize[1], cw);
            memcpy(d + w * h + cw * ch + r * cw, pic->data[2] + r * pic->linesize[2], cw);
        }
    }
    avcodec_close(c); av_free(c); av_free(pic);
    return 0;
}

static int g_trellis;
static int audio_roundtrip(AVCodec *enc, AVCodec *dec, const int16_t *pcm, int total, int frame_size,
                           uint8_t *chunks, int *chunk_bytes, int16_t *out_pcm, int *out_samples)
{
    AVCodecContext *e = avcodec_alloc_context(), *d = avcodec_alloc_context();
    int pos = 0, cb = 0, os = 0;
    uint8_t buf[FF_MIN_BUFFER_SIZE + 65536];
    int16_t *tmp = av_malloc(AVCODEC_MAX_AUDIO_FRAME_SIZE * 2);
    e->channels = 1; e->sample_rate = 22050; e->frame_size = frame_size; e->trellis = g_trellis;
    d->channels = 1; d->sample_rate = 22050;
    if (avcodec_open(e, enc) < 0 || avcodec_open(d, dec) < 0) return -1;
    e->frame_size = frame_size;
    while (pos + 2 * frame_size + 2 <= total) {
        int sz = avcodec_encode_audio(e, buf, sizeof(buf), pcm + pos);
        if (sz < 8) return -2;
        int two_n = buf[4] | (buf[5] << 8) | (buf[6] << 16) | (buf[7] << 24);
        memcpy(chunks + cb, buf, sz); cb += sz; pos += two_n;
        int bytes = AVCODEC_MAX_AUDIO_FRAME_SIZE * 2;
        if (avcodec_decode_audio2(d, tmp, &bytes, buf, sz) < 0) return -3;
        memcpy(out_pcm + os, tmp, bytes); os += bytes / 2;
    }
    *chunk_bytes = cb; *out_samples = os;
    avcodec_close(e); avcodec_close(d); av_free(e); av_free(d); av_free(tmp);
    return 0;
}

int main(int argc, char **argv)
{
    const int w = argc > 1 ? atoi(argv[1]) : 160, h = argc > 2 ? atoi(argv[2]) : 120, n = argc > 3 ? atoi(argv[3]) : 24;
    const int cw = (w + 1) / 2, ch = (h + 1) / 2, fb = w * h + 2 * cw * ch;
    int i, fail = 0;
    avcodec_init();
    av_log_set_level(AV_LOG_QUIET);
    unsetenv("AMVCUDA_LOOKAHEAD");                   /* the plain shims are the registered ones in this check */
    amvcuda_register_codecs();                       /* first match wins ...                              */
    register_avcodec(&amv_encoder); register_avcodec(&amv_decoder);   /* ... then what avcodec_register_all adds */
    register_avcodec(&adpcm_ima_amv_encoder); register_avcodec(&adpcm_ima_amv_decoder);
    register_avcodec(&sp5x_decoder); register_avcodec(&mjpeg_decoder);
    if (avcodec_find_decoder(CODEC_ID_AMV) != &amvcuda_amv_decoder || avcodec_find_decoder(CODEC_ID_SP5X) != &amvcuda_sp5x_decoder || avcodec_find_decoder(CODEC_ID_MJPEG) != &amvcuda_mjpeg_decoder || avcodec_find_encoder(CODEC_ID_AMV) != &amvcuda_amv_encoder ||
        avcodec_find_decoder(CODEC_ID_ADPCM_IMA_AMV) != &amvcuda_adpcm_ima_amv_decoder ||
        avcodec_find_encoder(CODEC_ID_ADPCM_IMA_AMV) != &amvcuda_adpcm_ima_amv_encoder) {
        printf("FAIL: lookup does not return the drop-in codecs\n");
        return 2;
    }
    uint8_t **ys = malloc(n * sizeof(*ys)), **us = malloc(n * sizeof(*us)), **vs = malloc(n * sizeof(*vs));
    for (i = 0; i < n; i++) {
        ys[i] = malloc(w * h); us[i] = malloc(cw * ch); vs[i] = malloc(cw * ch);
        make_frame(ys[i], us[i], vs[i], w, h, i);
    }
    int q;
    for (q = 0; q <= 2; q++) {
        const int quality = q == 0 ? 0 : (q == 1 ? 5 * FF_QP2LAMBDA : 31 * FF_QP2LAMBDA);
        Packet *pa = calloc(n, sizeof(Packet)), *pb = calloc(n, sizeof(Packet));
        int ra = encode_all(avcodec_find_encoder(CODEC_ID_AMV), w, h, n, quality, ys, us, vs, pa);
        int rb = encode_all(&amv_encoder, w, h, n, quality, ys, us, vs, pb);
        if (ra || rb) { printf("FAIL: encode returned %d / %d\n", ra, rb); return 3; }
        for (i = 0; i < n; i++)
            if (pa[i].size != pb[i].size || memcmp(pa[i].pk, pb[i].pk, pa[i].size)) { printf("FAIL: packet %d differs (quality %d)\n", i, quality); fail = 1; }
        uint8_t *da = malloc((size_t)n * fb), *db = malloc((size_t)n * fb);
        ra = decode_all(avcodec_find_decoder(CODEC_ID_AMV), w, h, n, pb, da);
        rb = decode_all(&amv_decoder, w, h, n, pb, db);
        if (ra || rb) { printf("FAIL: decode returned %d / %d\n", ra, rb); return 4; }
        if (memcmp(da, db, (size_t)n * fb)) { printf("FAIL: decoded planes differ (quality %d)\n", quality); fail = 1; }
        printf("video %dx%d x%d quality %d: packets and planes %s\n", w, h, n, quality, fail ? "DIFFER" : "identical");
        {   /* SP5X: the same scans behind a 14-byte header with literal FF bytes (sp5xdec.c:78-84; the reference has no
             * SP5X encoder).  Only packets whose FF count fits the reference's recode buffer (:51) are comparable. */
            Packet *ps = calloc(n, sizeof(Packet));
            int m = 0, j, k;
            for (i = 0; i < n; i++) {
                int ff = 0;
                uint8_t *d = calloc(1, pb[i].size + 14 + FF_INPUT_BUFFER_PADDING_SIZE);
                k = 14;
                for (j = 2; j < pb[i].size - 2; j++) { d[k++] = pb[i].pk[j]; if (pb[i].pk[j] == 0xff) { ff++; j++; } }
                if (ff <= 400) { ps[m].pk = d; ps[m].size = k; m++; } else free(d);
            }
            if (m) {
                ra = decode_all(avcodec_find_decoder(CODEC_ID_SP5X), w, h, m, ps, da);
                rb = decode_all(&sp5x_decoder, w, h, m, ps, db);
                if (ra || rb) { printf("FAIL: sp5x decode returned %d / %d\n", ra, rb); return 6; }
                if (memcmp(da, db, (size_t)m * fb)) { printf("FAIL: sp5x planes differ (quality %d)\n", quality); fail = 1; }
            }
            printf("sp5x  %dx%d x%d quality %d: planes %s\n", w, h, m, quality, fail ? "DIFFER" : "identical");
        }
        {   /* plain MJPEG: full JPEG frames from the reference's mjpeg_encoder (its rate control rewrites the DQT
             * segment from frame to frame), decoded by the drop-in and by the reference's mjpeg_decoder */
            Packet *pm = calloc(n, sizeof(Packet));
            int mfail = 0;
            ra = encode_all(&mjpeg_encoder, w, h, n, quality, ys, us, vs, pm);
            if (ra) { printf("FAIL: reference mjpeg encode returned %d\n", ra); return 7; }
            ra = decode_all(avcodec_find_decoder(CODEC_ID_MJPEG), w, h, n, pm, da);
            rb = decode_all(&mjpeg_decoder, w, h, n, pm, db);
            if (ra || rb) { printf("FAIL: mjpeg decode returned %d / %d\n", ra, rb); return 8; }
            if (memcmp(da, db, (size_t)n * fb)) { printf("FAIL: mjpeg planes differ (quality %d)\n", quality); fail = 1; mfail = 1; }
            printf("mjpeg %dx%d x%d quality %d: planes %s\n", w, h, n, quality, mfail ? "DIFFER" : "identical");
            {   /* the same through YUVJ422P (the encoder writes 2x2 / 1x2 / 1x2 sampling; chroma planes cw x h) */
                uint8_t **u2 = malloc(n * sizeof(*u2)), **v2 = malloc(n * sizeof(*v2));
                uint8_t *ea = malloc((size_t)n * (w * h + 2 * cw * h)), *eb = malloc((size_t)n * (w * h + 2 * cw * h));
                int r2, f2 = 0;
                for (i = 0; i < n; i++) {
                    u2[i] = malloc(cw * h); v2[i] = malloc(cw * h);
                    for (r2 = 0; r2 < h; r2++) { memcpy(u2[i] + r2 * cw, us[i] + (r2 / 2) * cw, cw); memcpy(v2[i] + r2 * cw, vs[i] + (r2 / 2) * cw, cw); }
                }
                g_pix_fmt = PIX_FMT_YUVJ422P;
                ra = encode_all(&mjpeg_encoder, w, h, n, quality, ys, u2, v2, pm);
                if (!ra) ra = decode_all(avcodec_find_decoder(CODEC_ID_MJPEG), w, h, n, pm, ea);
                rb = decode_all(&mjpeg_decoder, w, h, n, pm, eb);
                g_pix_fmt = PIX_FMT_YUVJ420P;
                if (ra || rb) { printf("FAIL: mjpeg 4:2:2 returned %d / %d\n", ra, rb); return 9; }
                if (memcmp(ea, eb, (size_t)n * (w * h + 2 * cw * h))) { printf("FAIL: mjpeg 4:2:2 planes differ (quality %d)\n", quality); fail = 1; f2 = 1; }
                printf("mjpeg 4:2:2 %dx%d x%d quality %d: planes %s\n", w, h, n, quality, f2 ? "DIFFER" : "identical");
                free(ea); free(eb);
            }
        }
        free(da); free(db);
    }
    {   /* what the drop-in costs per call, next to the callbacks it replaces: the same frames through the reference's CPU
         * codecs, through the shims one frame per call, and through the look-ahead (CODEC_CAP_DELAY) shims that hand
         * libamvcuda whole batches; outputs of all three must be the same bytes */
        const int scale = argc > 4 ? atoi(argv[4]) : 100;      /* percent of the default timing effort */
        if (argc > 5) setenv("AMVCUDA_LOOKAHEAD", argv[5], 1); /* queue depth of the look-ahead codecs (default 64) */
        const int reps_ref = 1 + 2 * scale / n, reps = 1 + 20 * scale / n, reps_q = 1 + 200 * scale / n;
        Packet *pr = calloc(n, sizeof(Packet)), *ps = calloc(n, sizeof(Packet)), *pq = calloc(n, sizeof(Packet));
        uint8_t *dr = malloc((size_t)n * fb), *ds = malloc((size_t)n * fb), *dq = malloc((size_t)n * fb);
        double f_er = 0, f_es = 0, f_eq = 0, f_dr = 0, f_ds = 0, f_dq = 0;
        int r1 = encode_timed(&amv_encoder, w, h, n, 0, ys, us, vs, pr, reps_ref, &f_er);
        int r2 = encode_timed(&amvcuda_amv_encoder, w, h, n, 0, ys, us, vs, ps, reps, &f_es);
        int r3 = encode_timed(&amvcuda_amv_encoder_delay, w, h, n, 0, ys, us, vs, pq, reps_q, &f_eq);
        if (r1 || r2 || r3) { printf("FAIL: timed encode returned %d / %d / %d\n", r1, r2, r3); return 12; }
        for (i = 0; i < n; i++)
            if (pr[i].size != ps[i].size || memcmp(pr[i].pk, ps[i].pk, pr[i].size) || pr[i].size != pq[i].size || memcmp(pr[i].pk, pq[i].pk, pr[i].size)) {
                printf("FAIL: timed run, packet %d differs (shim %d, look-ahead %d, reference %d bytes)\n", i, ps[i].size, pq[i].size, pr[i].size); fail = 1;
            }
        r1 = decode_timed(&amv_decoder, w, h, n, pr, dr, reps_ref, &f_dr);
        r2 = decode_timed(&amvcuda_amv_decoder, w, h, n, pr, ds, reps, &f_ds);
        r3 = decode_timed(&amvcuda_amv_decoder_delay, w, h, n, pr, dq, reps_q, &f_dq);
        if (r1 || r2 || r3) { printf("FAIL: timed decode returned %d / %d / %d\n", r1, r2, r3); return 13; }
        if (memcmp(dr, ds, (size_t)n * fb) || memcmp(dr, dq, (size_t)n * fb)) { printf("FAIL: timed run, decoded planes differ\n"); fail = 1; }
        printf("drop-in frames/s %dx%d (one host thread): encode reference %.0f | shim, one frame per call %.0f | shim, look-ahead queue %.0f ;"
               " decode reference %.0f | shim %.0f | look-ahead %.0f ; outputs %s\n", w, h, f_er, f_es, f_eq, f_dr, f_ds, f_dq,
               fail ? "DIFFER" : "identical");
    }
    for (g_trellis = 0; g_trellis <= 3; g_trellis += 3) {      /* plain encoder, then -trellis 3 */
        const int total = 22050 * 2 + 999, fs = 1378;
        int16_t *pcm = malloc(total * 2), *oa = malloc(total * 4), *ob = malloc(total * 4);
        uint8_t *ca = malloc(total), *cb = malloc(total);
        int cba = 0, cbb = 0, sa = 0, sb = 0;
        for (i = 0; i < total; i++) pcm[i] = (int16_t)(8000 * sin(i * 0.1254) + 2000 * sin(i * 0.3516) + (int)(rnd() % 200));
        int ra = audio_roundtrip(avcodec_find_encoder(CODEC_ID_ADPCM_IMA_AMV), avcodec_find_decoder(CODEC_ID_ADPCM_IMA_AMV),
                                 pcm, total, fs, ca, &cba, oa, &sa);
        int rb = audio_roundtrip(&adpcm_ima_amv_encoder, &adpcm_ima_amv_decoder, pcm, total, fs, cb, &cbb, ob, &sb);
        if (ra || rb) { printf("FAIL: audio returned %d / %d\n", ra, rb); return 5; }
        if (cba != cbb || memcmp(ca, cb, cba) || sa != sb || memcmp(oa, ob, sa * 2)) { printf("FAIL: audio differs\n"); fail = 1; }
        printf("audio %d samples (trellis %d): %d chunk bytes, %d decoded samples %s\n", total, g_trellis, cba, sa, fail ? "DIFFER" : "identical");
    }
    {   /* the `-s WxH` scaler the way ffmpeg.c drives it: sws_getContext once, sws_scale per frame (ffmpeg.c:757,1684) */
        static const int dims[][4] = { { 352, 288, 208, 176 }, { 160, 120, 320, 240 }, { 322, 244, 160, 120 }, { 208, 176, 208, 176 } };
        int d, t, p, fm;
        for (fm = 0; fm < 4; fm++)
        for (d = 0; d < 4; d++) {
            const int sf = fm & 1 ? PIX_FMT_YUVJ420P : PIX_FMT_YUV420P, df = fm & 2 ? PIX_FMT_YUVJ420P : PIX_FMT_YUV420P;
            const int iw = dims[d][0], ih = dims[d][1], ow = dims[d][2], oh = dims[d][3];
            const int icw = (iw + 1) / 2, ich = (ih + 1) / 2, ocw = (ow + 1) / 2, och = (oh + 1) / 2, ils = iw + 32, ols = ow + 16;
            uint8_t *src[4] = { malloc(ils * ih), malloc((ils / 2) * ich), malloc((ils / 2) * ich), NULL };
            uint8_t *da[4] = { malloc(ols * oh), malloc((ols / 2) * och), malloc((ols / 2) * och), NULL };
            uint8_t *db[4] = { malloc(ols * oh), malloc((ols / 2) * och), malloc((ols / 2) * och), NULL };
            int sst[4] = { ils, ils / 2, ils / 2, 0 }, dst[4] = { ols, ols / 2, ols / 2, 0 }, sfail = 0;
            struct SwsContext *ca = amvcuda_sws_getContext(iw, ih, sf, ow, oh, df, SWS_BICUBIC, NULL, NULL, NULL);
            struct SwsContext *cb = sws_getContext(iw, ih, sf, ow, oh, df, SWS_BICUBIC, NULL, NULL, NULL);
            if (!ca || !cb) { printf("FAIL: sws_getContext\n"); return 10; }
            for (t = 0; t < 3; t++) {
                for (p = 0; p < 3; p++) {
                    int k, nb = sst[p] * (p ? ich : ih);
                    for (k = 0; k < nb; k++) src[p][k] = (uint8_t)(t == 2 ? (rnd() & 1) * 255 : rnd());
                    memset(da[p], 0x55, dst[p] * (p ? och : oh)); memset(db[p], 0x55, dst[p] * (p ? och : oh));
                }
                if (amvcuda_sws_scale(ca, src, sst, 0, ih, da, dst) || sws_scale(cb, src, sst, 0, ih, db, dst)) { printf("FAIL: sws_scale\n"); return 10; }
                for (p = 0; p < 3; p++) if (memcmp(da[p], db[p], dst[p] * (p ? och : oh))) sfail = 1;
            }
            printf("scale %dx%d %s -> %dx%d %s x3: planes %s\n", iw, ih, fm & 1 ? "yuvj420p" : "yuv420p", ow, oh, fm & 2 ? "yuvj420p" : "yuv420p",
                   sfail ? "DIFFER" : "identical");
            if (sfail) fail = 1;
            amvcuda_sws_freeContext(ca); sws_freeContext(cb);
            (void)icw; (void)ocw;
        }
    }
    {   /* the audio resampler the way do_audio_out drives it: one call per decoded packet, ragged packet sizes */
        static const int cfg[][3] = { { 44100, 2, 1152 }, { 48000, 1, 1024 }, { 8000, 1, 160 }, { 32000, 2, 4608 }, { 22050, 2, 7 } };
        int d, k;
        for (d = 0; d < 5; d++) {
            const int rate = cfg[d][0], ch = cfg[d][1], pkt = cfg[d][2], total = pkt * 37 + 13;
            short *pcm = malloc(sizeof(short) * total * ch), *oa = malloc(sizeof(short) * (total * 3 + 4096)), *ob = malloc(sizeof(short) * (total * 3 + 4096));
            short *ta = malloc(sizeof(short) * (pkt * 16 + 4096)), *tb = malloc(sizeof(short) * (pkt * 16 + 4096));
            ReSampleContext *ra = amvcuda_audio_resample_init(1, ch, 22050, rate), *rb = audio_resample_init(1, ch, 22050, rate);
            int na = 0, nb = 0, pos = 0, afail = 0, call = 0;
            if (!ra || !rb) { printf("FAIL: audio_resample_init\n"); return 11; }
            for (k = 0; k < total * ch; k++) pcm[k] = (short)(k % 97 < 3 ? (rnd() & 1 ? 32767 : -32768) : 9000 * sin(k * 0.0731) + (int)(rnd() % 3000) - 1500);
            while (pos < total) {
                int n = call % 5 == 4 ? pkt / 3 + 1 : pkt, ka, kb;        /* every fifth packet is a short one */
                if (n > total - pos) n = total - pos;
                ka = amvcuda_audio_resample(ra, ta, pcm + pos * ch, n);
                kb = audio_resample(rb, tb, pcm + pos * ch, n);
                if (ka != kb || ka < 0 || memcmp(ta, tb, sizeof(short) * ka)) afail = 1;
                if (ka > 0) { memcpy(oa + na, ta, sizeof(short) * ka); na += ka; }
                if (kb > 0) { memcpy(ob + nb, tb, sizeof(short) * kb); nb += kb; }
                pos += n; call++;
            }
            printf("resample %d Hz x%d -> 22050 mono, %d calls: %d samples, per-call counts and samples %s\n", rate, ch, call, nb, afail ? "DIFFER" : "identical");
            if (afail) fail = 1;
            amvcuda_audio_resample_close(ra); audio_resample_close(rb);
        }
    }
    printf(fail ? "DROP-IN CHECK FAILED\n" : "DROP-IN CHECK OK\n");
    return fail;
}
