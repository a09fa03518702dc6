/*
 * TEST INFRASTRUCTURE ONLY -- never linked into, or called by, the product.
 *
 * Drives the UNMODIFIED reference container code (AMVmuxer libavformat 51.17.0: amvenc.c, avidec.c,
 * riff.c, aviobuf.c, utils.c; compiled in place by oracle/build_ref.sh into libamvref.so) through
 * libavformat's own public API, the way ffmpeg.c does for `-f amv`:
 *   mux:   av_alloc_format_context, av_new_stream x2 configured like new_video_stream /
 *          new_audio_stream (ffmpeg.c:2770-2960), avcodec_open of the two AMV encoders,
 *          av_set_parameters, av_write_header, av_interleaved_write_frame per packet, av_write_trailer
 *          (ffmpeg.c:1384-1900) into a dynamic memory buffer;
 *   demux: av_open_input_stream on the avi demuxer (AMV files are recognised by its amvh hooks,
 *          avidec.c:237,283,320,429-434) + av_read_frame.
 */
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include "avformat.h"

extern AVOutputFormat amv_muxer;
extern AVInputFormat avi_demuxer;
extern AVCodec amv_encoder, adpcm_ima_amv_encoder;

static int g_inited;
static void fmt_init(void)
{
    if (!g_inited) { avcodec_init(); av_log_set_level(AV_LOG_QUIET); g_inited = 1; }
}

/* returns the file size, or a negative error */
int64_t amvref_mux(int w, int h, int fps, int sample_rate, int n,
                   const uint8_t *vp, const uint64_t *voff, const uint32_t *vsz,
                   const uint8_t *ap, const uint64_t *aoff, const uint32_t *asz,
                   uint8_t *out, uint64_t cap)
{
    fmt_init();
    AVFormatContext *oc = av_alloc_format_context();
    AVFormatParameters params;
    AVStream *vst, *ast;
    uint8_t *buf = NULL;
    int64_t size = -1;
    int i;
    oc->oformat = &amv_muxer;
    vst = av_new_stream(oc, 0);
    ast = av_new_stream(oc, 1);
    avcodec_get_context_defaults2(vst->codec, CODEC_TYPE_VIDEO);
    vst->codec->codec_id = CODEC_ID_AMV;
    vst->codec->codec_type = CODEC_TYPE_VIDEO;
    vst->codec->time_base.num = 1; vst->codec->time_base.den = fps;
    vst->codec->width = w; vst->codec->height = h;
    vst->codec->pix_fmt = PIX_FMT_YUVJ420P;
    avcodec_get_context_defaults2(ast->codec, CODEC_TYPE_AUDIO);
    ast->codec->codec_id = CODEC_ID_ADPCM_IMA_AMV;
    ast->codec->codec_type = CODEC_TYPE_AUDIO;
    ast->codec->sample_rate = sample_rate;
    ast->codec->channels = 1;
    memset(&params, 0, sizeof params);
    if (av_set_parameters(oc, &params) < 0) return -2;
    if (avcodec_open(vst->codec, &amv_encoder) < 0) return -3;
    if (avcodec_open(ast->codec, &adpcm_ima_amv_encoder) < 0) return -4;
    if (url_open_dyn_buf(&oc->pb) < 0) return -5;
    if (av_write_header(oc) < 0) return -6;
    for (i = 0; i < n; i++) {
        AVPacket pkt;
        av_init_packet(&pkt);
        pkt.stream_index = 0; pkt.data = (uint8_t *)vp + voff[i]; pkt.size = (int)vsz[i]; pkt.flags |= PKT_FLAG_KEY;
        pkt.pts = i;                                  /* coded_frame->pts rescaled 1:1 (ffmpeg.c:826-827) */
        if (av_interleaved_write_frame(oc, &pkt) < 0) return -7;
        av_init_packet(&pkt);
        pkt.stream_index = 1; pkt.data = (uint8_t *)ap + aoff[i]; pkt.size = (int)asz[i]; pkt.flags |= PKT_FLAG_KEY;
        if (av_interleaved_write_frame(oc, &pkt) < 0) return -8;
    }
    if (av_write_trailer(oc) < 0) return -9;
    size = url_close_dyn_buf(&oc->pb, &buf);
    if (size >= 0 && (uint64_t)size <= cap) memcpy(out, buf, (size_t)size); else size = -10;
    av_free(buf);
    avcodec_close(vst->codec); avcodec_close(ast->codec);
    return size;
}

/* a read-only "file" in memory behind the reference's ByteIOContext callbacks (aviobuf.c:29-60) */
typedef struct { const uint8_t *p; int64_t size, pos; } MemFile;
static int mem_read(void *o, uint8_t *buf, int n)
{
    MemFile *m = o;
    if (m->pos >= m->size) return 0;
    if (n > m->size - m->pos) n = (int)(m->size - m->pos);
    memcpy(buf, m->p + m->pos, n); m->pos += n;
    return n;
}
static offset_t mem_seek(void *o, offset_t off, int whence)
{
    MemFile *m = o;
    if (whence == AVSEEK_SIZE) return m->size;
    if (whence == SEEK_CUR) off += m->pos; else if (whence == SEEK_END) off += m->size;
    if (off < 0 || off > m->size) return -1;
    m->pos = off;
    return off;
}

/* Walks an AMV file held in memory with the reference demuxer.  info = { width, height, fps (time base den / num),
 * sample_rate, nvideo, naudio }; packet payloads are concatenated into vdata / adata in read order with their sizes.
 * Returns the number of packets read, or a negative error. */
int amvref_demux(const uint8_t *file, uint64_t size, int *info,
                 uint8_t *vdata, uint32_t *vsz, uint8_t *adata, uint32_t *asz, int cap_units, uint64_t cap_bytes)
{
    fmt_init();
    ByteIOContext pb;
    AVFormatContext *ic = NULL;
    AVPacket pkt;
    uint64_t vpos = 0, apos = 0;
    int nv = 0, na = 0, i;
    MemFile mf = { file, (int64_t)size, 0 };
    static unsigned char iobuf[32768];
    if (init_put_byte(&pb, iobuf, sizeof iobuf, 0, &mf, mem_read, NULL, mem_seek) < 0) return -1;
    if (av_open_input_stream(&ic, &pb, "mem.amv", &avi_demuxer, NULL) < 0) return -2;
    memset(info, 0, 6 * sizeof(int));
    for (i = 0; i < (int)ic->nb_streams; i++) {
        AVCodecContext *c = ic->streams[i]->codec;
        if (c->codec_type == CODEC_TYPE_VIDEO) {
            info[0] = c->width; info[1] = c->height;
            info[2] = ic->streams[i]->r_frame_rate.num ? ic->streams[i]->r_frame_rate.num / ic->streams[i]->r_frame_rate.den
                                                       : ic->streams[i]->time_base.den / ic->streams[i]->time_base.num;
            if (c->codec_id != CODEC_ID_AMV) return -3;
        } else if (c->codec_type == CODEC_TYPE_AUDIO) {
            info[3] = c->sample_rate;
            if (c->codec_id != CODEC_ID_ADPCM_IMA_AMV) return -4;
        }
    }
    while (av_read_frame(ic, &pkt) >= 0) {
        int video = ic->streams[pkt.stream_index]->codec->codec_type == CODEC_TYPE_VIDEO;
        if (video) {
            if (nv >= cap_units || vpos + pkt.size > cap_bytes) return -5;
            memcpy(vdata + vpos, pkt.data, pkt.size); vsz[nv++] = pkt.size; vpos += pkt.size;
        } else {
            if (na >= cap_units || apos + pkt.size > cap_bytes) return -5;
            memcpy(adata + apos, pkt.data, pkt.size); asz[na++] = pkt.size; apos += pkt.size;
        }
        av_free_packet(&pkt);
    }
    info[4] = nv; info[5] = na;
    return nv + na;
}
