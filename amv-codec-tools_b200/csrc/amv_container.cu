// amv_container.cu -- the AMV container on the host: a packet index builder, so the
// GPU entry points can read a whole .amv file in place through (offset, size) arrays, and a muxer
// whose output is byte-identical to the reference's (libavformat/amvenc.c driven the way ffmpeg.c
// drives it).  Pure byte shuffling, no codec arithmetic, no device code; it lives in libamvcuda so
// the reference-side bindings get file I/O and codec calls from one library.
#include <stdint.h>
#include <string.h>

#include "../../include/amvcuda.h"

namespace {

inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline bool tag_is(const uint8_t *p, const char *t) { return memcmp(p, t, 4) == 0; }

struct Writer {
    uint8_t *p; uint64_t cap, pos; bool ok;
    void bytes(const void *src, uint64_t n) {
        if (pos + n <= cap) memcpy(p + pos, src, n); else ok = false;
        pos += n;
    }
    void tag(const char *t) { bytes(t, 4); }
    void le32(uint32_t v) { uint8_t b[4] = { (uint8_t)v, (uint8_t)(v >> 8), (uint8_t)(v >> 16), (uint8_t)(v >> 24) }; bytes(b, 4); }
    void le16(uint32_t v) { uint8_t b[2] = { (uint8_t)v, (uint8_t)(v >> 8) }; bytes(b, 2); }
    void u8(uint32_t v) { uint8_t b = (uint8_t)v; bytes(&b, 1); }
    // start_tag / end_tag of riff.c:27-45: a tag, a size patched when the chunk is closed
    uint64_t start(const char *t) { tag(t); le32(0); return pos; }
    void end(uint64_t start_pos) {
        const uint32_t v = (uint32_t)(pos - start_pos);
        if (start_pos <= cap && start_pos >= 4) { uint8_t *q = p + start_pos - 4; q[0] = (uint8_t)v; q[1] = (uint8_t)(v >> 8); q[2] = (uint8_t)(v >> 16); q[3] = (uint8_t)(v >> 24); }
    }
};

}  // namespace

extern "C" {

// Replaces, for AMV files, the reference's two file walkers: the AVI demuxer with its amvh hooks
// (libavformat/avidec.c:237,283,320,429-434,507-508: stream 0 = video CODEC_ID_AMV, stream 1 = audio
// CODEC_ID_ADPCM_IMA_AMV, chunk sizes of the lists ignored -- real device files leave them 0) and
// amvlib's AmvOpen / AmvReadNextFrame (C-AMVDecoder/amvlib/AMVDec.c:15-238).
AMV_API int amv_file_index(const uint8_t *file, uint64_t size, amv_file_info *info,
                           uint64_t *v_off, uint32_t *v_size, uint64_t *a_off, uint32_t *a_size, uint32_t cap) {
    if (!file || !info) return AMV_ERR_ARG;
    memset(info, 0, sizeof(*info));
    if (size < 12 || !tag_is(file, "RIFF") || !tag_is(file + 8, "AMV ")) return AMV_ERR_ARG;
    uint64_t p = 12;
    bool in_movi = false;
    int nstrf = 0;
    while (p + 8 <= size && !in_movi) {
        const uint8_t *c = file + p;
        const uint32_t csz = rd32(c + 4);
        if (tag_is(c, "LIST")) {                      // descend: list sizes are not trustworthy
            if (p + 12 > size) break;
            if (tag_is(c + 8, "movi")) { in_movi = true; info->movi_offset = p + 8; }
            p += 12;
        } else if (tag_is(c, "amvh")) {
            if (csz < 56 || p + 8 + 56 > size) return AMV_ERR_ARG;
            const uint8_t *h = c + 8;                 // amvenc.c:131-178 / amvlib AMVHeader.h:18-40
            info->us_per_frame = rd32(h);
            info->nb_frames_header = rd32(h + 16);
            info->width = (int)rd32(h + 32); info->height = (int)rd32(h + 36); info->fps = (int)rd32(h + 40);
            info->duration_s = h[52] + 60u * h[53] + 3600u * (h[54] | (h[55] << 8));
            p += 8 + (uint64_t)csz;
        } else if (tag_is(c, "strf")) {
            if (nstrf == 1 && csz >= 16 && p + 8 + 16 <= size) {     // second stream: WAVEFORMAT (riff.c:240-289)
                info->channels = c[8 + 2] | (c[8 + 3] << 8);
                info->sample_rate = (int)rd32(c + 8 + 4);
            }
            nstrf++;
            p += 8 + (uint64_t)csz;
        } else {
            p += 8 + (uint64_t)csz;                   // strh and anything unknown: skip by its own size
        }
    }
    if (!in_movi || info->width <= 0 || info->height <= 0) return AMV_ERR_ARG;
    // strictly alternating 00dc / 01wb chunks without even-byte padding (amvenc.c:300-323), then AMV_END_
    uint32_t nv = 0, na = 0;
    while (p + 8 <= size) {
        const uint8_t *c = file + p;
        const uint32_t csz = rd32(c + 4);
        const bool vid = tag_is(c, "00dc"), aud = tag_is(c, "01wb");
        if (!vid && !aud) break;
        if (p + 8 + (uint64_t)csz > size) { info->truncated = 1; break; }
        if (vid) { if (v_off && nv < cap) { v_off[nv] = p + 8; v_size[nv] = csz; } nv++; }
        else     { if (a_off && na < cap) { a_off[na] = p + 8; a_size[na] = csz; } na++; }
        p += 8 + (uint64_t)csz;
    }
    info->has_end_marker = p + 8 <= size && memcmp(file + p, "AMV_END_", 8) == 0;
    info->nvideo = nv; info->naudio = na;
    return AMV_OK;
}

// Replaces amv_muxer (libavformat/amvenc.c: avi_write_header :116-284, avi_write_packet :287-323,
// avi_write_trailer + avi_write_counters :72-114,325-343) fed the way ffmpeg.c feeds it: one video
// packet and one audio chunk per frame, alternating (amv_interleave_packet :378-406).
AMV_API int64_t amv_file_mux(const amv_mux_params *mp, int n,
                             const uint8_t *vpk, const uint64_t *v_off, const uint32_t *v_size,
                             const uint8_t *apk, const uint64_t *a_off, const uint32_t *a_size,
                             uint8_t *out, uint64_t cap) {
    if (!mp || n < 0 || (n > 0 && (!vpk || !v_off || !v_size || !apk || !a_off || !a_size)) || (!out && cap)) return AMV_ERR_ARG;
    if (mp->width <= 0 || mp->height <= 0 || mp->tb_num <= 0 || mp->tb_den <= 0 || mp->sample_rate <= 0) return AMV_ERR_ARG;
    const int vbr = mp->video_bit_rate ? mp->video_bit_rate : 200000;      // AVOption defaults "b" / "ab" (utils.c:429-)
    const int abr = mp->audio_bit_rate ? mp->audio_bit_rate : 64000;
    Writer w = { out, cap, 0, true };
    const uint64_t riff = w.start("RIFF");
    w.tag("AMV ");
    const uint64_t hdrl = w.start("LIST");
    w.tag("hdrl");
    w.tag("amvh"); w.le32(14 * 4);
    w.le32((uint32_t)(1000000ll * mp->tb_num / mp->tb_den));
    w.le32((uint32_t)((vbr + abr) / 8));
    w.le32(0);
    w.le32(0x800 | 0x10 | 0x100);                      // AMVF_TRUSTCKTYPE | AMVF_HASINDEX | AMVF_ISINTERLEAVED
    w.le32((uint32_t)n);                               // frames (avi_write_counters)
    w.le32(0); w.le32(2); w.le32(1024 * 1024);
    w.le32((uint32_t)mp->width); w.le32((uint32_t)mp->height);
    w.le32((uint32_t)mp->tb_den); w.le32(1); w.le32(0);
    const int dur = n / mp->tb_den;                    // the "HACK" of amvenc.c:96-107
    w.u8((uint32_t)(dur % 60)); w.u8((uint32_t)(dur / 60)); w.le16((uint32_t)(dur / 3600));
    {   // video stream
        const uint64_t strl = w.start("LIST"); w.tag("strl");
        const uint64_t strh = w.start("strh");
        w.tag("vids"); w.le32(0 /* codec_tag: CODEC_ID_AMV has no bmp tag */); w.le32(0); w.le16(0); w.le16(0); w.le32(0);
        w.le32((uint32_t)mp->tb_num); w.le32((uint32_t)mp->tb_den); w.le32(0); w.le32((uint32_t)n);
        w.le32(1024 * 1024); w.le32(0xffffffffu); w.le32(0); w.le32(0);
        w.le16((uint32_t)mp->width); w.le16((uint32_t)mp->height);
        w.end(strh);
        const uint64_t strf = w.start("strf");
        for (int i = 0; i < 9; i++) w.le32(0);
        w.end(strf);
        w.end(strl);
    }
    {   // audio stream
        const uint64_t strl = w.start("LIST"); w.tag("strl");
        const uint64_t strh = w.start("strh");
        w.tag("auds"); w.le32(1); w.le32(0); w.le16(0); w.le16(0); w.le32(0);
        w.le32((uint32_t)mp->tb_num); w.le32((uint32_t)mp->tb_den);       // amvenc.c:203-208: the VIDEO time base
        w.le32(0); w.le32((uint32_t)n);
        w.le32(2); w.le32(0); w.le16(0); w.le16(0);
        w.end(strh);
        const uint64_t strf = w.start("strf");
        w.le16(1); w.le16(1); w.le32((uint32_t)mp->sample_rate); w.le32((uint32_t)(abr / 8));    // put_wav_header, riff.c:240-289
        w.le16(2); w.le16(16);
        w.le32(0);
        w.end(strf);
        w.end(strl);
    }
    w.end(hdrl);
    const uint64_t movi = w.start("LIST");
    w.tag("movi");
    for (int i = 0; i < n; i++) {
        w.tag("00dc"); w.le32(v_size[i]); w.bytes(vpk + v_off[i], v_size[i]);
        w.tag("01wb"); w.le32(a_size[i]); w.bytes(apk + a_off[i], a_size[i]);
    }
    w.end(movi);
    w.tag("AMV_"); w.tag("END_");
    w.end(riff);
    if (!w.ok) return -(int64_t)w.pos;                 // too small: -(bytes needed)
    return (int64_t)w.pos;
}

}  // extern "C"
