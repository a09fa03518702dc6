/*
 * TEST INFRASTRUCTURE ONLY -- never linked into, or called by, the product.
 *
 * Batch driver around the UNMODIFIED reference amvlib (C-AMVDecoder/amvlib: AMVDec.c, AmvJpeg.c,
 * AdpcmIma.c), compiled in place by oracle/build_ref.sh into oracle/_ref/libamvlibref.so.  It goes
 * through amvlib's own decode entry points AmvVideoDecode / AmvAudioDecode (AMVDec.c:259-340) on an
 * AMVDecoder whose frame buffer we fill from memory instead of AmvReadNextFrame's fopen/fread
 * (AMVDec.c:150-238) -- the file walker is not on the codec path.
 *
 * amvlib keeps all JPEG decoder state in file-scope globals (AmvJpeg.c:429-462): one call at a time.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "AMVDec.h"

const char *amvlibref_version(void) { return "amv-codec-tools C-AMVDecoder/amvlib (AmvJpeg.c, AdpcmIma.c), compiled in place"; }

/* bytes per bitmap row the reference writes with: WIDTHBYTES(w*24) (AmvJpeg.c:420,1526) */
int amvlibref_line_bytes(int w) { return (w * 24 + 31) / 32 * 4; }

/* Decode n video packets to bottom-up BGR24 bitmaps of line_bytes(w) * h bytes each.
 * The reference allocates only w*h*3 bytes (AMVDec.c:275-283) but strides rows by WIDTHBYTES(w*24);
 * widths with w % 4 != 0 would overrun its buffer, so they are refused here.  ret[i] = the
 * reference's return value (0 ok, -1 format error). */
int amvlibref_video_decode(const uint8_t *pkts, const uint64_t *off, const uint32_t *size, int n,
                           int w, int h, uint8_t *bgr, int *ret)
{
    AMVDecoder amv;
    int i, lb = amvlibref_line_bytes(w);
    if ((w & 3) || w <= 0 || h <= 0) return -1;
    memset(&amv, 0, sizeof amv);
    amv.opened = 1;
    amv.amvinfo.dwWidth = (unsigned)w;
    amv.amvinfo.dwHeight = (unsigned)h;
    for (i = 0; i < n; i++) {
        /* slack after the packet: the bit reader looks ahead past EOI */
        unsigned char *copy = calloc(1, (size_t)size[i] + 64);
        if (!copy) return -2;
        memcpy(copy, pkts + off[i], size[i]);
        amv.framebuf.videobuff = copy;
        amv.framebuf.videobufflen = size[i];
        ret[i] = AmvVideoDecode(&amv);
        if (amv.videobuf.fbmpdat) memcpy(bgr + (size_t)i * lb * h, amv.videobuf.fbmpdat, (size_t)lb * h);
        free(copy);
    }
    free(amv.videobuf.fbmpdat);
    return n;
}

/* Decode n audio chunks through AmvAudioDecode.  pcm_off[i] = first sample of chunk i in pcm;
 * nsamp[i] receives the number of samples the reference produced (it decodes whole groups of four
 * data bytes, AdpcmIma.c:225-237, i.e. up to 6 samples from bytes past the chunk: the harness pads
 * the chunk with zero bytes so those are defined). */
int amvlibref_audio_decode(const uint8_t *chunks, const uint64_t *off, const uint32_t *size, int n,
                           int16_t *pcm, const uint64_t *pcm_off, uint32_t *nsamp, int *ret)
{
    AMVDecoder amv;
    int i;
    memset(&amv, 0, sizeof amv);
    amv.opened = 1;
    amv.amvinfo.nChannels = 1;
    for (i = 0; i < n; i++) {
        unsigned char *copy = calloc(1, (size_t)size[i] + 16);
        if (!copy) return -2;
        memcpy(copy, chunks + off[i], size[i]);
        amv.framebuf.audiobuff = copy;
        amv.framebuf.audiobufflen = size[i];
        ret[i] = AmvAudioDecode(&amv);
        nsamp[i] = 0;
        if (ret[i] == 0 && amv.audiobuf.audiodata) {
            nsamp[i] = amv.audiobuf.len / 2;
            memcpy(pcm + pcm_off[i], amv.audiobuf.audiodata, amv.audiobuf.len);
        }
        free(copy);
    }
    free(amv.audiobuf.audiodata);
    return n;
}
