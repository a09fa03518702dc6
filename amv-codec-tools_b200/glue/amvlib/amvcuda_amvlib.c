/*
 * amvcuda_amvlib.c -- reference-side binding for C-AMVDecoder/amvlib: the two decode entry points
 * of amvlib/AMVDec.h, AmvVideoDecode (AMVDec.c:259-286) and AmvAudioDecode (AMVDec.c:288-340),
 * re-implemented on top of libamvcuda's C ABI.  Same arguments, same buffers (amv->videobuf /
 * amv->audiobuf, malloc'd here and owned by the AMVDecoder exactly like the originals), same return
 * codes (0 ok, -1 bad state / undecodable data, -2 out of memory).  No codec arithmetic happens in
 * this file: it moves buffers and calls the GPU library.
 *
 * Build it INTO amvlib in place of the two functions with -DAMVCUDA_REPLACE_AMVLIB (then AMVDec.c's
 * own definitions must be left out), or next to them under the amvcuda_ prefix (what the check
 * harness glue/amvlib/amvlib_dropin_check.c does to compare both).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "amvcuda.h"
#ifdef AMVLIB_LONG32            /* amvlib was written for a 32-bit `long` (Win32) */
#define long int
#endif
#include "AMVDec.h"          /* the reference's own header */
#undef long

#ifdef AMVCUDA_REPLACE_AMVLIB
#define VIDEO_FN AmvVideoDecode
#define AUDIO_FN AmvAudioDecode
#else
#define VIDEO_FN amvcuda_AmvVideoDecode
#define AUDIO_FN amvcuda_AmvAudioDecode
#endif

/* amvlib has no per-decoder codec state to hang a context on (its own decoder keeps everything in
 * file-scope globals, AmvJpeg.c:429-462), so one lazily created libamvcuda context serves the process */
static amv_ctx *g_ctx;
static amv_ctx *ctx_get(void)
{
    if (!g_ctx && amv_create(NULL, &g_ctx) != AMV_OK) g_ctx = NULL;
    return g_ctx;
}
void amvcuda_amvlib_shutdown(void) { if (g_ctx) { amv_destroy(g_ctx); g_ctx = NULL; } }

int VIDEO_FN(AMVDecoder *amv)
{
    AMVInfo *info;
    FRAMEBUFF *fb;
    VIDEOBUFF *vb;
    amv_ctx *ctx;
    uint64_t off = 0;
    uint32_t size;
    int32_t status = 0;
    int w, h, line_bytes;
    size_t need;

    if (amv == NULL || !amv->opened) return -1;
    fb = &amv->framebuf;
    if (fb->videobuff == NULL || fb->videobufflen == 0) return -1;
    info = &amv->amvinfo;
    vb = &amv->videobuf;
    w = (int)info->dwWidth; h = (int)info->dwHeight;
    line_bytes = (w * 24 + 31) / 32 * 4;                 /* WIDTHBYTES(ImgWidth*24), AmvJpeg.c:420,1526 */
    vb->len = info->dwHeight * info->dwWidth * 3;        /* what the reference reports (AMVDec.c:275) */
    need = (size_t)line_bytes * h;                       /* what its row stride actually needs */
    if (need < vb->len) need = vb->len;
    if (vb->fbmpdat) free(vb->fbmpdat);
    vb->fbmpdat = (unsigned char *)malloc(need);
    if (vb->fbmpdat == NULL) return -2;
    memset(vb->fbmpdat, 0, need);
    if ((ctx = ctx_get()) == NULL) return -1;
    size = fb->videobufflen;
    if (amv_decode_frames_bgr24(ctx, fb->videobuff, size, &off, &size, 1, w, h, vb->fbmpdat, line_bytes, need, &status,
                                AMV_MEM_HOST) != AMV_OK)
        return -1;
    return status ? -1 : 0;
}

int AUDIO_FN(AMVDecoder *amv)
{
    FRAMEBUFF *fb;
    AUDIOBUFF *ab;
    amv_ctx *ctx;
    unsigned char *chunk;
    uint64_t off = 0, pcm_off = 0;
    uint32_t size, ndata, npad;
    int32_t status = 0;
    int rc;

    if (amv == NULL || !amv->opened) return -1;
    fb = &amv->framebuf;
    if (fb->audiobuff == NULL || fb->audiobufflen == 0) return -1;
    if (fb->audiobufflen <= 8) return -1;                /* AdpcmImaDecodeFrame: !buf_size (AdpcmIma.c:217-218) */
    ab = &amv->audiobuf;
    /* The reference reads the step index from ONE byte (AMVDec.c:313) and walks the data in groups of
     * four bytes (AdpcmIma.c:225-237), i.e. past the chunk when its length is not a multiple of four.
     * Hand the GPU decoder the chunk in exactly that shape: byte 3 cleared, data zero-padded. */
    ndata = fb->audiobufflen - 8;
    npad = (ndata + 3u) & ~3u;
    chunk = (unsigned char *)calloc(1, 8 + (size_t)npad);
    if (chunk == NULL) return -2;
    memcpy(chunk, fb->audiobuff, fb->audiobufflen);
    chunk[3] = 0;
    if (ab->audiodata) free(ab->audiodata);
    ab->audiodata = (short *)malloc((size_t)npad * 4 + 16);
    if (ab->audiodata == NULL) { free(chunk); return -2; }
    memset(ab->audiodata, 0, (size_t)npad * 4);
    if ((ctx = ctx_get()) == NULL) { free(chunk); return -1; }
    size = 8 + npad;
    rc = amv_adpcm_dec_chunks(ctx, chunk, size, &off, &size, 1, ab->audiodata, (uint64_t)npad * 2, &pcm_off, &status, AMV_MEM_HOST);
    free(chunk);
    if (rc != AMV_OK || status) return -1;
    ab->len = npad * 4;                                  /* declen: bytes of PCM produced (AMVDec.c:331-335) */
    return 0;
}
