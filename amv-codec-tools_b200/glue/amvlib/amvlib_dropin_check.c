/*
 * amvlib_dropin_check.c -- drives amvlib's decode entry points the way its callers do
 * (AMVDecoderDlg.cpp:368-372, AmvLibTest.cpp:36-85: fill amv->framebuf, AmvVideoDecode, AmvAudioDecode,
 * read amv->videobuf / amv->audiobuf) once with the reference's own functions and once with the
 * libamvcuda bindings (glue/amvlib/amvcuda_amvlib.c), and compares bitmaps and PCM byte for byte.
 * Input: a file of records written by the test: "AMVP" w h nvideo naudio, then per unit le32 size +
 * bytes.  Built by glue/build_dropin.sh where the reference tree is mounted; the binary travels.
 * Exit code 0 = identical.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#ifdef AMVLIB_LONG32
#define long int
#endif
#include "AMVDec.h"
#undef long

int amvcuda_AmvVideoDecode(AMVDecoder *amv);
int amvcuda_AmvAudioDecode(AMVDecoder *amv);
void amvcuda_amvlib_shutdown(void);

static unsigned char *read_unit(FILE *f, uint32_t *size)
{
    unsigned char *p;
    if (fread(size, 4, 1, f) != 1) return NULL;
    p = calloc(1, (size_t)*size + 64);            /* the reference's bit reader / 4-byte groups look past the end */
    if (!p || fread(p, 1, *size, f) != *size) return NULL;
    return p;
}

int main(int argc, char **argv)
{
    FILE *f;
    char magic[4];
    int32_t hdr[4];
    AMVDecoder ref, gpu;
    int i, bad = 0;
    if (argc < 2 || !(f = fopen(argv[1], "rb"))) { fprintf(stderr, "usage: %s packets.bin\n", argv[0]); return 2; }
    if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "AMVP", 4) || fread(hdr, 4, 4, f) != 4) return 2;
    memset(&ref, 0, sizeof ref); memset(&gpu, 0, sizeof gpu);
    ref.opened = gpu.opened = 1;
    ref.amvinfo.dwWidth = gpu.amvinfo.dwWidth = (unsigned)hdr[0];
    ref.amvinfo.dwHeight = gpu.amvinfo.dwHeight = (unsigned)hdr[1];
    ref.amvinfo.nChannels = gpu.amvinfo.nChannels = 1;
    for (i = 0; i < hdr[2]; i++) {
        uint32_t size;
        unsigned char *pk = read_unit(f, &size);
        int r0, r1;
        if (!pk) return 2;
        ref.framebuf.videobuff = gpu.framebuf.videobuff = pk;
        ref.framebuf.videobufflen = gpu.framebuf.videobufflen = size;
        r0 = AmvVideoDecode(&ref);
        r1 = amvcuda_AmvVideoDecode(&gpu);
        if (r0 != r1 || ref.videobuf.len != gpu.videobuf.len ||
            (r0 == 0 && memcmp(ref.videobuf.fbmpdat, gpu.videobuf.fbmpdat, ref.videobuf.len))) {
            printf("video frame %d differs (ret %d / %d)\n", i, r0, r1);
            bad++;
        }
        free(pk);
    }
    for (i = 0; i < hdr[3]; i++) {
        uint32_t size;
        unsigned char *ck = read_unit(f, &size);
        int r0, r1;
        if (!ck) return 2;
        ref.framebuf.audiobuff = gpu.framebuf.audiobuff = ck;
        ref.framebuf.audiobufflen = gpu.framebuf.audiobufflen = size;
        r0 = AmvAudioDecode(&ref);
        r1 = amvcuda_AmvAudioDecode(&gpu);
        if (r0 != r1 || ref.audiobuf.len != gpu.audiobuf.len ||
            (r0 == 0 && memcmp(ref.audiobuf.audiodata, gpu.audiobuf.audiodata, ref.audiobuf.len))) {
            printf("audio chunk %d differs (ret %d / %d, len %u / %u)\n", i, r0, r1, ref.audiobuf.len, gpu.audiobuf.len);
            bad++;
        }
        free(ck);
    }
    fclose(f);
    amvcuda_amvlib_shutdown();
    if (bad) { printf("AMVLIB DROP-IN CHECK FAILED: %d units differ\n", bad); return 1; }
    printf("AMVLIB DROP-IN CHECK OK: %d frames %dx%d, %d audio chunks identical\n", hdr[2], hdr[0], hdr[1], hdr[3]);
    return 0;
}
