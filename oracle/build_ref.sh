#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY (see oracle/README.md).
#
# Builds the *unmodified* reference AMV codec path (AMVmuxer's FFmpeg fork,
# libavcodec 51.47.1) from the sources where they lie under
# /root/reference, into oracle/_ref/libamvref.so.  No reference source is
# copied: gcc reads the .c files in place; the only generated inputs are a
# hand-written config.h (below; the reference's ./configure is NOT run) and
# our own harness oracle/ref_harness.c.
#
# Flags mirror SURVEY.md §8c: generic C paths only (no MMX/asm), so the
# compiled functions are ff_jpeg_fdct_islow, dct_quantize_c, simple_idct_put.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${AMV_REFERENCE_ROOT:-/root/reference}/AMVmuxer/ffmpeg"
OUT="$HERE/_ref"
if [ ! -d "$REF/libavcodec" ]; then
  echo "build_ref: reference tree not found at $REF (nothing built)" >&2
  exit 3
fi
mkdir -p "$OUT/cfg" "$OUT/obj"

# ---- hand-written config.h -------------------------------------------------
{
  echo "/* written by oracle/build_ref.sh -- NOT produced by the reference's configure */"
  echo "#ifndef AMVREF_CONFIG_H"
  echo "#define AMVREF_CONFIG_H"
  echo "#define FFMPEG_CONFIGURATION \"amvcuda-oracle generic C\""
  echo "#define ARCH_GENERIC 1"
  echo "#define HAVE_MALLOC_H 1"
  echo "#define HAVE_MEMALIGN 1"
  echo "#define HAVE_LRINTF 1"
  echo "#define HAVE_LRINT 1"
  echo "#define HAVE_LLRINT 1"
  echo "#define HAVE_ROUND 1"
  echo "#define HAVE_ROUNDF 1"
  echo "#define HAVE_MKSTEMP 1"
  echo "#define HAVE_FAST_64BIT 1"
  echo "#define CONFIG_ENCODERS 1"
  echo "#define CONFIG_DECODERS 1"
  echo "#define CONFIG_MUXERS 1"
  echo "#define CONFIG_DEMUXERS 1"
  echo "#define ENABLE_ENCODERS 1"
  echo "#define ENABLE_DECODERS 1"
  echo "#define ENABLE_SMALL 0"
  echo "#define ENABLE_GRAY 0"
  echo "#define ENABLE_GPL 0"
  echo "#define EXTERN_PREFIX \"\""
  echo "#define restrict __restrict__"
  echo "#define CONFIG_AMV_MUXER 1"
  echo "#define CONFIG_AVI_DEMUXER 1"
  # every ENABLE_<codec>_{EN,DE}CODER the compiled files mention: 0 except the AMV path
  on="AMV_DECODER AMV_ENCODER MJPEG_DECODER MJPEG_ENCODER SP5X_DECODER ADPCM_IMA_AMV_DECODER ADPCM_IMA_AMV_ENCODER"
  on="$on AMV_MUXER AVI_DEMUXER"
  grep -rhoE "ENABLE_[A-Z0-9_]+" "$REF"/libavcodec/*.c "$REF"/libavcodec/*.h "$REF"/libavutil/*.[ch] \
       "$REF"/libavformat/utils.c "$REF"/libavformat/avidec.c "$REF"/libavformat/amvenc.c "$REF"/libavformat/riff.c "$REF"/libavformat/*.h \
    | sort -u | while read -r m; do
      n="${m#ENABLE_}"
      case "$n" in ENCODERS|DECODERS|SMALL|GRAY|GPL) continue;; esac
      v=0; for o in $on; do if [ "$o" = "$n" ]; then v=1; fi; done
      echo "#define $m $v"
      if [ "$v" = 1 ]; then echo "#define CONFIG_$n 1"; fi
    done
  echo "#endif"
} > "$OUT/cfg/config.h"

CFLAGS="-O3 -fPIC -std=gnu99 -fgnu89-inline -fcommon -fno-strict-aliasing -fwrapv -w \
 -DHAVE_AV_CONFIG_H -D_ISOC9X_SOURCE -D_GNU_SOURCE \
 -I$OUT/cfg -I$REF -I$REF/libavcodec -I$REF/libavutil -I$REF/libswscale"

AVCODEC="utils opt imgconvert dsputil simple_idct jfdctint jfdctfst jrevdct faandct \
 mpegvideo mpegvideo_enc mpeg12data mjpeg mjpegenc mjpegdec sp5xdec adpcm bitstream \
 ratecontrol motion_est error_resilience eval h263 jpeglsdec jpegls golomb \
 parser raw imgresample resample resample2"
AVUTIL="mem log rational mathematics integer intfloat_readwrite crc fifo string"
# container layer: the AMV muxer, the AVI/AMV demuxer and what they stand on
AVFORMAT="utils aviobuf avio riff amvenc avidec cutils"

objs=""
for f in $AVCODEC; do
  [ -f "$REF/libavcodec/$f.c" ] || { echo "skip $f"; continue; }
  gcc $CFLAGS -c "$REF/libavcodec/$f.c" -o "$OUT/obj/avc_$f.o"
  objs="$objs $OUT/obj/avc_$f.o"
done
for f in $AVUTIL; do
  [ -f "$REF/libavutil/$f.c" ] || { echo "skip $f"; continue; }
  gcc $CFLAGS -c "$REF/libavutil/$f.c" -o "$OUT/obj/avu_$f.o"
  objs="$objs $OUT/obj/avu_$f.o"
done
for f in $AVFORMAT; do
  [ -f "$REF/libavformat/$f.c" ] || { echo "skip $f"; continue; }
  gcc $CFLAGS -I"$REF/libavformat" -c "$REF/libavformat/$f.c" -o "$OUT/obj/avf_$f.o"
  fobjs="${fobjs:-} $OUT/obj/avf_$f.o"
done
gcc $CFLAGS -c "$HERE/ref_harness.c" -o "$OUT/obj/ref_harness.o"
gcc $CFLAGS -I"$REF/libavformat" -c "$HERE/ref_container_harness.c" -o "$OUT/obj/ref_container_harness.o"
gcc -shared -o "$OUT/libamvref.so" "$OUT/obj/ref_harness.o" "$OUT/obj/ref_container_harness.o" $objs $fobjs -lm -Wl,--no-undefined
echo "built $OUT/libamvref.so"

# ---- amvlib (C-AMVDecoder/amvlib), compiled in place ------------------------
# The library is Win32 code: AmvJpeg.c includes <windows.h>/<io.h> and "AmvDec.h" (the file is
# AMVDec.h), and the sources assume a 32-bit `long` (AMVHeader.h:6 DWORD, AMVDec.c:154, AmvJpeg.c:431).
# Instead of patching a copy, three generated shim headers stand in for the missing ones, and
# amvlib_long32.h (force-included) pulls in the libc headers first and then makes `long` 32 bits
# wide for the library's own code, which is what it was written for.
AMVLIB="${AMV_REFERENCE_ROOT:-/root/reference}/C-AMVDecoder/amvlib"
if [ -d "$AMVLIB" ]; then
  SH="$OUT/cfg/amvlib_shim"
  mkdir -p "$SH"
  cat > "$SH/amvlib_long32.h" <<EOT
/* written by oracle/build_ref.sh */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#define long int
EOT
  cat > "$SH/windows.h" <<EOT
/* written by oracle/build_ref.sh: the few Win32 names AmvJpeg.c mentions (only its JPEG-file ->
   BMP-file helper, which the tests never call, uses the file and Global* functions) */
#ifndef AMVLIB_SHIM_WINDOWS_H
#define AMVLIB_SHIM_WINDOWS_H
#include <stdint.h>
typedef uint32_t DWORD; typedef int32_t LONG; typedef uint16_t WORD; typedef uint8_t BYTE;
typedef char *LPSTR; typedef int HFILE; typedef void *HGLOBAL; typedef int BOOL;
#pragma pack(push, 2)
typedef struct { WORD bfType; DWORD bfSize; WORD bfReserved1, bfReserved2; DWORD bfOffBits; } BITMAPFILEHEADER;
#pragma pack(pop)
typedef struct { DWORD biSize; LONG biWidth, biHeight; WORD biPlanes, biBitCount; DWORD biCompression, biSizeImage;
                 LONG biXPelsPerMeter, biYPelsPerMeter; DWORD biClrUsed, biClrImportant; } BITMAPINFOHEADER, *LPBITMAPINFOHEADER;
typedef struct { BYTE b, g, r, x; } RGBQUAD;
#define HFILE_ERROR (-1)
#define OF_READ 0
#define GHND 0
#define BI_RGB 0
#define MAKEWORD(a, b) ((WORD)(((BYTE)(a)) | ((WORD)((BYTE)(b))) << 8))
static inline HFILE _lopen(const char *n, int m) { (void)n; (void)m; return HFILE_ERROR; }
static inline LONG _llseek(HFILE f, LONG o, int w) { (void)f; (void)o; (void)w; return 0; }
static inline LONG _hread(HFILE f, void *b, LONG n) { (void)f; (void)b; (void)n; return 0; }
static inline HFILE _lclose(HFILE f) { (void)f; return 0; }
static inline HFILE _lcreat(const char *n, int a) { (void)n; (void)a; return HFILE_ERROR; }
static inline LONG _lwrite(HFILE f, const char *b, LONG n) { (void)f; (void)b; (void)n; return 0; }
static inline HGLOBAL GlobalAlloc(int f, DWORD n) { (void)f; return calloc(1, n ? n : 1); }
static inline void *GlobalLock(HGLOBAL h) { return h; }
static inline int GlobalUnlock(HGLOBAL h) { (void)h; return 0; }
static inline HGLOBAL GlobalFree(HGLOBAL h) { free(h); return 0; }
#endif
EOT
  echo "/* written by oracle/build_ref.sh */" > "$SH/io.h"
  echo "#include \"$AMVLIB/AMVDec.h\"" > "$SH/AmvDec.h"
  ACF="-O2 -fPIC -std=gnu99 -fcommon -fno-strict-aliasing -fwrapv -w -include $SH/amvlib_long32.h -I$SH -I$AMVLIB"
  aobjs=""
  for f in AMVDec AmvJpeg AdpcmIma; do
    gcc $ACF -c "$AMVLIB/$f.c" -o "$OUT/obj/amvlib_$f.o"
    aobjs="$aobjs $OUT/obj/amvlib_$f.o"
  done
  gcc $ACF -c "$HERE/amvlib_harness.c" -o "$OUT/obj/amvlib_harness.o"
  gcc -shared -o "$OUT/libamvlibref.so" "$OUT/obj/amvlib_harness.o" $aobjs -lm -Wl,--no-undefined
  echo "built $OUT/libamvlibref.so"
fi
