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
  # every ENABLE_<codec>_{EN,DE}CODER the compiled files mention: 0 except the AMV path
  on="AMV_DECODER AMV_ENCODER MJPEG_DECODER MJPEG_ENCODER SP5X_DECODER ADPCM_IMA_AMV_DECODER ADPCM_IMA_AMV_ENCODER"
  grep -rhoE "ENABLE_[A-Z0-9_]+" "$REF"/libavcodec/*.c "$REF"/libavcodec/*.h "$REF"/libavutil/*.[ch] \
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
 -I$OUT/cfg -I$REF -I$REF/libavcodec -I$REF/libavutil"

AVCODEC="utils opt imgconvert dsputil simple_idct jfdctint jfdctfst jrevdct faandct \
 mpegvideo mpegvideo_enc mpeg12data mjpeg mjpegenc mjpegdec sp5xdec adpcm bitstream \
 ratecontrol motion_est error_resilience eval h263 jpeglsdec jpegls golomb \
"
AVUTIL="mem log rational mathematics integer intfloat_readwrite crc fifo"

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
gcc $CFLAGS -c "$HERE/ref_harness.c" -o "$OUT/obj/ref_harness.o"
gcc -shared -o "$OUT/libamvref.so" "$OUT/obj/ref_harness.o" $objs -lm -Wl,--no-undefined
echo "built $OUT/libamvref.so"
