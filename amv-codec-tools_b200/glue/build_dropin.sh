#!/usr/bin/env bash
# Builds glue/_build/dropin_check: the AVCodec shims (glue/ffmpeg/amvcuda_codecs.c) compiled against
# the reference's own headers, linked with the reference's unmodified libavcodec objects (produced
# by oracle/build_ref.sh) and with libamvcuda.so.  Needs the reference tree; the binary travels.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF="${AMV_REFERENCE_ROOT:-/root/reference}/AMVmuxer/ffmpeg"
OBJ="$ROOT/oracle/_ref/obj"
[ -d "$REF/libavcodec" ] || { echo "build_dropin: reference tree absent, keeping prebuilt binary" >&2; exit 0; }
[ -d "$OBJ" ] || "$ROOT/oracle/build_ref.sh" > /dev/null
mkdir -p "$HERE/_build"
CFLAGS="-O2 -std=gnu99 -fgnu89-inline -fcommon -w -DHAVE_AV_CONFIG_H -D_GNU_SOURCE \
 -I$ROOT/oracle/_ref/cfg -I$REF -I$REF/libavcodec -I$REF/libavutil -I$REF/libswscale -I$ROOT/include"
gcc $CFLAGS -c "$HERE/ffmpeg/amvcuda_codecs.c" -o "$HERE/_build/amvcuda_codecs.o"
gcc $CFLAGS -c "$HERE/ffmpeg/amvcuda_resample.c" -o "$HERE/_build/amvcuda_resample.o"
# the checker is an API *user* of libavcodec (no HAVE_AV_CONFIG_H: that poisons printf/malloc)
gcc ${CFLAGS/-DHAVE_AV_CONFIG_H/} -c "$HERE/ffmpeg/dropin_check.c" -o "$HERE/_build/dropin_check.o"
OBJS=$(ls "$OBJ"/avc_*.o "$OBJ"/avu_*.o)
gcc -o "$HERE/_build/dropin_check" "$HERE/_build/dropin_check.o" "$HERE/_build/amvcuda_codecs.o" "$HERE/_build/amvcuda_resample.o" $OBJS \
    -L"$HERE/../lib" -lamvcuda -Wl,-rpath,'$ORIGIN/../../lib' -lm
echo "built $HERE/_build/dropin_check"

# ---- amvlib binding: glue/amvlib/amvcuda_amvlib.c against amvlib's own header, checked next to the
# reference's unmodified AmvVideoDecode / AmvAudioDecode
AMVLIB="${AMV_REFERENCE_ROOT:-/root/reference}/C-AMVDecoder/amvlib"
SH="$ROOT/oracle/_ref/cfg/amvlib_shim"
if [ -d "$AMVLIB" ] && [ -f "$OBJ/amvlib_AMVDec.o" ]; then
  ACF="-O2 -std=gnu99 -fcommon -w -DAMVLIB_LONG32 -I$AMVLIB -I$ROOT/include"
  gcc $ACF -c "$HERE/amvlib/amvcuda_amvlib.c" -o "$HERE/_build/amvcuda_amvlib.o"
  gcc $ACF -c "$HERE/amvlib/amvlib_dropin_check.c" -o "$HERE/_build/amvlib_dropin_check.o"
  gcc -o "$HERE/_build/amvlib_dropin_check" "$HERE/_build/amvlib_dropin_check.o" "$HERE/_build/amvcuda_amvlib.o" \
      "$OBJ/amvlib_AMVDec.o" "$OBJ/amvlib_AmvJpeg.o" "$OBJ/amvlib_AdpcmIma.o" \
      -L"$HERE/../lib" -lamvcuda -Wl,-rpath,'$ORIGIN/../../lib' -lm
  echo "built $HERE/_build/amvlib_dropin_check"
fi
