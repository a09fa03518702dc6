#!/bin/sh
# TEST-ONLY: builds libamvcuda_emul.so -- the kernel sources of libamvcuda compiled as plain C++ against the SIMT
# emulator in this directory (see cuda_runtime.h).  Never shipped, never loaded by the package.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
CSRC="$HERE/../../../amv-codec-tools_b200/csrc"
OUT="$HERE/_build"
CXX="${CXX:-g++}"
FLAGS="-O1 -g -rdynamic -std=c++17 -fPIC -DAMV_EMUL -I$HERE -w -fno-strict-aliasing"
LDFLAGS=""
if [ -n "$ASAN" ]; then     # ASAN=1: the same library under AddressSanitizer ("device" and pinned memory are heap blocks here)
    OUT="$HERE/_build_asan"
    LIBASAN=""
    for c in "$CXX" g++ /usr/bin/g++; do          # a compiler that ships the sanitizer runtime
        a="$($c -print-file-name=libasan.so 2>/dev/null || true)"
        case "$a" in /*) CXX="$c"; LIBASAN="$a"; break;; esac
    done
    [ -n "$LIBASAN" ] || { echo "no libasan.so" >&2; exit 3; }
    FLAGS="$FLAGS -fsanitize=address -fno-omit-frame-pointer"
    LDFLAGS="-fsanitize=address"
fi
mkdir -p "$OUT"
# one build at a time (pytest-xdist workers start this script together)
if command -v flock >/dev/null 2>&1; then exec 9>"$OUT/.lock"; flock 9; fi
pids=""
for f in amv_api amv_dec amv_enc amv_adpcm amv_amvlib amv_container amv_range amv_resample; do
    $CXX $FLAGS -x c++ -include cuda_runtime.h -c "$CSRC/$f.cu" -o "$OUT/$f.o" &
    pids="$pids $!"
done
$CXX $FLAGS -c "$HERE/simt_rt.cpp" -o "$OUT/simt_rt.o" &
pids="$pids $!"
for p in $pids; do wait $p; done
$CXX -shared $LDFLAGS -o "$OUT/libamvcuda_emul.so" "$OUT"/*.o -lpthread
[ -z "$ASAN" ] || echo "$LIBASAN"        # ASAN=1: the runtime to preload, then the library
echo "$OUT/libamvcuda_emul.so"
