#!/bin/sh
# TEST-ONLY: builds libamvcuda_emul.so -- the kernel sources of libamvcuda compiled as plain C++ against the SIMT
# emulator in this directory (see cuda_runtime.h).  Never shipped, never loaded by the package.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
CSRC="$HERE/../../../amv-codec-tools_b200/csrc"
OUT="$HERE/_build"
mkdir -p "$OUT"
CXX="${CXX:-g++}"
FLAGS="-O1 -g -rdynamic -std=c++17 -fPIC -DAMV_EMUL -I$HERE -w -fno-strict-aliasing"
pids=""
for f in amv_api amv_dec amv_enc amv_adpcm amv_amvlib amv_container amv_range amv_resample; do
    $CXX $FLAGS -x c++ -include cuda_runtime.h -c "$CSRC/$f.cu" -o "$OUT/$f.o" &
    pids="$pids $!"
done
$CXX $FLAGS -c "$HERE/simt_rt.cpp" -o "$OUT/simt_rt.o" &
pids="$pids $!"
for p in $pids; do wait $p; done
$CXX -shared -o "$OUT/libamvcuda_emul.so" "$OUT"/*.o -lpthread
echo "$OUT/libamvcuda_emul.so"
