#!/bin/bash
# Build libdexb200.so (sm_100a only) in-tree: dex-tts_b200/dexb200/libdexb200.so
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="$HERE/csrc"
OUT="$HERE/dexb200/libdexb200.so"
OBJ="$HERE/build"
mkdir -p "$OBJ"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall ${DEXB_NVCC_EXTRA:-}"
pids=()
for f in api engine gemm attn posconv kernels_unet kernels_dit kernels_misc kernels_stft tiv tv align mas vocoder; do
  if [ ! -f "$OBJ/$f.o" ] || [ -n "$(find "$SRC" -newer "$OBJ/$f.o" \( -name '*.cu' -o -name '*.cuh' \) -print -quit)" ] || [ "$HERE/../include/dexb200.h" -nt "$OBJ/$f.o" ]; then
    nvcc $FLAGS -c "$SRC/$f.cu" -o "$OBJ/$f.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT.tmp" "$OBJ"/*.o -lcudart
mv -f "$OUT.tmp" "$OUT"          # atomic replace: a snapshot of the tree never sees a half-written library
echo "built $OUT"
