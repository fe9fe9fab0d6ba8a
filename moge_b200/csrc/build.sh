#!/bin/bash
# Build libmoge_b200.so for sm_100a (in-tree; the .so travels to the GPU box with the repo snapshot).
set -e
cd "$(dirname "$0")"
OUT=../_lib
mkdir -p $OUT build
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I../../include $MG_EXTRA_FLAGS"
# a change of flags (e.g. MG_EXTRA_FLAGS=-DMG_ATT_DEBUG) invalidates every object
if [ "$(cat build/.flags 2>/dev/null)" != "$FLAGS" ]; then rm -f build/*.o; echo "$FLAGS" > build/.flags; fi
pids=()
for f in umma_rows umma2 umma_tiles conv64 convh attention elementwise pack tmap engine; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ -n "$(find . -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer build/$f.o)" ] || [ ../../include/moge_b200.h -nt build/$f.o ]; then
    ( nvcc $FLAGS -Xptxas -v -c $f.cu -o build/$f.o > build/$f.log 2>&1 || { echo "FAILED $f"; grep -E "error" build/$f.log | head -20; exit 1; } ) &
    pids+=($!)
  fi
done
rc=0
for p in "${pids[@]}"; do wait $p || rc=1; done
[ $rc -eq 0 ] || exit 1
nvcc -shared -o $OUT/libmoge_b200.so build/*.o -lcudart
echo "built $OUT/libmoge_b200.so"
