#!/bin/bash
# Builds lib/libunislam_b200.so (C-ABI, no torch dependency) for sm_100a. Usage: csrc/build.sh [extra nvcc flags]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT" "$HERE/_obj"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $@"
pids=()
for f in encode field field_bwd field_tc sampling composite loss adam; do
  nvcc $FLAGS -c "$HERE/$f.cu" -o "$HERE/_obj/$f.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libunislam_b200.so" "$HERE"/_obj/{encode,field,field_bwd,field_tc,sampling,composite,loss,adam}.o -lcudart
echo "built $OUT/libunislam_b200.so"
