#!/bin/bash
# Builds lib/libunislam_b200.so (C-ABI, no torch dependency) for sm_100a.
# Usage: csrc/build.sh [extra nvcc flags]      (USL_LIB_NAME=libfoo.so csrc/build.sh -DX=1 builds a side-by-side variant
#                                               for A/B runs: select it with USL_LIB_PATH, see _lib.py)
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../lib"
NAME="${USL_LIB_NAME:-libunislam_b200.so}"
OBJ="$HERE/_obj/${NAME%.so}"
mkdir -p "$OUT" "$OBJ"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $@"
SRCS="encode field field_bwd field_tc sampling composite loss adam collective mesh cull metrics"
pids=()
for f in $SRCS; do
  nvcc $FLAGS -c "$HERE/$f.cu" -o "$OBJ/$f.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
OBJS=""
for f in $SRCS; do OBJS="$OBJS $OBJ/$f.o"; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/$NAME" $OBJS -lcudart
echo "built $OUT/$NAME"
