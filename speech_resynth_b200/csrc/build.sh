#!/bin/bash
# Build the sm_100a libraries in-tree:
#   speech_resynth_b200/libsrb.so        the product library
#   speech_resynth_b200/libsrb_tight.so  the same sources with -DSRB_SPLIT (tight-precision mode, see include/srb.h)
#   build/libsrb_trace.so                (`build.sh trace`) instrumented build for tools/trace_*.py
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O2"
SRCS="srb_convgemm srb_elementwise srb_attention_tc srb_mrf_fused srb_pair_fused"
mkdir -p ../../build

compile() {   # <object dir> <extra flags...>
  local obj=$1; shift
  mkdir -p $obj
  for f in $SRCS; do
    $NVCC $FLAGS "$@" -c $f.cu -o $obj/$f.o &
    pids+=($!)
  done
}
link() {      # <object dir> <output>
  local objs=""
  for f in $SRCS; do objs="$objs $1/$f.o"; done
  $NVCC -shared -o $2 $objs -lcudart
  echo "built $(realpath $2)"
}

pids=()
if [ "$1" = "trace" ]; then
  compile ../../build/trace -DSRB_TRACE=1
  for p in "${pids[@]}"; do wait $p; done
  link ../../build/trace ../../build/libsrb_trace.so
  exit 0
fi
compile ../../build/product
[ "$1" = "product" ] || compile ../../build/tight -DSRB_SPLIT=1
for p in "${pids[@]}"; do wait $p; done
link ../../build/product ../libsrb.so
[ "$1" = "product" ] || link ../../build/tight ../libsrb_tight.so
