#!/bin/bash
# Build libsrb.so (sm_100a only) in-tree: speech_resynth_b200/libsrb.so
set -e
cd "$(dirname "$0")"
OUT=../libsrb.so
NVCC=${NVCC:-nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O2"
if [ "$1" = "trace" ]; then
  # instrumented build for tools/trace_kernels.py (kept out of the package directory)
  EXTRA="$EXTRA -DSRB_TRACE=1"
  OUT=../../build/libsrb_trace.so
  mkdir -p ../../build/trace
  OBJ=../../build/trace
else
  OBJ=../../build
fi
mkdir -p ../../build
pids=()
for f in srb_convgemm srb_elementwise srb_attention_tc srb_mrf_fused; do
  $NVCC $FLAGS $EXTRA -c $f.cu -o $OBJ/$f.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o $OUT $OBJ/srb_convgemm.o $OBJ/srb_elementwise.o $OBJ/srb_attention_tc.o $OBJ/srb_mrf_fused.o -lcudart
echo "built $(realpath $OUT)"
