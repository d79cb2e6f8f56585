#!/bin/bash
# Build libsrb.so (sm_100a only) in-tree: speech_resynth_b200/libsrb.so
set -e
cd "$(dirname "$0")"
OUT=../libsrb.so
NVCC=${NVCC:-nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O2"
mkdir -p ../../build
pids=()
for f in srb_convgemm srb_elementwise srb_attention srb_attention_tc srb_mrf_fused; do
  $NVCC $FLAGS $EXTRA -c $f.cu -o ../../build/$f.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o $OUT ../../build/srb_convgemm.o ../../build/srb_elementwise.o ../../build/srb_attention.o ../../build/srb_attention_tc.o ../../build/srb_mrf_fused.o -lcudart
echo "built $(realpath $OUT)"
