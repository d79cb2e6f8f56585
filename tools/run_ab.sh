#!/bin/bash
# same-box A/B, alternating.  Two forms:
#   tools/run_ab.sh <tag> [kernel-check filters...]            build/libsrb_base.so (SRB_DEBUG_LIB) vs speech_resynth_b200/libsrb.so
#   AB_ENV="SRB_X=0" tools/run_ab.sh <tag> [filters...]        the shipped library with and without the knob
TAG=${1:-ab}; shift
mkdir -p gpurun_out
timeout 600 python tools/gpu_check.py "$@" > gpurun_out/${TAG}_kchecks.log 2>&1; echo "kchecks rc=$?" >> gpurun_out/${TAG}_kchecks.log
tail -4 gpurun_out/${TAG}_kchecks.log
if [ -n "$AB_ENV" ]; then BASE_ENV="$AB_ENV"; else BASE_ENV="SRB_DEBUG_LIB=$PWD/build/libsrb_base.so"; fi
for i in 1 2; do
  env $BASE_ENV timeout 300 python bench.py --no-config3 --steps 20 --ops gpurun_out/ops_${TAG}_base$i.csv > gpurun_out/bench_${TAG}_base$i.json 2> gpurun_out/bench_${TAG}_base$i.err
  timeout 300 python bench.py --no-config3 --steps 20 --ops gpurun_out/ops_${TAG}_new$i.csv > gpurun_out/bench_${TAG}_new$i.json 2> gpurun_out/bench_${TAG}_new$i.err
done
python tools/cmp_ops.py ${TAG}_base1 ${TAG}_new1 ${TAG}_base2 ${TAG}_new2 | head -60
