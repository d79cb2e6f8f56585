#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_check.py pair_fused > gpurun_out/pair_kchecks_first.log 2>&1; rc=$?; echo "first rc=$rc" >> gpurun_out/pair_kchecks_first.log
tail -3 gpurun_out/pair_kchecks_first.log | cut -c1-200
if [ $rc -ne 0 ]; then exit 0; fi
SRB_PAIR_FUSED=0 timeout 300 python bench.py --no-config3 --steps 20 --ops gpurun_out/ops_pq_base1.csv > gpurun_out/bench_pq_base1.json 2> gpurun_out/bench_pq_base1.err
timeout 300 python bench.py --no-config3 --steps 20 --ops gpurun_out/ops_pq_new1.csv > gpurun_out/bench_pq_new1.json 2> gpurun_out/bench_pq_new1.err
python tools/cmp_ops.py pq_base1 pq_new1 | grep -E "ms/step|pair_fused|40020, 64, 64|total ms"
