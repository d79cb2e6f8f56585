#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "qk_rope or attention or qkv" > gpurun_out/r12_kchecks.log 2>&1; echo "rc=$?" >> gpurun_out/r12_kchecks.log
tail -2 gpurun_out/r12_kchecks.log
SRB_FUSED_QKV=0 timeout 300 python bench.py --no-config3 --ops gpurun_out/r12_ops_split.csv > gpurun_out/r12_bench_split.json 2> gpurun_out/r12_bench_split.err
timeout 300 python bench.py --no-config3 --ops gpurun_out/r12_ops_fused.csv > gpurun_out/r12_bench_fused.json 2> gpurun_out/r12_bench_fused.err
SRB_FUSED_QKV=0 timeout 300 python bench.py --no-config3 > gpurun_out/r12_bench_split2.json 2> /dev/null
timeout 300 python bench.py --no-config3 > gpurun_out/r12_bench_fused2.json 2> /dev/null
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q -x > gpurun_out/r12_e2e.log 2>&1; echo "rc=$?" >> gpurun_out/r12_e2e.log
tail -2 gpurun_out/r12_e2e.log
grep -h "qk_rope\|v_transposed\|attention" gpurun_out/r12_ops_*.csv
python - <<'PY'
import json
for f in ("r12_bench_split","r12_bench_fused","r12_bench_split2","r12_bench_fused2"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["ms_per_step"],3), round(d["value"],1), d["gpu_launches"])
    except Exception as e: print(f, "ERR", e)
PY
