#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2 3; do
  for m in 0 1; do
    SRB_CLUSTER_MODE=$m timeout 300 python bench.py --no-config3 --steps 20 > gpurun_out/r14_bench_cm${m}_$rep.json 2> /dev/null
  done
done
python - <<'PY'
import json
for rep in (1,2,3):
    for m in (0,1):
        d=json.load(open(f"gpurun_out/r14_bench_cm{m}_{rep}.json")); print(rep, "mode", m, round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["clocks"]["sm_mhz"])
PY
