#!/bin/bash
# full GPU suite + smoke + determinism at the round's defaults, then one bench line with the per-op table
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.txt
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/suite_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/suite_pytest.log
tail -4 gpurun_out/suite_pytest.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/suite_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/suite_smoke.log
tail -2 gpurun_out/suite_smoke.log
timeout 600 python tools/determinism_check.py > gpurun_out/suite_determinism.log 2>&1; echo "rc=$?" >> gpurun_out/suite_determinism.log
tail -4 gpurun_out/suite_determinism.log
timeout 600 python bench.py --ops gpurun_out/suite_ops.csv > gpurun_out/suite_bench.json 2> gpurun_out/suite_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/suite_bench.json").read().strip().splitlines()[-1])
    print("bench ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["clocks"],
          "whole_step", d["roofline"]["whole_step"], "c3", d["config3"]["ms"], d["config3"]["value"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/suite_bench.err").read()[-1500:])
PY
grep -E "80040, 32, 32|ffn_out_norm" gpurun_out/suite_ops.csv | cut -c1-120
