#!/bin/bash
# tight-precision mode: its tests, then the product library's kernel checks and a short bench (nothing may have moved)
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.txt
timeout 900 python -m pytest tests/test_gpu_tight.py -q -x > gpurun_out/tight_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tight_pytest.log
tail -30 gpurun_out/tight_pytest.log
cat gpurun_out/parity_errors.txt
timeout 600 python tools/gpu_check.py > gpurun_out/tight_kchecks.log 2>&1; echo "kchecks rc=$?" >> gpurun_out/tight_kchecks.log
tail -3 gpurun_out/tight_kchecks.log
timeout 300 python bench.py --no-config3 --steps 20 --ops gpurun_out/ops_tight_pass.csv > gpurun_out/bench_tight_pass.json 2> gpurun_out/bench_tight_pass.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_tight_pass.json").read().strip().splitlines()[-1])
    print("bench ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d.get("clocks"))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_tight_pass.err").read()[-1500:])
PY
