#!/bin/bash
# BASELINE configs[3] (vocoder alone, 256 x 500) and configs[4] (16 x 3000, NFE sweep) bench lines + the reference arm
mkdir -p gpurun_out
timeout 500 python bench.py --config 4 --ops gpurun_out/c4_ops.csv > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err; echo "c4 rc=$?"
timeout 700 python bench.py --config 5 --ops gpurun_out/c5_ops.csv > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err; echo "c5 rc=$?"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/ref_bench.json 2> gpurun_out/ref_bench.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ("c4_bench", "c5_bench", "ref_bench"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, {k: d[k] for k in ("value", "ms_per_step", "unit") if k in d}, str(d.get("config", ""))[:160])
        for k in ("sweep", "stages", "roofline"):
            if k in d: print("   ", k, str(d[k])[:700])
    except Exception as e:
        print(f, "failed", e)
PY
