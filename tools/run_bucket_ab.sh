#!/bin/bash
# configs[2] record with different bucket budgets (tile_budget,max_batch,max_waste), same box
mkdir -p gpurun_out
for ba in 296,160,0.12 592,320,0.12 888,480,0.15 1184,640,0.2; do
  tag=$(echo $ba | tr ',' '_')
  timeout 600 python bench.py --steps 3 --warmup 3 --c3-buckets $ba > gpurun_out/bench_c3_$tag.json 2> gpurun_out/bench_c3_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_c3_{tag}.json").read().strip().splitlines()[-1])
    c = d["config3"]
    print(tag, "buckets", c["buckets"], "steady ms %.1f" % c["ms"], "value %.0f" % c["value"], "cold %.0f" % c["cold_value"],
          "arena %.1f GB" % c["arena_gb"], "passes", [round(p["ms"], 1) for p in c["passes"]], "padded TF %.0f" % c["padded_tflop"])
except Exception as e:
    print(tag, "failed", e); print(open(f"gpurun_out/bench_c3_{tag}.err").read()[-800:])
PY
done
