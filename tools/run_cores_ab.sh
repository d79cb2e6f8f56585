#!/bin/bash
# co-residency experiment for the vocoder's parallel resblock chains (env knobs only, same box, alternating)
mkdir -p gpurun_out
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --no-config3 --steps 30 --warmup 5 > gpurun_out/bench_co_$tag.json 2> gpurun_out/bench_co_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_co_{tag}.json").read().strip().splitlines()[-1])
    print(f"{tag:28s} ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  sm {d['clocks']['sm_mhz']}")
except Exception as e:
    print(tag, "failed", e); print(open(f"gpurun_out/bench_co_{tag}.err").read()[-600:])
PY
}
run base1 SRB_X=0
run nofork_voc SRB_FORK=1
run ws_occ1 SRB_WS_OCC=1 SRB_WS_A_STAGES=2
run ws_occ1_g128 SRB_WS_OCC=1 SRB_WS_A_STAGES=2 SRB_GEN128_STAGES=2,3
run base2 SRB_X=0
run g128_only SRB_GEN128_STAGES=2,3
run ws_occ2 SRB_WS_OCC=2 SRB_WS_A_STAGES=2
run ws_occ1_b SRB_WS_OCC=1 SRB_WS_A_STAGES=2
