#!/bin/bash
# CTA-pair / multicast modes of the BN = 256 GEMM tiles, per op (same box A/B)
mkdir -p gpurun_out
for m in 0 1 2; do
  SRB_CLUSTER_MODE=$m timeout 300 python bench.py --no-config3 --ops gpurun_out/r13_ops_cm$m.csv > gpurun_out/r13_bench_cm$m.json 2> gpurun_out/r13_bench_cm$m.err
done
SRB_CLUSTER_MODE=0 timeout 300 python bench.py --no-config3 > gpurun_out/r13_bench_cm0b.json 2> /dev/null
python - <<'PY'
import json, csv
for m in ("cm0","cm1","cm2","cm0b"):
    try:
        d=json.load(open(f"gpurun_out/r13_bench_{m}.json")); print(m, round(d["ms_per_step"],3), round(d["value"],1))
    except Exception as e: print(m, "ERR", e)
for m in (0,1,2):
    print("mode", m)
    for r in csv.DictReader(open(f"gpurun_out/r13_ops_cm{m}.csv")):
        if r["op"] in ("srb_cfm_ffn_out_norm","srb_cfm_ffn_glu","srb_cfm_attn_out_norm","srb_cfm_qk_rope","srb_cfm_embed") or ("256, 256" in r["tag"]):
            print("  ", r["op"], r["tag"][:40], r["launches"], r["avg_ms"], r["tflops"])
PY
