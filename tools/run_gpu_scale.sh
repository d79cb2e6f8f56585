#!/bin/bash
# scaling pass: bench.py (config 2 weak scaling + config 3 strong scaling record) at N GPUs of one box
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus $N > gpurun_out/r2s_bench_${N}gpu.json 2> gpurun_out/r2s_bench_${N}gpu.err; echo "rc=$?" >> gpurun_out/r2s_bench_${N}gpu.err
tail -c 300 gpurun_out/r2s_bench_${N}gpu.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2s_bench_${N}gpu.json"))
print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["e2e"]["value"])
c=d["config3"]; print({k:c[k] for k in c if k not in ("passes","workload","bucket_shapes_rank0")})
for p in c["passes"]: print({k:(v if not isinstance(v,list) else [round(x,1) for x in v]) for k,v in p.items()})
PY
