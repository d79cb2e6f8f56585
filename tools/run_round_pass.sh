#!/bin/bash
# round-2 GPU pass 1: tests, smoke, bench (config 2 + 3), configs 4 / 5, calibration
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke.log
timeout 600 python bench.py --ops gpurun_out/r2_ops.csv > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?" >> gpurun_out/r2_bench.err
timeout 400 python tools/calibrate_buckets.py > gpurun_out/r2_calib.jsonl 2> gpurun_out/r2_calib.err
timeout 400 python bench.py --config 4 --ops gpurun_out/r2_ops_c4.csv > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err
timeout 600 python bench.py --config 5 --ops gpurun_out/r2_ops_c5.csv > gpurun_out/r2_bench_c5.json 2> gpurun_out/r2_bench_c5.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
tail -3 gpurun_out/r2_pytest.log; cat gpurun_out/r2_smoke.log | tail -2; tail -c 600 gpurun_out/r2_bench.err
