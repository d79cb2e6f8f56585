#!/bin/bash
# round-2 GPU pass 2 (2 GPUs): full GPU test-suite including the torchrun world=2 sharded test, 2-GPU bench
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 > gpurun_out/r2b_bench_2gpu.json 2> gpurun_out/r2b_bench_2gpu.err; echo "bench2 rc=$?" >> gpurun_out/r2b_bench_2gpu.err
tail -4 gpurun_out/r2b_pytest.log; tail -c 400 gpurun_out/r2b_bench_2gpu.err
