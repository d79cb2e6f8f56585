#!/bin/bash
# 2-GPU pass: the torchrun world=2 sharded test (NCCL), then the 2-GPU bench (config 2 weak scaling + config 3 strong scaling)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -3 gpurun_out/r2b_pytest.log
bash tools/run_gpu_scale.sh 2
