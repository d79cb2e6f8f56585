#!/bin/bash
# full GPU suite + smoke + determinism at the round's defaults
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r15_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r15_pytest.log
tail -3 gpurun_out/r15_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r15_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r15_smoke.log
tail -2 gpurun_out/r15_smoke.log
timeout 600 python tools/determinism_check.py > gpurun_out/r15_determinism.log 2>&1; echo "rc=$?" >> gpurun_out/r15_determinism.log
tail -5 gpurun_out/r15_determinism.log
