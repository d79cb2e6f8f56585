#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_units.py -q -x > gpurun_out/units_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/units_pytest.log
tail -25 gpurun_out/units_pytest.log | cut -c1-300
timeout 300 python tools/time_units.py > gpurun_out/units_timing.json 2> gpurun_out/units_timing.err; cat gpurun_out/units_timing.json; tail -3 gpurun_out/units_timing.err
