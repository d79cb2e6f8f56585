#!/bin/bash
# fused resblock pair (srb_hifigan_pair_fused): its kernel checks first (bounded), then a same-box A/B and the e2e goldens
mkdir -p gpurun_out
timeout 300 python tools/gpu_check.py pair_fused > gpurun_out/pair_kchecks_first.log 2>&1; rc=$?; echo "first rc=$rc" >> gpurun_out/pair_kchecks_first.log
tail -12 gpurun_out/pair_kchecks_first.log | cut -c1-300
if [ $rc -ne 0 ]; then exit 0; fi
AB_ENV="SRB_PAIR_FUSED=0" bash tools/run_ab.sh pair pair_fused conv_res_act
timeout 600 python -m pytest tests/test_gpu_e2e.py -q -x -k "golden or config4 or full_size" 2>&1 | tail -3
grep -E "waveform" gpurun_out/parity_errors.txt | tail -6
