#!/bin/bash
# what paces srb_hifigan_pair_fused: the launch with its MMAs / its output stores removed (timing only)
mkdir -p gpurun_out
( for k in 3 7; do for d in 0 1 2 3; do SRB_PAIR_DEBUG=$d timeout 120 python tools/time_op.py pair $k; done; done ) 2>&1 | grep -v Warn | tee gpurun_out/pair_probe.log
