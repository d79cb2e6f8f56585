#!/bin/bash
# ncu evidence of the round (run under gpurun, one GPU): (1) launch list of the bench command, (2) --set full captures of
# the GEMM / attention / MRF kernels and of the CUDA-core (HBM-bound) kernels at the config-2 shape, condensed to CSV on
# the box (the reports themselves are too large to travel).  Numbers printed by runs under ncu are never bench values.
mkdir -p gpurun_out /tmp/ncu
export SRB_GRAPHS=0
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_ncu_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-config3 > gpurun_out/r02_ncu_launches.log 2>&1
ncu --set full --clock-control none -k regex:"stage_inputs|embed_gather|unit_lengths|posconv_norm|post_tanh" -c 10 \
    -o /tmp/ncu/hbm_kernels -f python tools/ncu_hbm_kernels.py 2 > gpurun_out/r02_ncu_hbm.log 2>&1
ncu --set full --clock-control none -k regex:"convgemm_kernel" -s 4 -c 12 \
    -o /tmp/ncu/convgemm_step -f python tools/ncu_hbm_kernels.py 1 > gpurun_out/r02_ncu_convgemm.log 2>&1
ncu --set full --clock-control none -k regex:"attn_tc_kernel|mrf_fused" -c 4 \
    -o /tmp/ncu/attn_mrf -f python tools/ncu_hbm_kernels.py 1 > gpurun_out/r02_ncu_attn_mrf.log 2>&1
ncu --set full --clock-control none -k regex:"convgemm_kernel|split_rows|kmeans_decode" -c 3 \
    -o /tmp/ncu/units -f python tools/time_units.py > gpurun_out/r02_ncu_units.log 2>&1
python tools/ncu_summary.py "round 2, config-2 shape (64 x 500 units), ncu --set full --clock-control none; cold-cache, serialised launches" \
    /tmp/ncu/hbm_kernels.ncu-rep /tmp/ncu/convgemm_step.ncu-rep /tmp/ncu/attn_mrf.ncu-rep /tmp/ncu/units.ncu-rep > gpurun_out/r02_ncu_summary.csv
wc -l gpurun_out/r02_ncu_summary.csv gpurun_out/r02_ncu_launches.csv
