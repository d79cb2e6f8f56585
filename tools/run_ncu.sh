#!/bin/bash
# ncu evidence of the round (run under gpurun, one GPU): (1) launch list of the bench command, (2) --set full captures of
# the dominant GEMM and of the CUDA-core (HBM-bound) kernels at the config-2 shape.  Numbers printed by runs under ncu are
# never bench values.
mkdir -p gpurun_out
export SRB_GRAPHS=0
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_ncu_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-config3 > gpurun_out/r02_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"stage_inputs|embed_gather|unit_lengths|posconv_norm|post_tanh" -c 12 \
    -o gpurun_out/r02_hbm_kernels -f python tools/ncu_hbm_kernels.py 2 > gpurun_out/r02_ncu_hbm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"convgemm_kernel" -s 20 -c 30 \
    -o gpurun_out/r02_convgemm -f python tools/ncu_hbm_kernels.py 1 > gpurun_out/r02_ncu_convgemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"attn_tc_kernel|mrf_fused" -c 6 \
    -o gpurun_out/r02_attn_mrf -f python tools/ncu_hbm_kernels.py 1 > gpurun_out/r02_ncu_attn_mrf.log 2>&1
ls -la gpurun_out/*.ncu-rep
