#!/bin/bash
# the two captures tools/run_ncu.sh's launch counts miss: the fused MRF kernels and the unit quantiser's score GEMM
mkdir -p gpurun_out /tmp/ncu
export SRB_GRAPHS=0
ncu --set full --clock-control none -k regex:"mrf_fused" -c 2 -o /tmp/ncu/mrf -f python tools/ncu_hbm_kernels.py 1 > gpurun_out/r02_ncu_mrf.log 2>&1
ncu --set full --clock-control none -k regex:"convgemm_kernel|kmeans_decode" -c 2 -o /tmp/ncu/units_gemm -f python tools/time_units.py > gpurun_out/r02_ncu_units_gemm.log 2>&1
python tools/ncu_summary.py "extra captures" /tmp/ncu/mrf.ncu-rep /tmp/ncu/units_gemm.ncu-rep > gpurun_out/r02_ncu_summary_extra.csv
cat gpurun_out/r02_ncu_summary_extra.csv | cut -c1-300
