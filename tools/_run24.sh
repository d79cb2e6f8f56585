set -x
cd $GRAFT_REPO_ROOT
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k "regex:convgemm_kernelILi256ELi64ELi[123]E|attn_tc|posconv" -c 7 -f -o gpurun_out/prof_final_a python tools/profile_step.py 64 500 1 > gpurun_out/ncu24a.log 2>&1; echo "ncu a rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k "regex:convgemm_kernelILi64ELi64ELi0ELi0ELi1E|mrf_fused" -c 4 -f -o gpurun_out/prof_final_b python tools/profile_step.py 64 500 1 > gpurun_out/ncu24b.log 2>&1; echo "ncu b rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:convgemm_kernel|attn_tc_kernel|mrf_fused_kernel|posconv_norm_kernel|post_tanh_kernel|embed_gather_kernel|unit_lengths_kernel|prior_prepare_kernel" --launch-skip 418 -c 836 --csv --log-file gpurun_out/launches24.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch24.log 2>&1; echo "ncu rc=$?"
