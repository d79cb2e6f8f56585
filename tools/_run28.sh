set -x
cd $GRAFT_REPO_ROOT
SRB_POSCONV_ROWS=8 timeout 300 python tools/gpu_check.py posconv > gpurun_out/check28.log 2>&1; tail -2 gpurun_out/check28.log
SRB_POSCONV_ROWS=8 timeout 600 python bench.py --steps 5 --warmup 3 --ops gpurun_out/ops_p8.csv > gpurun_out/bench_p8.json 2> gpurun_out/bench_p8.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --ops gpurun_out/ops_p16.csv > gpurun_out/bench_p16.json 2> gpurun_out/bench_p16.err; echo "bench rc=$?"
tail -2 gpurun_out/bench_p16.err
