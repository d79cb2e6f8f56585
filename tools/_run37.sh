set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest37.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest37.log
timeout 300 python tools/determinism_check.py > gpurun_out/det37.log 2>&1; echo "det rc=$?"; tail -4 gpurun_out/det37.log
timeout 600 python bench.py --steps 10 --warmup 3 --ops gpurun_out/ops_r37.csv > gpurun_out/bench_r37.json 2> gpurun_out/bench_r37.err; echo "bench rc=$?"
