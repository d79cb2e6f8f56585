set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest46.log 2>&1; echo "pytest rc=$?"
timeout 300 python tools/determinism_check.py > gpurun_out/det46.log 2>&1; echo "determinism rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke46.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --ops gpurun_out/ops_final46.csv > gpurun_out/bench46.json 2> gpurun_out/bench46.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref46.json 2> gpurun_out/bench_ref46.err; echo "ref rc=$?"
