cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest50.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/pytest50.log
timeout 300 python tools/determinism_check.py > gpurun_out/det50.log 2>&1; echo "determinism rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke50.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --ops gpurun_out/ops_final50.csv > gpurun_out/bench50.json 2> gpurun_out/bench50.err; echo "bench rc=$?"
wc -l gpurun_out/bench50.json
