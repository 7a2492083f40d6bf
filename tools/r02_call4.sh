set -x
cd $GRAFT_REPO_ROOT
timeout -k 5 900 python -m pytest tests -q -m gpu -s > gpurun_out/r02_pytest_gpu_c.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^E  |OVER BUDGET" gpurun_out/r02_pytest_gpu_c.log | cut -c1-900 | tail -30
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_c.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r02_smoke_c.log | cut -c1-600
