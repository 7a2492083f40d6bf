set -x
cd $GRAFT_REPO_ROOT
timeout -k 5 900 python -m pytest tests -q -m gpu -x -s > gpurun_out/r02_pytest_gpu_a.log 2>&1; echo "pytest rc=$?"
grep -E "^\[|passed|failed|Error|error" gpurun_out/r02_pytest_gpu_a.log | cut -c1-600 | tail -60
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_a.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r02_smoke_a.log
