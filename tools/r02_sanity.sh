cd $GRAFT_REPO_ROOT
timeout -k 5 600 python -m pytest tests -q -m gpu > gpurun_out/r02f_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02f_pytest_gpu.log | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02f_smoke.log | cut -c1-250
