cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_layers_gpu.py tests/test_w2v_gpu.py tests/test_fullsize_properties_gpu.py -q -x > gpurun_out/r02_pytest_layers.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r02_pytest_layers.log | cut -c1-300
