cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_w2v_heads_gpu.py tests/test_tensor_profiler.py -q -x > gpurun_out/r02_pytest_ctc.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r02_pytest_ctc.log | cut -c1-400
