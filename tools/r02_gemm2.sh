cd $GRAFT_REPO_ROOT
timeout 300 ./tools/selftest_gemm > gpurun_out/r02_selftest_gemm_b.log 2>&1; echo "selftest rc=$?"
grep -E "FAIL|PASSED|FAILED|watchdog|rc=|error" gpurun_out/r02_selftest_gemm_b.log | head -20
grep "time " gpurun_out/r02_selftest_gemm_b.log
for args in "6000 768 768 0 1 1 cold" "6000 768 3072 0 1 1 cold" "6000 768 3072 0 1 2 cold" "6000 3072 768 0 1 2 cold"; do
  timeout 120 python tools/gemm_trace.py $args 2>&1 | tail -9
done > gpurun_out/r02_gemm_trace_b.log 2>&1
cat gpurun_out/r02_gemm_trace_b.log
