cd $GRAFT_REPO_ROOT
timeout 300 ./tools/selftest_gemm > gpurun_out/r02_selftest_gemm_d.log 2>&1; echo "selftest rc=$?"
grep -E "FAIL|PASSED|FAILED|watchdog|rc=|error" gpurun_out/r02_selftest_gemm_d.log | head -20
grep "time " gpurun_out/r02_selftest_gemm_d.log | head -28
