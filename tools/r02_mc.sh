cd $GRAFT_REPO_ROOT
TETHYS_SELFTEST_MC=1 timeout 300 ./tools/selftest_gemm > gpurun_out/mc_selftest.log 2>&1; echo "selftest rc=$?"
grep -E "engine \(4|mc_|FAIL|PASSED|FAILED|watchdog|error" gpurun_out/mc_selftest.log | head -40 | cut -c1-200
grep -E "time " gpurun_out/mc_selftest.log | grep -v "8192\|qk \|pv \|dec " | cut -c1-170
