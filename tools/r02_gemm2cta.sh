set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
TETHYS_GEMM_CTAS=1 timeout 300 ./tools/selftest_gemm > gpurun_out/r02_selftest_gemm_pair.log 2>&1; echo "selftest rc=$?"
grep -E "FAIL|PASSED|FAILED|watchdog|rc=|error" gpurun_out/r02_selftest_gemm_pair.log | head -20
grep "time " gpurun_out/r02_selftest_gemm_pair.log
