cd $GRAFT_REPO_ROOT
timeout 600 ./tools/selftest_attn > gpurun_out/r02_selftest_attn_fwd2.log 2>&1; echo "selftest_attn rc=$?"
cat gpurun_out/r02_selftest_attn_fwd2.log | tail -30
TETHYS_ATTN_FWD=1 timeout 300 ./tools/selftest_attn prof > gpurun_out/r02_selftest_attn_fwd1.log 2>&1
cat gpurun_out/r02_selftest_attn_fwd1.log | tail -3
