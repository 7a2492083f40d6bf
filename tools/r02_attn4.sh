cd $GRAFT_REPO_ROOT
timeout 120 ./tools/selftest_attn > gpurun_out/r02_selftest_attn_bwd2.log 2>&1; echo "selftest_attn rc=$?"
cat gpurun_out/r02_selftest_attn_bwd2.log | tail -22
