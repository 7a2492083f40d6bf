cd $GRAFT_REPO_ROOT
timeout 600 ./tools/selftest_attn > gpurun_out/r02_selftest_attn_fwd2.log 2>&1; echo "selftest_attn rc=$?"
cat gpurun_out/r02_selftest_attn_fwd2.log | tail -22
