cd $GRAFT_REPO_ROOT
timeout 300 ./tools/selftest_attn prof > gpurun_out/r02_attn_prof_plain.log 2>&1; tail -2 gpurun_out/r02_attn_prof_plain.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_fwd2 -c 1 -o gpurun_out/r02_attn_fwd2 -f ./tools/selftest_attn prof > gpurun_out/r02_attn_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02_attn_ncu.log
