cd $GRAFT_REPO_ROOT
timeout 200 python tools/profile_step.py > /dev/null 2>&1 || exit 1
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gn_gelu_bwd1 -s 6 -c 1 -o gpurun_out/r02e_gnbwd1 -f python tools/profile_step.py > gpurun_out/r02e_ncu_gnbwd1.log 2>&1; echo "ncu gnbwd1 rc=$?"
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv0_fwd -c 1 -o gpurun_out/r02e_conv0 -f python tools/profile_step.py > gpurun_out/r02e_ncu_conv0.log 2>&1; echo "ncu conv0 rc=$?"
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gn_gelu_fwd -c 1 -o gpurun_out/r02e_gnfwd -f python tools/profile_step.py > gpurun_out/r02e_ncu_gnfwd.log 2>&1; echo "ncu gnfwd rc=$?"
ls -la gpurun_out/*.ncu-rep
