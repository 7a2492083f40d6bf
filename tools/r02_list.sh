cd $GRAFT_REPO_ROOT
timeout 200 python tools/profile_step.py > gpurun_out/r02d_profile_step.log 2>&1 && \
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02d_launches_w2v.csv python tools/profile_step.py > gpurun_out/r02d_ncu_launch.log 2>&1; echo "ncu list rc=$?"
python tools/launch_summary.py gpurun_out/r02d_launches_w2v.csv > gpurun_out/r02d_launch_summary_w2v.txt 2>&1; head -45 gpurun_out/r02d_launch_summary_w2v.txt
python tools/launch_summary.py -g gpurun_out/r02d_launches_w2v.csv > gpurun_out/r02d_launch_summary_w2v_bygrid.txt 2>&1
