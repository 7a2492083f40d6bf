cd $GRAFT_REPO_ROOT
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b_launches_w2v_base_15s_b8.csv python tools/profile_step.py > gpurun_out/r02_prof_w2v.log 2>&1; echo "ncu w2v rc=$?"
python tools/launch_summary.py gpurun_out/r02b_launches_w2v_base_15s_b8.csv > gpurun_out/r02b_launch_summary_w2v.txt; head -50 gpurun_out/r02b_launch_summary_w2v.txt
