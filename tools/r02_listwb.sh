cd $GRAFT_REPO_ROOT
timeout 200 python tools/profile_step.py --family whisper --size base > gpurun_out/r02d_profile_step_whisper_base.log 2>&1 && \
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02d_launches_whisper_base.csv python tools/profile_step.py --family whisper --size base > gpurun_out/r02d_ncu_launch_whisper_base.log 2>&1; echo "ncu list rc=$?"
python tools/launch_summary.py gpurun_out/r02d_launches_whisper_base.csv -g > gpurun_out/r02d_launch_summary_whisper_base.txt 2>&1; head -60 gpurun_out/r02d_launch_summary_whisper_base.txt
