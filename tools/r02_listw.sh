cd $GRAFT_REPO_ROOT
timeout 200 python tools/profile_step.py --family whisper > gpurun_out/r02d_profile_step_whisper.log 2>&1 && \
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02d_launches_whisper.csv python tools/profile_step.py --family whisper > gpurun_out/r02d_ncu_launch_whisper.log 2>&1; echo "ncu list rc=$?"
python tools/launch_summary.py gpurun_out/r02d_launches_whisper.csv > gpurun_out/r02d_launch_summary_whisper.txt 2>&1; head -48 gpurun_out/r02d_launch_summary_whisper.txt
