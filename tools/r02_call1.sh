set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -2; nproc
timeout 900 python tools/parity_report.py --out gpurun_out/r02_parity_report_before.json > gpurun_out/r02_parity_before.log 2>&1; echo "parity rc=$?"
tail -12 gpurun_out/r02_parity_before.log
timeout 200 python tools/profile_step.py --family whisper > gpurun_out/r02_profile_whisper.log 2>&1 && \
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_whisper_small_30s_b4.csv python tools/profile_step.py --family whisper > gpurun_out/r02_ncu_whisper.log 2>&1; echo "ncu list rc=$?"
cat gpurun_out/r02_profile_whisper.log | tail -2
