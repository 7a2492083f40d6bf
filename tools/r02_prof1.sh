cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_w2v_e.json 2> gpurun_out/r02_bench_w2v_e.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_w2v_e.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','step_frac_of_bf16_sustained')}, d['e2e']['value'], d.get('cpu_baseline'))
for e in d['extra']['workloads']: print(e.get('config',{}).get('workload'), e.get('ms_per_step'), e.get('value'), e.get('step_frac_of_bf16_sustained'), e.get('error'))
PY
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_w2v_base_15s_b8.csv python tools/profile_step.py > gpurun_out/r02_prof_w2v.log 2>&1; echo "ncu w2v rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_whisper_small_30s_b4.csv python tools/profile_step.py --family whisper > gpurun_out/r02_prof_whisper.log 2>&1; echo "ncu whisper rc=$?"
python tools/launch_summary.py gpurun_out/r02_launches_w2v_base_15s_b8.csv > gpurun_out/r02_launch_summary_w2v.txt; head -48 gpurun_out/r02_launch_summary_w2v.txt
python tools/launch_summary.py gpurun_out/r02_launches_whisper_small_30s_b4.csv > gpurun_out/r02_launch_summary_whisper.txt; head -36 gpurun_out/r02_launch_summary_whisper.txt
