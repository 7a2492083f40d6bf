cd $GRAFT_REPO_ROOT
timeout -k 5 600 python -m pytest tests -q -m gpu -x > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02c_pytest_gpu.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02c_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02c_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02c_bench_w2v.json 2> gpurun_out/r02c_bench_w2v.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02c_bench_w2v.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','step_frac_of_bf16_sustained')}, d['e2e']['value'])
for k in d['kernel_rooflines']: print(f"{k['frac']:.3f} {k['us']:8.1f} us  {k['kernel'][:100]}")
PY
timeout 200 python tools/profile_step.py > gpurun_out/r02c_profile_step.log 2>&1 && \
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02c_launches_w2v.csv python tools/profile_step.py > gpurun_out/r02c_ncu_launch.log 2>&1; echo "ncu list rc=$?"
python tools/launch_summary.py gpurun_out/r02c_launches_w2v.csv > gpurun_out/r02c_launch_summary_w2v.txt 2>&1; head -50 gpurun_out/r02c_launch_summary_w2v.txt
timeout 100 ./tools/selftest_attn > gpurun_out/r02c_selftest_attn.log 2>&1; tail -8 gpurun_out/r02c_selftest_attn.log
