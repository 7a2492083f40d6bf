cd $GRAFT_REPO_ROOT
timeout -k 5 600 python -m pytest tests -q -m gpu > gpurun_out/r02f_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02f_pytest_gpu.log | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02f_smoke.log | cut -c1-250
for wl in whisper_small_30s whisper_base_30s; do
timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02f_bench_${wl}_n1.json 2> gpurun_out/last_$wl.err; echo "bench $wl rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02f_bench_${wl}_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])
for k in d['kernel_rooflines']:
    if 'attn' in k['kernel']: print(f"{k['frac']:.3f} {k['us']:8.1f} us  {k['kernel'][:100]}")
PY
done
