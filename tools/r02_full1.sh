cd $GRAFT_REPO_ROOT
timeout -k 5 1200 python -m pytest tests -q -m gpu -x > gpurun_out/r02_pytest_gpu_d.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_pytest_gpu_d.log | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_w2v_d.json 2> gpurun_out/r02_bench_w2v_d.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_w2v_d.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','step_frac_of_bf16_sustained')}, d['e2e']['value'])
for k in d['kernel_rooflines']: print(f"{k['frac']:.3f} {k['us']:8.1f} us  {k['kernel'][:100]}")
PY
