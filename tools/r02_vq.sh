cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_w2v_gpu.py tests/test_layers_gpu.py tests/test_golden_gpu.py tests/test_ref_golden_gpu.py tests/test_fullsize_properties_gpu.py -q -x > gpurun_out/r02_pytest_vq.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02_pytest_vq.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02_bench_w2v_f.json 2> gpurun_out/r02_bench_w2v_f.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_w2v_f.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','step_frac_of_bf16_sustained')}, d['e2e']['value'])
PY
