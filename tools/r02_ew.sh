cd $GRAFT_REPO_ROOT
timeout -k 5 600 python -m pytest tests/test_w2v_gpu.py tests/test_whisper_gpu.py tests/test_dropout_gpu.py tests/test_golden_gpu.py -q -m gpu -x > gpurun_out/ew_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/ew_pytest.log | cut -c1-300
for v in 0 1; do
TETHYS_EW_DIRECT=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ew${v}.json 2> gpurun_out/ew${v}.err; echo "bench direct=$v rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/ew${v}.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])
PY
done
