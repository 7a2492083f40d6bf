cd $GRAFT_REPO_ROOT
timeout -k 5 600 python -m pytest tests/test_w2v_gpu.py tests/test_whisper_gpu.py tests/test_norm_ops_gpu.py tests/test_w2v_heads_gpu.py tests/test_layers_gpu.py -q -m gpu -x > gpurun_out/ln_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/ln_pytest.log | cut -c1-300
for v in 0 1; do
for wl in w2v_base_15s whisper_small_30s; do
TETHYS_LN_BWD_DIRECT=$v timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ln${v}_$wl.json 2> gpurun_out/ln${v}_$wl.err; echo "bench direct=$v $wl rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/ln${v}_$wl.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])
PY
done
done
