cd $GRAFT_REPO_ROOT
timeout -k 5 600 python -m pytest tests/test_whisper_gpu.py tests/test_graph_gpu.py tests/test_golden_gpu.py tests/test_ref_golden_gpu.py tests/test_whisper_generate_gpu.py tests/test_layers_gpu.py -q -m gpu -x > gpurun_out/kv_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/kv_pytest.log | cut -c1-200
for v in 0 1; do
for wl in whisper_small_30s whisper_base_30s; do
TETHYS_NO_SIDE_KV=$v timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/kv${v}_$wl.json 2> gpurun_out/kv${v}_$wl.err; echo "bench nosidekv=$v $wl rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/kv${v}_$wl.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])
PY
done
done
