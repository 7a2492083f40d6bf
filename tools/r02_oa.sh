cd $GRAFT_REPO_ROOT
timeout -k 5 600 python -m pytest tests/test_whisper_gpu.py tests/test_graph_gpu.py tests/test_golden_gpu.py tests/test_ref_golden_gpu.py tests/test_checkpoint.py tests/test_cli.py -q -m gpu -x > gpurun_out/oa_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/oa_pytest.log | cut -c1-200
for v in 1 0; do
for wl in whisper_small_30s whisper_base_30s; do
TETHYS_OVERLAP_ADAM=$v timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/oa${v}_$wl.json 2> gpurun_out/oa${v}_$wl.err; echo "bench overlap=$v $wl rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/oa${v}_$wl.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])
PY
done
done
