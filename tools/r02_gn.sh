cd $GRAFT_REPO_ROOT
timeout -k 5 600 python -m pytest tests/test_gemm_gpu.py tests/test_w2v_gpu.py tests/test_norm_ops_gpu.py tests/test_fullsize_properties_gpu.py tests/test_golden_gpu.py tests/test_ref_golden_gpu.py -q -m gpu -x > gpurun_out/gn_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/gn_pytest.log | cut -c1-300
for v in 0 1; do
TETHYS_NO_FUSED_GN=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/gn${v}.json 2> gpurun_out/gn${v}.err; echo "bench nofuse=$v rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/gn${v}.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])
PY
done
