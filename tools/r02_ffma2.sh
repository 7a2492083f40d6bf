cd $GRAFT_REPO_ROOT
timeout -k 5 600 python -m pytest tests/test_w2v_gpu.py tests/test_golden_gpu.py tests/test_ref_golden_gpu.py tests/test_w2v_heads_gpu.py -q -m gpu -x > gpurun_out/ffma2_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/ffma2_pytest.log | cut -c1-200
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -k regex:"conv0|gn_gelu_bwd1_ring" --csv --log-file gpurun_out/ffma2_conv0.csv python tools/profile_step.py > /dev/null 2>&1; echo "ncu rc=$?"
grep -E "conv0|gn_gelu" gpurun_out/ffma2_conv0.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ffma2_bench.json 2> gpurun_out/ffma2_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/ffma2_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])
PY
