cd $GRAFT_REPO_ROOT
timeout 300 ./tools/selftest_gemm > gpurun_out/w16_selftest.log 2>&1; echo "selftest rc=$?"
grep -E "FAIL|PASSED|FAILED|mismatch" gpurun_out/w16_selftest.log | head -20 | cut -c1-200
grep -E "time " gpurun_out/w16_selftest.log | grep -v "8192\|qk \|pv \|eng=1" | cut -c1-170
timeout -k 5 600 python -m pytest tests/test_gemm_gpu.py tests/test_dropout_gpu.py tests/test_w2v_gpu.py tests/test_whisper_gpu.py -q -m gpu -x > gpurun_out/w16_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/w16_pytest.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/w16_bench.json 2> gpurun_out/w16_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/w16_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])
for k in d['kernel_rooflines'][:6]: print(f"{k['frac']:.3f} {k['us']:8.1f} us  {k.get('us_warm_l2',0):6.1f} warm  {k['kernel'][:100]}")
PY
