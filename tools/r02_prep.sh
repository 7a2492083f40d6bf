cd $GRAFT_REPO_ROOT
timeout 100 ./tools/selftest_attn > gpurun_out/prep_selftest_attn.log 2>&1; echo "selftest attn rc=$?"; tail -7 gpurun_out/prep_selftest_attn.log | cut -c1-170
timeout -k 5 600 python -m pytest tests/test_attention_gpu.py tests/test_gemm_gpu.py tests/test_w2v_gpu.py tests/test_whisper_gpu.py -q -m gpu -x > gpurun_out/prep_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/prep_pytest.log | cut -c1-200
for wl in w2v_base_15s whisper_small_30s; do
timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/prep_$wl.json 2> gpurun_out/prep_$wl.err; echo "bench $wl rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/prep_$wl.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])
for k in d['kernel_rooflines'][6:8]: print(f"{k['frac']:.3f} {k['us']:8.1f} us  {k['kernel'][:100]}")
PY
done
