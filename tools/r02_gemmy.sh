cd $GRAFT_REPO_ROOT
timeout 300 ./tools/selftest_gemm > gpurun_out/r02e_selftest_gemm.log 2>&1; echo "selftest rc=$?"; grep -E "time|PASSED|FAILED" gpurun_out/r02e_selftest_gemm.log | grep -v "8192" | cut -c1-170
timeout -k 5 600 python -m pytest tests -q -m gpu -x > gpurun_out/r02e_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02e_pytest_gpu.log | cut -c1-200
for wl in w2v_base_15s whisper_small_30s whisper_base_30s; do
timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02e_$wl.json 2> gpurun_out/r02e_$wl.err; echo "bench $wl rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02e_$wl.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])
PY
done
