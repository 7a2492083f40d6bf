cd $GRAFT_REPO_ROOT
timeout -k 5 600 python -m pytest tests -q -m gpu -x > gpurun_out/pdl_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pdl_pytest.log | cut -c1-300
for pdl in 1 0; do
for wl in w2v_base_15s whisper_small_30s; do
TETHYS_PDL=$pdl timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/pdl${pdl}_$wl.json 2> gpurun_out/pdl${pdl}_$wl.err; echo "bench pdl=$pdl $wl rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/pdl${pdl}_$wl.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])
PY
done
done
