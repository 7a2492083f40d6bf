cd $GRAFT_REPO_ROOT
for mode in 1 2; do
cp gpurun_scratch/lib_mode$mode.so tethys_speech_b200/libtethys.so
for wl in w2v_base_15s whisper_small_30s; do
timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/pdlm${mode}_$wl.json 2> gpurun_out/pdlm${mode}_$wl.err; echo "bench mode=$mode $wl rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/pdlm${mode}_$wl.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])
PY
done
done
