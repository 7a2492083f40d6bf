cd $GRAFT_REPO_ROOT
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02f_bench_w2v_base_15s_n1.json 2> gpurun_out/r02f_bench_w2v.err; echo "bench rc=$?"; tail -3 gpurun_out/r02f_bench_w2v.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02f_bench_w2v_base_15s_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','step_frac_of_bf16_sustained','gpu_launches','simt_downgrades')}, d['e2e'], d['cpu_baseline'])
for k in d['kernel_rooflines']: print(f"{k['frac']:.3f} {k['us']:8.1f} us  {k.get('us_warm_l2',0):6.1f} warm  {k['kernel'][:110]}")
for w in d.get('extra',{}).get('workloads',[]): print(w.get('config',{}).get('workload'), w.get('value'), w.get('ms_per_step'), w.get('e2e',{}).get('value'))
PY
