# Command list behind the profiles/r02f_* evidence (each ncu pass only after the same command ran plain and exited 0).
cd $GRAFT_REPO_ROOT
timeout -k 5 600 python -m pytest tests -q -m gpu > gpurun_out/r02f_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02f_pytest_gpu.log | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02f_smoke.log | cut -c1-250
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02f_bench_w2v_base_15s_n1.json 2> gpurun_out/r02f_bench_w2v.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02f_bench_reference_arm.json 2> gpurun_out/r02f_bench_ref.err; echo "bench ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02f_bench_w2v_base_15s_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','step_frac_of_bf16_sustained','gpu_launches','simt_downgrades')}, d['e2e']['value'], d['cpu_baseline'])
for k in d['kernel_rooflines']: print(f"{k['frac']:.3f} {k['us']:8.1f} us  {k.get('us_warm_l2',0):6.1f} warm  {k['kernel'][:110]}")
for w in d.get('extra',{}).get('workloads',[]): print(w.get('config',{}).get('workload'), w.get('value'), w.get('ms_per_step'))
r=json.loads(open('gpurun_out/r02f_bench_reference_arm.json').read().strip().splitlines()[-1]); print('ref', r.get('value'), r.get('cpu_baseline'))
PY
timeout 200 python tools/profile_step.py > /dev/null 2>&1 && \
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02f_launches_w2v_base_15s_b8.csv python tools/profile_step.py > gpurun_out/r02f_ncu_launch_w2v.log 2>&1; echo "ncu list w2v rc=$?"
python tools/launch_summary.py gpurun_out/r02f_launches_w2v_base_15s_b8.csv > gpurun_out/r02f_launch_summary_w2v.txt 2>&1; head -30 gpurun_out/r02f_launch_summary_w2v.txt
timeout 200 python tools/profile_step.py --family whisper > /dev/null 2>&1 && \
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02f_launches_whisper_small_30s_b4.csv python tools/profile_step.py --family whisper > gpurun_out/r02f_ncu_launch_whisper.log 2>&1; echo "ncu list whisper rc=$?"
python tools/launch_summary.py gpurun_out/r02f_launches_whisper_small_30s_b4.csv > gpurun_out/r02f_launch_summary_whisper.txt 2>&1; head -12 gpurun_out/r02f_launch_summary_whisper.txt
timeout 100 ./tools/selftest_gemm prof > /dev/null 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 1 -o gpurun_out/r02f_gemm_ffn1 -f ./tools/selftest_gemm prof > gpurun_out/r02f_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 100 ./tools/selftest_attn prof > /dev/null 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"attn_fwd2|attn_bwd2" -c 2 -o gpurun_out/r02f_attn -f ./tools/selftest_attn prof > gpurun_out/r02f_ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"gn_gelu_bwd1_ring" -s 6 -c 1 -o gpurun_out/r02f_gnbwd1 -f python tools/profile_step.py > gpurun_out/r02f_ncu_gnbwd1.log 2>&1; echo "ncu gnbwd1 rc=$?"
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"ew_colsum_ring|ln_bwd_ring|conv0_wgrad" -c 3 -o gpurun_out/r02f_ew_ln -f python tools/profile_step.py > gpurun_out/r02f_ncu_ew_ln.log 2>&1; echo "ncu ew/ln rc=$?"
ls -la gpurun_out/*.ncu-rep | tail
