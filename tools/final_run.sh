set -x
cd $GRAFT_REPO_ROOT
timeout -k 5 420 python -m pytest tests -q -m gpu > gpurun_out/r01b_pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/r01b_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01b_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r01b_smoke.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r01b_bench_w2v.json 2> gpurun_out/r01b_bench_w2v.err; echo "bench rc=$?"
timeout 300 python bench.py --workload whisper_small_30s --steps 20 --warmup 3 > gpurun_out/r01b_bench_whisper.json 2> gpurun_out/r01b_bench_whisper.err; echo "bench whisper rc=$?"
timeout 200 python tools/profile_step.py > gpurun_out/r01b_profile_step.log 2>&1 && \
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01b_launches.csv python tools/profile_step.py > gpurun_out/r01b_ncu_launch.log 2>&1; echo "ncu list rc=$?"
timeout 100 ./tools/selftest_gemm prof > /dev/null 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 1 -o gpurun_out/r01b_gemm_ffn1 -f ./tools/selftest_gemm prof > gpurun_out/r01b_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -15
