cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/check_dist_graph.py > gpurun_out/r02_check_dist_graph_n2.log 2>&1; echo "check_dist_graph rc=$?"; tail -6 gpurun_out/r02_check_dist_graph_n2.log | cut -c1-300
timeout 300 $TR tools/check_dist_oracle.py > gpurun_out/r02_check_dist_oracle_n2.log 2>&1; echo "check_dist_oracle rc=$?"; tail -8 gpurun_out/r02_check_dist_oracle_n2.log | cut -c1-300
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,REG,NVLS,TUNING timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-extra > gpurun_out/r02_bench_n2_native.json 2> gpurun_out/r02_bench_n2_native.err; echo "bench native rc=$?"
grep -iE "nvls|register|ncclMemAlloc" gpurun_out/r02_bench_n2_native.err | head -12 | cut -c1-250
TETHYS_NATIVE_COMM=0 timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-extra > gpurun_out/r02_bench_n2_torch.json 2> gpurun_out/r02_bench_n2_torch.err; echo "bench torch rc=$?"
timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 --workload whisper_small_30s --no-extra > gpurun_out/r02_bench_n2_whisper_native.json 2> gpurun_out/r02_bench_n2_whisper_native.err; echo "whisper native rc=$?"
TETHYS_NATIVE_COMM=0 timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 --workload whisper_small_30s --no-extra > gpurun_out/r02_bench_n2_whisper_torch.json 2> gpurun_out/r02_bench_n2_whisper_torch.err; echo "whisper torch rc=$?"
python - <<'PY'
import json
for f in ("r02_bench_n2_native","r02_bench_n2_torch","r02_bench_n2_whisper_native","r02_bench_n2_whisper_torch"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["allreduce"][:120])
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
