cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_grad_buckets_gpu.py tests/test_abi.py -q -x 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/check_dist_graph.py > gpurun_out/r02_check_dist_graph_n2.log 2>&1; echo "check_dist_graph rc=$?"; grep -E "PASSED|FAILED|Error|error" gpurun_out/r02_check_dist_graph_n2.log | head -8 | cut -c1-300
timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-extra > gpurun_out/r02_bench_n2_native.json 2> gpurun_out/r02_bench_n2_native.err; echo "bench native rc=$?"
timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 --workload whisper_small_30s --no-extra > gpurun_out/r02_bench_n2_whisper_native.json 2> gpurun_out/r02_bench_n2_whisper_native.err; echo "whisper native rc=$?"
timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 --workload whisper_base_30s --no-extra > gpurun_out/r02_bench_n2_whisper_base.json 2> gpurun_out/r02_bench_n2_whisper_base.err; echo "whisper base rc=$?"
timeout 400 python bench.py --steps 10 --warmup 3 --workload whisper_base_30s --no-extra --no-cpu-baseline > gpurun_out/r02_bench_n1_whisper_base.json 2> gpurun_out/r02_bench_n1_whisper_base.err; echo "whisper base n1 rc=$?"
python - <<'PY'
import json
for f in ("r02_bench_n2_native","r02_bench_n2_whisper_native","r02_bench_n2_whisper_base","r02_bench_n1_whisper_base"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["allreduce"][:100])
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
