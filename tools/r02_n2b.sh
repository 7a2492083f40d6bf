cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/check_dist_graph.py > gpurun_out/r02_check_dist_graph_n2.log 2>&1; echo "check_dist_graph rc=$?"; grep -E "PASSED|FAILED|Error|error" gpurun_out/r02_check_dist_graph_n2.log | head -8 | cut -c1-300
timeout 400 $TR tools/check_dist_oracle.py > gpurun_out/r02_check_dist_oracle_n2.log 2>&1; echo "check_dist_oracle rc=$?"; grep -E "^\[rank 0\]|Error|error|Traceback" gpurun_out/r02_check_dist_oracle_n2.log | head -24 | cut -c1-400
timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-extra > gpurun_out/r02_bench_n2_native.json 2> gpurun_out/r02_bench_n2_native.err; echo "bench native rc=$?"
timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 --workload whisper_small_30s --no-extra > gpurun_out/r02_bench_n2_whisper_native.json 2> gpurun_out/r02_bench_n2_whisper_native.err; echo "whisper native rc=$?"
python - <<'PY'
import json
for f in ("r02_bench_n2_native","r02_bench_n2_whisper_native"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["allreduce"][:160])
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
