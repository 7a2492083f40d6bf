cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 500 $TR bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02e_bench_n8.json 2> gpurun_out/r02e_bench_n8.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02e_bench_n8.json") if l.startswith("{")][-1])
print("main", d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["allreduce"][:120])
for w in d.get("extra",{}).get("workloads",[]):
    print(w.get("config",{}).get("workload"), w.get("value"), w.get("ms_per_step"), (w.get("e2e") or {}).get("value"), w.get("error"))
PY
timeout 300 $TR tools/comm_bench.py > gpurun_out/r02e_comm_bench_n8.log 2>&1; echo "comm bench rc=$?"; tail -8 gpurun_out/r02e_comm_bench_n8.log
