cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 500 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02f_bench_w2v_base_15s_n2.json 2> gpurun_out/r02f_bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/r02f_bench_n2.err | cut -c1-300
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02f_bench_w2v_base_15s_n2.json") if l.startswith("{")][-1])
print("main", d["value"], d["ms_per_step"], d["e2e"])
for w in d.get("extra",{}).get("workloads",[]):
    print(w.get("config",{}).get("workload"), w.get("value"), w.get("ms_per_step"), (w.get("e2e") or {}).get("value"), w.get("error"))
PY
timeout 200 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02f_bench_ref_n2.json 2> gpurun_out/r02f_bench_ref_n2.err; echo "ref n2 rc=$?"; tail -c 400 gpurun_out/r02f_bench_ref_n2.json
