cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=TUNING,REG timeout 500 $TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_n8.out 2> gpurun_out/r02_bench_n8.err; echo "bench n8 rc=$?"
grep -E "^\{" gpurun_out/r02_bench_n8.out > gpurun_out/r02_bench_w2v_base_15s_n8.json
grep -E "AllReduce: [0-9]{7,} Bytes" gpurun_out/r02_bench_n8.out | sed 's/.*NCCL INFO //' | sort | uniq -c | sort -rn | head -8
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_w2v_base_15s_n8.json').read().strip().splitlines()[-1])
    print('main', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['allreduce'][:150])
    for e in d['extra']['workloads']: print(e.get('config',{}).get('workload'), e.get('ms_per_step'), e.get('value'), e.get('error'))
except Exception as ex:
    print('ERR', ex); print(open('gpurun_out/r02_bench_n8.err').read()[-2500:])
PY
