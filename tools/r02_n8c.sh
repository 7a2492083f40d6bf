cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 $TR bench.py --gpus 8 --steps 10 --warmup 3 --workload whisper_small_30s --no-extra > gpurun_out/n8c_$name.json 2> gpurun_out/n8c_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/n8c_$name.json") if l.startswith("{")][-1])
    print("$name", round(d["value"],1), round(d["ms_per_step"],3), round(d["e2e"]["value"],1))
except Exception as e:
    print("$name ERR", e); print(open("gpurun_out/n8c_$name.err").read()[-800:])
PY
}
run base A=0
run margin16 TETHYS_SM_MARGIN=16
run margin32 TETHYS_SM_MARGIN=32
run ctas8_margin8 NCCL_MAX_CTAS=8 TETHYS_SM_MARGIN=8
run ctas4_margin4 NCCL_MAX_CTAS=4 TETHYS_SM_MARGIN=4
