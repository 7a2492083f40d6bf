cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for v in "" 1; do
TETHYS_PACK_ON_MAIN=$v timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 3 --workload whisper_small_30s --no-extra > gpurun_out/n2d_whisper_$v.json 2> gpurun_out/n2d_whisper_$v.err; echo "whisper pack_on_main=$v rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/n2d_whisper_$v.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
PY
done
timeout 300 $TR tools/check_dist_graph.py > gpurun_out/r02_check_dist_graph_n2.log 2>&1; echo "check_dist_graph rc=$?"; grep -E "PASSED|FAILED|Error|error" gpurun_out/r02_check_dist_graph_n2.log | head -8 | cut -c1-300
