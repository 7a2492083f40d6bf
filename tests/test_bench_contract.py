"""CPU-side checks of bench.py's measurement contract: the reference arm (`--impl reference`, the oracle's CPU restatement timed on the
host cores) prints one JSON line with the keys the driver reads, its `config` is the very object our arm prints for the same command
line, and under torchrun only rank 0 works. (Our arm needs a B200 and is exercised by the driver / profiles/.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_workload_config_names_the_workload_for_both_arms():
    sys.path.insert(0, ROOT)
    import bench

    for wl, (family, size, n_samples, secs, gflop) in bench.WORKLOADS.items():
        for world in (1, 2, 8):
            c = bench.workload_config(wl, 0, world)
            assert c["workload"] == wl and c["audio_seconds"] == secs and c["parallelism"] == f"dp{world}"
            assert c["per_gpu_batch"] == bench.DEFAULT_BATCH[family] and c["global_batch"] == world * c["per_gpu_batch"]
            assert ("none" in c["allreduce"]) == (world == 1) and "L2" in c["l2"]
            assert not any(k in c for k in ("cuda_graph", "collectives"))        # nothing about HOW an arm runs it
    assert bench.workload_config("w2v_base_15s", 4, 2)["global_batch"] == 8


def test_reference_arm_line_and_rank_gating():
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "w2v_tiny_2s", "--steps", "2",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train_samples_per_sec" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert "batch 1" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    sys.path.insert(0, ROOT)
    import bench

    assert d["config"] == bench.workload_config("w2v_tiny_2s", 0, 1)
    # under torchrun every rank but 0 exits 0 without work or output
    env2 = dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r2 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "w2v_tiny_2s", "--gpus", "2"],
                        capture_output=True, text=True, timeout=120, env=env2, cwd=ROOT)
    assert r2.returncode == 0 and not [ln for ln in r2.stdout.splitlines() if ln.startswith("{")]
