"""World-size-2 CPU (gloo) tests of the multi-replica host logic: the Strategy shim that stands in for
MultiWorkerMirroredStrategy (bucketed SUM all-reduce of the flat gradient arena = K21, weight broadcast = K23, scalar loss
reduce = K22) and the two reduce conventions of the reference: Whisper sums un-normalised gradients
(W:829-836), Wav2Vec2 divides the loss by N, clips locally BEFORE the reduce and per-variable AFTER it (V:1231-1246).
The GPU kernels cannot run here; gradients come from the CPU oracle and travel through the same Strategy code."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from oracle import tf_ops as T
    from oracle import wav2vec2_oracle as O
    from tethys_speech_b200.runtime import ReduceOp, Strategy

    torch.set_num_threads(1)
    st = Strategy(backend="gloo")
    assert st.num_replicas_in_sync == world
    res = {}
    # K23 broadcast: replicas start from different "initial" arenas, the chief's wins
    arena = torch.full((1000,), float(rank + 1))
    st.broadcast_(arena)
    res["bcast"] = float(arena.sum())
    # K21 bucketed all-reduce over a flat arena with a ragged last bucket
    g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    st.all_reduce_sum_(g, bucket_elems=300)
    res["allreduce_ok"] = bool(torch.equal(g, torch.arange(1000, dtype=torch.float32) * 3))
    # K22 loss reduce
    res["loss_sum"] = float(st.reduce(ReduceOp.SUM, torch.tensor(0.5 + rank)))
    # asynchronous bucket reduce with a packing hook: `pre` must have run before the collective reads the bucket
    bucket = torch.zeros(64)
    st.all_reduce_async_(bucket, pre=lambda: bucket.fill_(float(rank + 1)))
    st.join_async()
    res["async_pre"] = float(bucket[0])
    # Wav2Vec2 convention end to end on the oracle's gradients
    cfg = O.Wav2Vec2Config("tiny")
    w = O.randomize_weights(O.init_weights(cfg, seed=0, dtype=torch.float64), seed=1)
    gen = torch.Generator().manual_seed(100 + rank)
    wave = torch.randn(1, 1600, generator=gen, dtype=torch.float64)
    Tn = O.num_frames(cfg, 1600)
    neg = O.negative_indices_from_random(torch.randint(0, Tn, (1, Tn), generator=gen), cfg.num_negatives)
    out, grads = O.loss_and_grads(cfg, w, wave, neg, loss_div=float(world))
    names = list(w)
    clipped, _ = T.clip_by_global_norm([grads[k] for k in names], 1.0)                 # local, pre-reduce (V:1243)
    flat = torch.cat([c.reshape(-1) for c in clipped])
    st.all_reduce_sum_(flat, bucket_elems=1 << 20)                                       # SUM, not mean (A-13)
    res["flat"] = flat.numpy().copy()                                                    # by value: the worker exits before the parent reads
    res["local"] = torch.cat([c.reshape(-1) for c in clipped]).numpy().copy()
    res["scaled_loss"] = float(st.reduce(ReduceOp.SUM, (out["loss"] / world).detach()))
    res["loss"] = float(out["loss"])
    q.put((rank, res))
    st.barrier()


@pytest.mark.timeout(300)
def test_strategy_shim_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        assert got[r]["bcast"] == 1000.0            # everyone holds the chief's (rank 0) values
        assert got[r]["allreduce_ok"]
        assert got[r]["loss_sum"] == 0.5 + 1.5
        assert got[r]["async_pre"] == 3.0           # 1 + 2: both replicas' buckets were packed before the SUM
    # the reduced gradient is the SUM of the two locally clipped gradients, identical on both ranks
    import numpy as np

    assert np.array_equal(got[0]["flat"], got[1]["flat"])
    assert np.allclose(got[0]["flat"], got[0]["local"] + got[1]["local"], rtol=0, atol=1e-15)
    # each local part has global norm <= 1 (clip before reduce), the sum may exceed it (hence the post-reduce clipnorm)
    assert np.linalg.norm(got[0]["local"]) <= 1.0 + 1e-12 and np.linalg.norm(got[1]["local"]) <= 1.0 + 1e-12
    # the reported loss is the SUM over replicas of loss/N = the mean of the per-replica losses (V:1231, V:1260)
    mean_loss = 0.5 * (got[0]["loss"] + got[1]["loss"])
    assert abs(got[0]["scaled_loss"] - mean_loss) < 1e-6 * mean_loss       # the scalar travels as float32, like TF's loss


# ---- the reference's own launch: no torchrun, every pod runs `python speech_jobs/*_dist.py` with the TFJob operator's TF_CONFIG ----
def _tf_config_worker(task_type, task_index, cluster, q):
    import json

    sys.path.insert(0, ROOT)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        os.environ.pop(k, None)
    os.environ["TF_CONFIG"] = json.dumps({"cluster": cluster, "task": {"type": task_type, "index": task_index}, "environment": "cloud"})
    from tethys_speech_b200.runtime import Strategy

    torch.set_num_threads(1)
    st = Strategy(backend="gloo")
    g = torch.full((8,), float(st.rank + 1))
    st.all_reduce_sum_(g)
    arena = torch.full((4,), float(10 * (st.rank + 1)))
    st.broadcast_(arena)
    q.put((task_type, task_index, st.rank, st.num_replicas_in_sync, float(g[0]), float(arena[0])))
    st.barrier()


@pytest.mark.timeout(300)
def test_strategy_rendezvous_from_tf_config_world3_gloo():
    """W:1037-1047 / sample_tfjobs/*.yaml: CHIEF + 2 WORKER pods, TF_CONFIG only. Replica order = chief, worker 0, worker 1; the chief's
    TFJob port is the rendezvous address."""
    from tethys_speech_b200.runtime import rendezvous_from_tf_config

    port = _free_port()
    cluster = {"chief": [f"127.0.0.1:{port}"], "worker": ["127.0.0.1:2223", "127.0.0.1:2224"]}
    assert rendezvous_from_tf_config({"cluster": cluster, "task": {"type": "worker", "index": 1}}) == (3, 2, "127.0.0.1", port)
    assert rendezvous_from_tf_config({"cluster": {"CHIEF": cluster["chief"], "WORKER": cluster["worker"], "ps": ["x:1"]},
                                      "task": {"type": "chief", "index": 0}}) == (3, 0, "127.0.0.1", port)
    assert rendezvous_from_tf_config({}) is None and rendezvous_from_tf_config("") is None
    assert rendezvous_from_tf_config({"cluster": {"worker": ["h:1"]}, "task": {"type": "worker", "index": 0}}) is None     # one task
    with pytest.raises(ValueError):
        rendezvous_from_tf_config({"cluster": cluster, "task": {"type": "ps", "index": 0}})
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    tasks = [("chief", 0), ("worker", 0), ("worker", 1)]
    procs = [ctx.Process(target=_tf_config_worker, args=(t, i, cluster, q)) for t, i in tasks]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=240) for _ in tasks)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(t, i, r) for t, i, r, *_ in got] == [("chief", 0, 0), ("worker", 0, 1), ("worker", 1, 2)]
    for _, _, _, world, red, bc in got:
        assert world == 3 and red == 6.0 and bc == 10.0          # 1 + 2 + 3; everyone holds the chief's arena
