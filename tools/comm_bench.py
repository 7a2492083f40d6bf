#!/usr/bin/env python
"""Under torchrun: where does the distributed Wav2Vec2 step's time go? (1) the bare all-reduce of the gradient arena through
ts_comm (registered buffer) and through torch.distributed, (2) the graphed distributed step, (3) the same step with the collective
skipped (TETHYS_SKIP_AR=1: numerically wrong, timing only). Prints per-rank numbers from rank 0 (max over ranks)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tethys_speech_b200 import wav2vec2 as W
from tethys_speech_b200.runtime import Adam, Strategy

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
st = Strategy()
rank, world = st.rank, st.world


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(); st.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda", dtype=torch.float64)
    st.dist.all_reduce(t, op=st.dist.ReduceOp.MAX)
    return float(t)


n = 92297728
for dt, name in ((torch.bfloat16, "bf16"), (torch.float32, "fp32")):
    if st.comm is not None:
        buf = st.alloc(n, dt)
        ms = timed(lambda: st._native_all_reduce(buf))
        if rank == 0:
            print(f"ts_comm all-reduce {name} {n * buf.element_size() / 1e6:.0f} MB (registered): {ms:.3f} ms  algbw {n * buf.element_size() / ms / 1e6:.0f} GB/s", flush=True)
    tb = torch.zeros(n, dtype=dt, device="cuda")
    ms = timed(lambda: st.dist.all_reduce(tb))
    if rank == 0:
        print(f"torch.distributed all-reduce {name}: {ms:.3f} ms  algbw {n * tb.element_size() / ms / 1e6:.0f} GB/s", flush=True)
    del tb

rng = np.random.default_rng(1234 + rank)
x = torch.from_numpy(rng.standard_normal((8, 240000), dtype=np.float32)).cuda()
for skip in ("0", "1"):
    os.environ["TETHYS_SKIP_AR"] = skip
    model = W.Wav2Vec2ForPreTraining(W.Wav2Vec2Config("base"), precision="bf16", device=local, seed=0)
    model.broadcast_weights(st)
    opt = Adam(learning_rate=3e-5, epsilon=1e-8, clipnorm=1.0)
    gstep, segs = W.make_graphed_distributed_step(st, model, opt, x)
    ms = timed(lambda: gstep(x), iters=10)
    if rank == 0:
        print(f"graphed distributed step (collective {'SKIPPED' if skip == '1' else 'on'}): {ms:.3f} ms", flush=True)
    del gstep, segs, model, opt
    torch.cuda.empty_cache()
st.barrier()
