#!/usr/bin/env python
"""Run a few Wav2Vec2 train steps and bracket exactly one steady-state step with cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python tools/profile_step.py
Also usable without ncu (prints the step time)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from tethys_speech_b200 import wav2vec2 as W

ap = argparse.ArgumentParser()
ap.add_argument("--size", default="base")
ap.add_argument("--samples", type=int, default=240000)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--steps", type=int, default=1)
args = ap.parse_args()

torch.cuda.set_device(0)
model = W.Wav2Vec2ForPreTraining(W.Wav2Vec2Config(args.size), precision=args.precision, device=0)
opt = W.Adam(learning_rate=3e-5, epsilon=1e-8, clipnorm=1.0)
rng = np.random.default_rng(1234)
x = torch.from_numpy(rng.standard_normal((args.batch, args.samples), dtype=np.float32)).cuda()
for _ in range(args.warmup):
    W.train_step(model, (x, None), opt)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
for _ in range(args.steps):
    loss = W.train_step(model, (x, None), opt)
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
model._prog.ctx.watchdog()
print(f"step_ms={e0.elapsed_time(e1) / args.steps:.3f} loss={float(loss):.4f}")
