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
ap.add_argument("--family", default="w2v", choices=["w2v", "whisper"])
ap.add_argument("--size", default=None, help="w2v: tiny/small/base/large (default base); whisper: tiny/base/small (default small)")
ap.add_argument("--samples", type=int, default=240000)
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--steps", type=int, default=1)
args = ap.parse_args()

torch.cuda.set_device(0)
rng = np.random.default_rng(1234)
if args.family == "w2v":
    B = args.batch or 8
    model = W.Wav2Vec2ForPreTraining(W.Wav2Vec2Config(args.size or "base"), precision=args.precision, device=0)
    opt = W.Adam(learning_rate=3e-5, epsilon=1e-8, clipnorm=1.0)
    x = torch.from_numpy(rng.standard_normal((B, args.samples), dtype=np.float32)).cuda()
    batch = (x, None)
else:
    from tethys_speech_b200 import whisper as W  # noqa: F811

    B = args.batch or 4
    model = W.create_whisper_model(args.size or "small", precision=args.precision, device=0)
    opt = W.Adam(learning_rate=1e-4)
    labels = np.zeros((B, 100), dtype=np.int32)
    for i in range(B):
        n = int(rng.integers(50, 90))
        labels[i, 0] = 1; labels[i, 1:n - 1] = rng.integers(3, 100, size=n - 2); labels[i, n - 1] = 2
    batch = (torch.from_numpy(rng.standard_normal((B, 80, 3000), dtype=np.float32)).cuda(), torch.from_numpy(labels).cuda())
for _ in range(args.warmup):
    W.train_step(model, batch, opt)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
for _ in range(args.steps):
    loss = W.train_step(model, batch, opt)
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
model._prog.ctx.watchdog()
print(f"step_ms={e0.elapsed_time(e1) / args.steps:.3f} loss={float(loss):.4f}")
