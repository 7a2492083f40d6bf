#!/usr/bin/env python
"""N-GPU check (run under torchrun): the graphed distributed steps (CUDA graphs around eager NCCL all-reduces) must produce the same parameters as the plain eager distributed_train_step after the same steps (dropout off)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tethys_speech_b200 import wav2vec2 as W2V
from tethys_speech_b200 import whisper as WH
from tethys_speech_b200.runtime import Adam, Strategy

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
strategy = Strategy()
rank = strategy.rank
rng = np.random.default_rng(100 + rank)
ok = True

# ---- Wav2Vec2 (tiny preset, fp32 so that the comparison is tight) ----
ma = W2V.Wav2Vec2ForPreTraining(W2V.Wav2Vec2Config("tiny"), precision="fp32", device=local, seed=0)
mb = W2V.Wav2Vec2ForPreTraining(W2V.Wav2Vec2Config("tiny"), precision="fp32", device=local, seed=0)
ma.broadcast_weights(strategy); mb.set_weights(ma.get_weights())
oa = Adam(learning_rate=1e-4, epsilon=1e-8, clipnorm=1.0); ob = Adam(learning_rate=1e-4, epsilon=1e-8, clipnorm=1.0)
x = torch.from_numpy(rng.standard_normal((2, 6400), dtype=np.float32)).cuda()
T = ma.num_frames(6400)
neg = ma._sample_negative_indices(T, 2)[:, 0, :].contiguous()
p0 = ma._prog.params.clone()
gstep, segs = W2V.make_graphed_distributed_step(strategy, mb, ob, x, dropout=False, warmup=1)
mb._sample_negative_indices = lambda T_, B_: neg.unsqueeze(1)          # same negatives on both paths
for _ in range(2):
    lb = gstep(x)
for _ in range(3):                                                         # 1 warm-up + 2 graphed steps on model b
    la = W2V.distributed_train_step(strategy, ma, (x, None), oa, neg_indices=neg, dropout=False)
torch.cuda.synchronize()
upd = float((ma._prog.params - p0).norm())
err = float((ma._prog.params - mb._prog.params).norm()) / upd
print(f"[rank {rank}] w2v: loss eager {float(la):.5f} graphed {float(lb):.5f}  param diff / update = {err:.2e}  buckets={len(segs.items)}")
ok = ok and err < 1e-2 and oa.iterations == ob.iterations

# ---- Whisper (tiny shapes via config edits, fp32) ----
def small():
    c = WH.WhisperConfig()
    c.d_model, c.d_ff, c.encoder_layers, c.decoder_layers = 128, 256, 2, 2
    c.encoder_attention_heads = c.decoder_attention_heads = 2
    c.vocab_size, c.n_ctx, c.decoder_start_token_id = 512, 64, 500
    return c
wa = WH.WhisperForConditionalGeneration(small(), precision="fp32", device=local, seed=0)
wb = WH.WhisperForConditionalGeneration(small(), precision="fp32", device=local, seed=0)
wa.broadcast_weights(strategy); wb.set_weights(wa.get_weights())
oa = Adam(learning_rate=1e-4); ob = Adam(learning_rate=1e-4)
f = torch.from_numpy(rng.standard_normal((2, 80, 128), dtype=np.float32)).cuda()
lab = torch.from_numpy(rng.integers(0, 100, size=(2, 20)).astype(np.int32)).cuda()
p0 = wa._prog.params.clone()
gstep, segs = WH.make_graphed_distributed_step(strategy, wb, ob, f, lab, dropout=False, warmup=1, bucket_elems=1 << 16)
for _ in range(2):
    lb = gstep(f, lab)
for _ in range(3):
    la = WH.distributed_train_step(strategy, wa, (f, lab), oa, dropout=False)
torch.cuda.synchronize()
upd = float((wa._prog.params - p0).norm())
err = float((wa._prog.params - wb._prog.params).norm()) / upd
print(f"[rank {rank}] whisper: loss eager {float(la):.5f} graphed {float(lb):.5f}  param diff / update = {err:.2e}  items={len(segs.items)}")
ok = ok and err < 1e-2 and oa.iterations == ob.iterations
# ---- bf16 compute: bf16 gradient buckets (default) against fp32 buckets (TETHYS_AR_DTYPE=fp32) -----------------------------
# same weights, same data, dropout off, 3 steps each: the two runs may differ by the bf16 rounding of the summed gradient only
def _w2v_bf16_run(ar_dtype):
    os.environ["TETHYS_AR_DTYPE"] = ar_dtype
    m = W2V.Wav2Vec2ForPreTraining(W2V.Wav2Vec2Config("tiny"), precision="bf16", device=local, seed=0)
    m.broadcast_weights(strategy)
    o = Adam(learning_rate=1e-4, epsilon=1e-8, clipnorm=1.0)
    assert m._prog.ar_bf16() == (ar_dtype == "bf16")
    step_fn, _ = W2V.make_graphed_distributed_step(strategy, m, o, x, dropout=False, warmup=1)
    m._rng.manual_seed(77)
    losses = [float(step_fn(x)) for _ in range(3)]
    eager = float(W2V.distributed_train_step(strategy, m, (x, None), o, neg_indices=neg, dropout=False))   # eager path, same bucket dtype
    return losses, eager


l16, e16 = _w2v_bf16_run("bf16")
l32, e32 = _w2v_bf16_run("fp32")
os.environ.pop("TETHYS_AR_DTYPE", None)
bad = [abs(a - b) > 2e-2 * abs(b) for a, b in zip(l16 + [e16], l32 + [e32])]
print(f"[rank {rank}] w2v bf16: losses with bf16 buckets {l16} / fp32 buckets {l32}; eager {e16:.5f} / {e32:.5f}", flush=True)
print(f"[rank {rank}] {'BF16-BUCKET CHECK FAILED' if any(bad) else 'BF16-BUCKET CHECK PASSED'}", flush=True)
ok = ok and not any(bad)


def _whisper_bf16_run(ar_dtype):
    os.environ["TETHYS_AR_DTYPE"] = ar_dtype
    m = WH.WhisperForConditionalGeneration(small(), precision="bf16", device=local, seed=0)
    m.broadcast_weights(strategy)
    o = Adam(learning_rate=1e-4)
    step_fn, _ = WH.make_graphed_distributed_step(strategy, m, o, f, lab, dropout=False, warmup=1, bucket_elems=1 << 16)
    losses = [float(step_fn(f, lab)) for _ in range(3)]
    eager = float(WH.distributed_train_step(strategy, m, (f, lab), o, dropout=False))    # backward_allreduce_overlapped path
    return losses, eager


w16, we16 = _whisper_bf16_run("bf16")
w32, we32 = _whisper_bf16_run("fp32")
os.environ.pop("TETHYS_AR_DTYPE", None)
bad = [abs(a - b) > 2e-2 * abs(b) for a, b in zip(w16 + [we16], w32 + [we32])]
print(f"[rank {rank}] whisper bf16: losses with bf16 buckets {w16} / fp32 buckets {w32}; eager {we16:.5f} / {we32:.5f}", flush=True)
print(f"[rank {rank}] {'WHISPER BF16-BUCKET CHECK FAILED' if any(bad) else 'WHISPER BF16-BUCKET CHECK PASSED'}", flush=True)
ok = ok and not any(bad)

strategy.dist.barrier()
strategy.dist.destroy_process_group()
print(f"[rank {rank}] {'DIST GRAPH CHECK PASSED' if ok else 'DIST GRAPH CHECK FAILED'}")
sys.exit(0 if ok else 1)

