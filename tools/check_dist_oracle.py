#!/usr/bin/env python
"""N-GPU check over REAL NCCL (run under torchrun, N = 2): the CUDA distributed train steps — eager and the single-graph native
path (ts_comm collectives captured in the step's CUDA graph) — against the ORACLE's N-replica step
(oracle.train_step(peer_grads=...)) for both reduce conventions of the reference:
  Wav2Vec2  V:1186-1260  loss / N -> gradients -> LOCAL clip_by_global_norm(1.0) -> all-reduce SUM -> clipnorm(1.0) -> Adam
  Whisper   W:819-848    gradients of the local mean loss -> all-reduce SUM (not divided by N) -> Adam; loss = SUM over replicas
Every rank knows every rank's (seeded) data, so each rank evaluates the oracle for all replicas on its host cores and compares
its own post-all-reduce gradient arena, the reduced loss and the updated weights. fp32 compute: bars 1e-5 (gradients), 1e-4 (loss),
3e-3 (weight CHANGE after an Adam step: Adam's early updates are ~ lr * sign(g), which amplifies 1e-6 gradient differences
wherever |g| ~ 0; the error grows roughly linearly over the first steps — 1.4e-3, 1.9e-3, 2.2e-3 measured for steps 1-3). Uses oracle/ as the checker only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import tf_ops as T
from oracle import wav2vec2_oracle as OV
from oracle import whisper_oracle as OW
from tethys_speech_b200 import wav2vec2 as W2V
from tethys_speech_b200 import whisper as WH
from tethys_speech_b200.runtime import Adam, Strategy

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
strategy = Strategy()
rank, N = strategy.rank, strategy.world
ok = True


def rel(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-300))


def report(tag, worst_g, loss, want_loss, worst_u, gbar=1e-5, lbar=1e-4, ubar=3e-3):
    global ok
    good = worst_g[1] <= gbar and abs(loss - want_loss) <= lbar * abs(want_loss) and worst_u[1] <= ubar
    ok = ok and good
    print(f"[rank {rank}] {tag}: reduced-gradient worst {worst_g[0]} {worst_g[1]:.2e} (bar {gbar:.0e}); loss gpu {loss:.6f} oracle {want_loss:.6f}; "
          f"weight-change worst {worst_u[0]} {worst_u[1]:.2e} (bar {ubar:.0e})  {'ok' if good else 'FAIL'}", flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# Wav2Vec2 (tiny preset, fp32)
# ---------------------------------------------------------------------------------------------------------------------
ocfg = OV.Wav2Vec2Config("tiny")
w0 = OV.randomize_weights(OV.init_weights(ocfg, seed=0, dtype=torch.float64), seed=1)
Tn = OV.num_frames(ocfg, 3200)
data = []
for r in range(N):
    g = torch.Generator().manual_seed(500 + r)
    wave = torch.randn(2, 3200, generator=g, dtype=torch.float64)
    neg = OV.negative_indices_from_random(torch.randint(0, Tn, (2, Tn), generator=g), ocfg.num_negatives)
    data.append((wave, neg))
names = list(w0)
for mode in ("eager", "graph"):
    model = W2V.Wav2Vec2ForPreTraining(W2V.Wav2Vec2Config("tiny"), precision="fp32", device=local, seed=0)
    model.set_weights({k: v.float() for k, v in w0.items()})
    model.broadcast_weights(strategy)
    opt = Adam(learning_rate=3e-5, epsilon=1e-8, clipnorm=1.0)
    w = {k: v.clone() for k, v in w0.items()}
    mo = {k: torch.zeros_like(v) for k, v in w.items()}
    vo = {k: torch.zeros_like(v) for k, v in w.items()}
    wave0, neg0 = data[rank]
    def oracle_reduced(w):
        clipped, losses = [], []
        for r in range(N):
            _, g_r = OV.loss_and_grads(ocfg, w, data[r][0], data[r][1], loss_div=float(N))
            c_r, _ = T.clip_by_global_norm([g_r[k] for k in names], 1.0)
            clipped.append(dict(zip(names, c_r)))
            losses.append(float(OV.forward(ocfg, w, data[r][0], data[r][1])["loss"]) / N)
        return clipped, losses

    steps = (1, 2)
    if mode == "graph":
        # building the graphs runs ONE real distributed step on the example batch (warm-up): the oracle takes it too, unchecked
        model._sample_negative_indices = lambda T_, B_: neg0.cuda().unsqueeze(1)
        gstep, segs = W2V.make_graphed_distributed_step(strategy, model, opt, wave0.float().cuda(), dropout=False, warmup=1)
        clipped, _ = oracle_reduced(w)
        OV.train_step(ocfg, w, mo, vo, 1, wave0, neg0, lr=3e-5, eps=1e-8, num_replicas=N, peer_grads=[clipped[r] for r in range(N) if r != rank])
        steps = (2, 3)
    for t in steps:
        before = {k: v.clone() for k, v in w.items()}
        if mode == "eager":
            loss = W2V.distributed_train_step(strategy, model, (wave0.float(), None), opt, neg_indices=neg0, dropout=False)
        else:
            loss = gstep(wave0.float().cuda())
        torch.cuda.synchronize()
        clipped, losses = oracle_reduced(w)
        want = {k: sum(c[k] for c in clipped) for k in names}
        prog = model._prog
        ge = {k: rel(prog.view(prog.grads, k), want[k]) for k in names if float(want[k].abs().max()) > 1e-12}
        wg = max(ge.items(), key=lambda kv: kv[1] / (1.0 if want[kv[0]].dim() > 1 else 3.0))
        wg = (wg[0], wg[1] / (1.0 if want[wg[0]].dim() > 1 else 3.0))
        OV.train_step(ocfg, w, mo, vo, t, wave0, neg0, lr=3e-5, eps=1e-8, num_replicas=N,
                      peer_grads=[clipped[r] for r in range(N) if r != rank])
        got = model.get_weights()
        ue = {k: rel(got[k].double().cpu() - before[k], w[k] - before[k]) for k in
              ("encoder.layers.0.attention.q_proj.kernel", "fe.conv1.kernel", "project_hid.dense.kernel", "quantizer.codevectors",
               "encoder.layers.3.feed_forward.output_dense.kernel", "feature_projection.kernel")}
        report(f"w2v {mode} step {t}", wg, float(loss), sum(losses), max(ue.items(), key=lambda kv: kv[1]))
    del model, opt

# ---------------------------------------------------------------------------------------------------------------------
# Whisper (reduced config, fp32)
# ---------------------------------------------------------------------------------------------------------------------
ocfg = OW.WhisperConfig("small")
def shrink(c):
    c.d_model, c.d_ff = 128, 256
    c.encoder_layers = c.decoder_layers = 2
    c.encoder_attention_heads = c.decoder_attention_heads = 2
    c.vocab_size, c.n_mels, c.n_ctx, c.decoder_start_token_id = 203, 16, 64, 200
    return c
shrink(ocfg)
w0 = OW.randomize_weights(OW.init_weights(ocfg, seed=4, dtype=torch.float64), seed=5)
data = []
for r in range(N):
    g = torch.Generator().manual_seed(700 + r)
    data.append((torch.randn(2, ocfg.n_mels, 100, generator=g, dtype=torch.float64), torch.randint(0, 100, (2, 24), generator=g, dtype=torch.int32)))
for mode in ("eager", "graph"):
    model = WH.WhisperForConditionalGeneration(shrink(WH.WhisperConfig()), precision="fp32", device=local, seed=4)
    model.set_weights({k: v.float() for k, v in w0.items()})
    model.broadcast_weights(strategy)
    opt = Adam(learning_rate=1e-4)
    w = {k: v.clone() for k, v in w0.items()}
    mo = {k: torch.zeros_like(v) for k, v in w.items()}
    vo = {k: torch.zeros_like(v) for k, v in w.items()}
    f0, l0 = data[rank]
    steps = (1, 2)
    if mode == "graph":
        gstep, segs = WH.make_graphed_distributed_step(strategy, model, opt, f0.float().cuda(), l0.cuda(), dropout=False, warmup=1,
                                                       bucket_elems=1 << 16)
        grads = [OW.loss_and_grads(ocfg, w, data[r][0], data[r][1])[1] for r in range(N)]
        OW.train_step(ocfg, w, mo, vo, 1, f0, l0, peer_grads=[grads[r] for r in range(N) if r != rank])    # the warm-up step
        steps = (2, 3)
    for t in steps:
        before = {k: v.clone() for k, v in w.items()}
        if mode == "eager":
            loss = WH.distributed_train_step(strategy, model, (f0.float(), l0), opt, dropout=False)
        else:
            loss = gstep(f0.float().cuda(), l0.cuda())
        torch.cuda.synchronize()
        grads, losses = [], []
        for r in range(N):
            o_r, g_r = OW.loss_and_grads(ocfg, w, data[r][0], data[r][1])
            grads.append(g_r); losses.append(float(o_r["loss"]))
        want = {k: sum(g_[k] for g_ in grads) for k in grads[0]}
        prog = model._prog
        gscale = max(float(v.abs().max()) for v in want.values())
        ge = {k: rel(prog.view(prog.grads, k), want[k]) for k in want if float(want[k].abs().max()) > 1e-12 * max(1.0, gscale)}
        wg = max(ge.items(), key=lambda kv: kv[1] / (1.0 if want[kv[0]].dim() > 1 else 3.0))
        wg = (wg[0], wg[1] / (1.0 if want[wg[0]].dim() > 1 else 3.0))
        OW.train_step(ocfg, w, mo, vo, t, f0, l0, peer_grads=[grads[r] for r in range(N) if r != rank])
        got = model.get_weights()
        keys = [k for k in got if k.endswith("kernel")][:6]
        ue = {k: rel(got[k].double().cpu() - before[k], w[k] - before[k]) for k in keys}
        report(f"whisper {mode} step {t}", wg, float(loss), sum(losses), max(ue.items(), key=lambda kv: kv[1]))
    del model, opt

strategy.check()
info = strategy.info()
if rank == 0:
    print(f"[rank 0] communicator: {info}")
strategy.barrier()
print(f"[rank {rank}] DIST ORACLE CHECK {'PASSED' if ok else 'FAILED'}", flush=True)
# (the communicator is left to process exit: CUDA graphs that captured its collectives are still alive here)
if strategy.dist is not None:
    strategy.dist.destroy_process_group()
sys.exit(0 if ok else 1)
