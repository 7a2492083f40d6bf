"""The CUDA distributed train steps against the oracle's N-replica step (oracle.train_step(peer_grads=...)) for BOTH reduce
conventions of the reference:
  Wav2Vec2  V:1186-1260  loss / N -> gradients -> LOCAL clip_by_global_norm(1.0) -> all-reduce SUM -> per-variable clipnorm(1.0) -> Adam
  Whisper   W:819-848    gradients of the local mean loss -> all-reduce SUM, NOT divided by N -> Adam; returned loss = SUM of replica losses
This file runs on ONE GPU (the driver's `-m gpu` box): the second replica is a sibling model object on the same device whose
locally clipped gradient arena is added where the NCCL all-reduce would add it (a Strategy whose collectives are sums over the
emulated peers) — every kernel of the distributed path (loss scale, local clip, clipnorm, Adam, bf16 compute) runs for real and
is compared with the ORACLE, not with another CUDA run. The same comparison over real NCCL on 2 GPUs is
tools/check_dist_oracle.py (torchrun; log under profiles/)."""
import pytest
import torch

from conftest import check_bf16_grads, rel_l2

pytestmark = pytest.mark.gpu


def _peer_strategy(peer_arenas, peer_losses):
    from tethys_speech_b200.runtime import Strategy

    class PeerStrategy(Strategy):
        def __init__(self):                      # no process group: the peers live in this process
            self.world, self.rank, self.local_rank, self.dist = 1 + len(peer_arenas), 0, 0, None

        def all_reduce_sum_(self, flat, bucket_elems=None):
            for a in peer_arenas:
                flat.add_(a)

        def reduce(self, op, value, axis=None):
            return value + sum(peer_losses)

    return PeerStrategy()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_w2v_two_replica_step_matches_oracle(precision):
    from oracle import tf_ops as T
    from oracle import wav2vec2_oracle as O
    from tethys_speech_b200 import wav2vec2 as W

    N = 2
    ocfg = O.Wav2Vec2Config("tiny")
    w0 = O.randomize_weights(O.init_weights(ocfg, seed=0, dtype=torch.float64), seed=1)
    reps = [W.Wav2Vec2ForPreTraining(W.Wav2Vec2Config("tiny"), precision=precision, seed=0) for _ in range(N)]
    opts = [W.Adam(learning_rate=3e-5, epsilon=1e-8, clipnorm=1.0) for _ in range(N)]
    for m_ in reps:
        m_.set_weights({k: v.float() for k, v in w0.items()})
    Tn = O.num_frames(ocfg, 3200)
    data = []
    for r in range(N):
        g = torch.Generator().manual_seed(500 + r)
        wave = torch.randn(2, 3200, generator=g, dtype=torch.float64)
        neg = O.negative_indices_from_random(torch.randint(0, Tn, (2, Tn), generator=g), ocfg.num_negatives)
        data.append((wave, neg))
    w = {k: v.clone() for k, v in w0.items()}
    mo = {k: torch.zeros_like(v) for k, v in w.items()}
    vo = {k: torch.zeros_like(v) for k, v in w.items()}
    names = list(w)
    for t in (1, 2):
        before = {k: v.clone() for k, v in w.items()}
        # replica 1 (the peer): forward with loss / N, backward, local clip (V:1231-1243); its arena is what NCCL would add
        wave1, neg1 = data[1]
        out1 = reps[1](wave1.float(), training=True, neg_indices=neg1, loss_div=float(N), dropout=False)
        idx1 = out1["code_indices"].cpu().clone()
        reps[1].gradient()
        opts[1].local_clip(reps[1], 1.0)
        peer_arena = reps[1]._prog.grads.clone()
        peer_loss = (out1["loss"] / N).clone()
        # replica 0: the product's distributed step; its all-reduce adds the peer's arena
        strategy = _peer_strategy([peer_arena], [peer_loss])
        wave0, neg0 = data[0]
        loss = W.distributed_train_step(strategy, reps[0], (wave0.float(), None), opts[0], neg_indices=neg0, dropout=False)
        idx0 = reps[0]._last["out"]["code_indices"].cpu().clone()
        torch.cuda.synchronize()
        # oracle: replica 1's locally clipped gradients as peer_grads of replica 0's step
        inj = precision != "fp32"
        _, g1 = O.loss_and_grads(ocfg, w, wave1, neg1, loss_div=float(N), code_indices=idx1 if inj else None)
        c1, _ = T.clip_by_global_norm([g1[k] for k in names], 1.0)
        # the gradient arena after the all-reduce = sum over replicas of the locally clipped gradients (before clipnorm)
        _, g0 = O.loss_and_grads(ocfg, w, wave0, neg0, loss_div=float(N), code_indices=idx0 if inj else None)
        c0, _ = T.clip_by_global_norm([g0[k] for k in names], 1.0)
        want = {k: a + b for k, a, b in zip(names, c0, c1)}
        prog = reps[0]._prog
        gerrs = {k: rel_l2(prog.view(prog.grads, k), want[k]) for k in names if float(want[k].abs().max()) > 1e-12}
        if precision == "fp32":
            badg = {k: v for k, v in gerrs.items() if not v <= (1e-5 if want[k].dim() > 1 else 3e-5)}
            assert not badg, (t, badg)
        else:
            with T.bf16_storage():
                _, e0 = O.loss_and_grads(ocfg, w, wave0, neg0, loss_div=float(N), code_indices=idx0)
                _, e1 = O.loss_and_grads(ocfg, w, wave1, neg1, loss_div=float(N), code_indices=idx1)
            ec0, _ = T.clip_by_global_norm([e0[k] for k in names], 1.0)
            ec1, _ = T.clip_by_global_norm([e1[k] for k in names], 1.0)
            emu = {k: rel_l2(a + b, want[k]) for k, a, b in zip(names, ec0, ec1) if k in gerrs}
            assert not check_bf16_grads(f"w2v 2-replica reduced gradient step {t}", gerrs, emu)
        o1 = O.forward(ocfg, w, wave1, neg1, code_indices=idx1 if inj else None)
        oout = O.train_step(ocfg, w, mo, vo, t, wave0, neg0, lr=3e-5, eps=1e-8, num_replicas=N, peer_grads=[dict(zip(names, c1))],
                            code_indices=idx0 if inj else None)
        want_loss = float(oout["loss"]) / N + float(o1["loss"]) / N          # V:1260: SUM of the scaled losses
        ltol = 1e-4 if precision == "fp32" else 2e-2
        assert abs(float(loss) - want_loss) <= ltol * abs(want_loss), (t, float(loss), want_loss)
        got = reps[0].get_weights()
        worst = 0.0
        for k in ("encoder.layers.0.attention.q_proj.kernel", "fe.conv1.kernel", "project_hid.dense.kernel", "quantizer.codevectors",
                  "encoder.layers.3.feed_forward.output_dense.kernel", "feature_projection.kernel"):
            d_gpu = got[k].double().cpu() - before[k]
            d_ref = w[k] - before[k]
            e = rel_l2(d_gpu, d_ref)
            worst = max(worst, e)
            # Adam's first updates are ~ lr * sign(g): compare the CHANGE. Asserted in fp32 only: in bf16 the gradients carry
            # ~1e-2 relative error (checked above against their budget), which a sign-like update amplifies wherever |g| is
            # near zero; the clipnorm + Adam kernel itself is the same fp32 code in both modes
            if precision == "fp32":
                assert e < 2e-3, (t, k, e)
        print(f"[w2v 2-replica {precision} step {t}] loss gpu {float(loss):.6f} oracle {want_loss:.6f}; worst update error {worst:.2e}")
        reps[1].set_weights(reps[0].get_weights())                            # mirrored variables


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_whisper_two_replica_step_matches_oracle(precision, monkeypatch):
    from oracle import whisper_oracle as O
    from tethys_speech_b200 import whisper as W

    monkeypatch.setenv("TETHYS_NO_OVERLAP", "1")      # the emulated peers have no NCCL stream to overlap with
    N = 2
    ocfg = O.WhisperConfig("small")
    cfgs = [W.WhisperConfig() for _ in range(N)]
    for c in [ocfg] + cfgs:
        c.d_model, c.d_ff = 128, 256
        c.encoder_layers = c.decoder_layers = 2
        c.encoder_attention_heads = c.decoder_attention_heads = 2
        c.vocab_size, c.n_mels, c.n_ctx, c.decoder_start_token_id = 203, 16, 64, 200
    w0 = O.randomize_weights(O.init_weights(ocfg, seed=4, dtype=torch.float64), seed=5)
    reps = [W.WhisperForConditionalGeneration(c, precision=precision, seed=4) for c in cfgs]
    opts = [W.Adam(learning_rate=1e-4) for _ in range(N)]
    for m_ in reps:
        m_.set_weights({k: v.float() for k, v in w0.items()})
    data = []
    for r in range(N):
        g = torch.Generator().manual_seed(700 + r)
        data.append((torch.randn(2, ocfg.n_mels, 100, generator=g, dtype=torch.float64),
                     torch.randint(0, 100, (2, 24), generator=g, dtype=torch.int32)))
    w = {k: v.clone() for k, v in w0.items()}
    mo = {k: torch.zeros_like(v) for k, v in w.items()}
    vo = {k: torch.zeros_like(v) for k, v in w.items()}
    for t in (1, 2):
        before = {k: v.clone() for k, v in w.items()}
        f1, l1 = data[1]
        out1 = reps[1](f1.float(), labels=l1, training=True, dropout=False)
        reps[1].gradient()
        peer_arena = reps[1]._prog.grads.clone()
        peer_loss = out1["loss"].clone()
        strategy = _peer_strategy([peer_arena], [peer_loss])
        f0, l0 = data[0]
        loss = W.distributed_train_step(strategy, reps[0], (f0.float(), l0), opts[0], dropout=False)
        torch.cuda.synchronize()
        o1, g1 = O.loss_and_grads(ocfg, w, f1, l1)
        _, g0 = O.loss_and_grads(ocfg, w, f0, l0)
        prog = reps[0]._prog
        gscale = max(float(v.abs().max()) for v in g0.values())
        want = {k: g0[k] + g1[k] for k in g0}
        gerrs = {k: rel_l2(prog.view(prog.grads, k), want[k]) for k in want if float(want[k].abs().max()) > 1e-12 * max(1.0, gscale)}
        if precision == "fp32":
            badg = {k: v for k, v in gerrs.items() if not v <= (1e-5 if want[k].dim() > 1 else 3e-5)}
            assert not badg, (t, badg)
        else:
            from oracle import tf_ops as T

            with T.bf16_storage():
                _, e0 = O.loss_and_grads(ocfg, w, f0, l0)
                _, e1 = O.loss_and_grads(ocfg, w, f1, l1)
            assert not check_bf16_grads(f"whisper 2-replica reduced gradient step {t}", gerrs,
                                        {k: rel_l2(e0[k] + e1[k], want[k]) for k in gerrs})
        oout = O.train_step(ocfg, w, mo, vo, t, f0, l0, peer_grads=[g1])      # W:829-836: raw SUM, no 1/N
        want_loss = float(oout["loss"]) + float(o1["loss"])                    # W:848: SUM of the replica losses
        ltol = 1e-4 if precision == "fp32" else 2e-2
        assert abs(float(loss) - want_loss) <= ltol * abs(want_loss), (t, float(loss), want_loss)
        got = reps[0].get_weights()
        worst = 0.0
        for k in ("lm_head.kernel", "encoder.conv1.kernel", "decoder.layers.1.encoder_attn.k_proj.kernel", "encoder.layers.0.self_attn.q_proj.kernel",
                  "decoder.layers.0.feed_forward.fc1.kernel"):
            e = rel_l2(got[k].double().cpu() - before[k], w[k] - before[k])
            worst = max(worst, e)
            if precision == "fp32":      # see the Wav2Vec2 test: the update is sign-like, bf16 is judged on the gradients
                assert e < 5e-3, (t, k, e)
        print(f"[whisper 2-replica {precision} step {t}] loss gpu {float(loss):.6f} oracle {want_loss:.6f}; worst update error {worst:.2e}")
        reps[1].set_weights(reps[0].get_weights())
