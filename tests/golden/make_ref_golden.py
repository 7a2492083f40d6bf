#!/usr/bin/env python
"""Generates tests/golden/ref_*.npz by RUNNING THE REFERENCE'S OWN CODE: the unmodified /root/reference/speech_jobs scripts
are imported on oracle/tf_shim.py (TensorFlow is not installable; see tests/test_reference_pinning.py) in float64, the
seeded weights of the oracle's initialisers are loaded into the reference's tf.Variables, and the reference's model classes and
step functions produce: forward outputs, loss, a few gradients, and the weights after two optimiser steps of the reference's
distributed_train_step. The fixtures travel to the GPU box (where /root/reference does not exist):
  tests/test_ref_golden.py        CPU: the oracle reproduces them to 1e-10;
  tests/test_ref_golden_gpu.py    GPU: the CUDA path (fp32 parity mode, through the C-ABI) reproduces them to 1e-5.
Run from the repo root in the build container:  python tests/golden/make_ref_golden.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_runner as R  # noqa: E402
from oracle import tf_shim  # noqa: E402
from oracle import wav2vec2_oracle as WO  # noqa: E402
from oracle import whisper_oracle as HO  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
MAX_ELEMS = 1500


def sub(t):
    """Fixtures stay small: a tensor with more than MAX_ELEMS elements is stored as a strided sample of its flattened form
    (the consumer applies the same `sub`)."""
    a = t.detach().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
    flat = a.reshape(-1)
    step = -(-flat.size // MAX_ELEMS)
    return flat[::step].copy()

W2V_GRADS = ("fe.conv0.kernel", "fe.conv2.gn.gamma", "encoder.layers.1.attention.q_proj.kernel", "encoder.layers.3.feed_forward.output_dense.bias",
             "quantizer.codevectors", "project_q.dense.kernel")
WH_GRADS = ("encoder.conv1.kernel", "encoder.layers.1.self_attn.q_proj.kernel", "decoder.embed_tokens.embeddings",
            "decoder.layers.0.self_attn.k_proj.kernel", "decoder.layers.1.encoder_attn.v_proj.bias", "lm_head.kernel")


def whisper_edit(c):
    c.d_model, c.d_ff = 64, 128
    c.encoder_layers = c.decoder_layers = 2
    c.encoder_attention_heads = c.decoder_attention_heads = 2
    c.vocab_size, c.n_mels, c.n_ctx, c.decoder_start_token_id = 203, 16, 64, 200


def w2v(seed=11, size="tiny", B=2, N=3200):
    tf = tf_shim.install()
    ref = R.load("wav2vec2_dist")
    tf_shim.seed(seed)
    ocfg = WO.Wav2Vec2Config(size)
    w0 = WO.randomize_weights(WO.init_weights(ocfg, seed, torch.float64), seed + 1)
    wave = torch.randn(B, N, generator=torch.Generator().manual_seed(seed), dtype=torch.float32).double()   # fp32-representable
    model = R.build_w2v(ref, size, wave)
    vm = R.w2v_variable_map(model)
    R.set_weights(vm, w0)
    rec = {"seed": seed, "size": size, "wave": wave.float().numpy()}
    tf_shim.RANDOM_LOG.clear()
    out = model(wave, training=True)
    logits, closs = model._compute_contrastive_loss(out["projected_states"], out["projected_quantized_features"])
    loss = closs + model.diversity_loss_weight * model._compute_diversity_loss(out["codevector_perplexity"])
    draw = [t for k, t in tf_shim.RANDOM_LOG if k == "uniform"][-1]
    rec["random_ints"] = draw.numpy()                         # the tf.random.uniform draw of V:919; negatives derive from it
    tf_shim.RANDOM_LOG.clear()
    rec["loss"] = float(loss)
    rec["contrastive_loss"] = float(closs)
    rec["perplexity"] = float(out["codevector_perplexity"])
    rec["logits_sub"] = sub(logits)
    rec["last_hidden_sub"] = sub(out["last_hidden_state"])
    rec["extract_features_sub"] = sub(out["extract_features"])
    enc = model.wav2vec2.quantizer(model.wav2vec2.feature_projection_layer_norm(model.wav2vec2.feature_projection(out["extract_features"])), training=True)["encodings"]
    rec["code_indices"] = enc.argmax(-1).numpy().astype(np.int64)          # [G,B,T] from the reference's one-hot encodings
    gr = R.grads_by_name(vm, model, loss)
    for k in W2V_GRADS:
        rec["grad::" + k] = sub(gr[k])
    # two steps of the reference's distributed_train_step (V:1186-1260), single replica
    strategy = tf.distribute.MultiWorkerMirroredStrategy()
    opt = tf.keras.optimizers.Adam(learning_rate=3e-5, beta_1=0.9, beta_2=0.999, epsilon=1e-8, clipnorm=1.0)
    draws, losses = [], []
    for _ in range(2):
        tf_shim.RANDOM_LOG.clear()
        losses.append(float(ref.distributed_train_step(strategy, model, (wave, torch.zeros(B)), opt)))
        draws.append([t for k, t in tf_shim.RANDOM_LOG if k == "uniform"][-1].numpy())
    rec["step_random_ints"] = np.stack(draws)
    rec["step_losses"] = np.array(losses)
    for k in W2V_GRADS:
        rec["delta2::" + k] = sub(vm[k].detach() - w0[k])                     # weight change after 2 steps
    np.savez_compressed(os.path.join(OUT, "ref_w2v_tiny.npz"), **rec)


def whisper(seed=13, B=2, Tm=100, S=12):
    tf = tf_shim.install()
    ref = R.load("whisper_dist")
    tf_shim.seed(seed)
    for name, dtype in (("f64", torch.float64), ("f32", torch.float32)):
        tf_shim.set_floatx(dtype)
        ocfg = HO.WhisperConfig("small")
        whisper_edit(ocfg)
        w0 = HO.randomize_weights(HO.init_weights(ocfg, seed, dtype), seed + 1)
        g = torch.Generator().manual_seed(seed)
        feats = torch.randn(B, ocfg.n_mels, Tm, generator=g, dtype=torch.float32).to(dtype)
        labels = torch.randint(0, 100, (B, S), generator=g, dtype=torch.int32)
        model = R.build_whisper(ref, whisper_edit, feats, labels)
        vm = R.whisper_variable_map(model)
        R.set_weights(vm, w0)
        out = model(feats, labels=labels, training=True)
        if name == "f32":
            # float32 run: TF's own arithmetic for the anti-causal mask (score + (-1e9) absorbs the score, App. C-1)
            np.savez_compressed(os.path.join(OUT, "ref_whisper_small_cfg_f32.npz"), seed=seed, feats=feats.float().numpy(), labels=labels.numpy(),
                                loss=float(out["loss"]), logits_sub=sub(out["logits"]))
            break
        rec = {"seed": seed, "feats": feats.float().numpy(), "labels": labels.numpy(), "loss": float(out["loss"]),
               "logits_sub": sub(out["logits"]), "encoder_sub": sub(out["encoder_last_hidden_state"])}
        gr = R.grads_by_name(vm, model, out["loss"])
        for k in WH_GRADS:
            rec["grad::" + k] = sub(gr[k])
        strategy = tf.distribute.MultiWorkerMirroredStrategy()
        opt = tf.keras.optimizers.Adam(learning_rate=1e-4)
        rec["step_losses"] = np.array([float(ref.distributed_train_step(strategy, model, (feats, labels), opt)) for _ in range(2)])
        for k in WH_GRADS:
            rec["delta2::" + k] = sub(vm[k].detach() - w0[k])
        np.savez_compressed(os.path.join(OUT, "ref_whisper_small_cfg.npz"), **rec)
    tf_shim.set_floatx(torch.float64)


if __name__ == "__main__":
    assert R.available(), "needs /root/reference (build container only)"
    w2v()
    whisper()
    print({f: os.path.getsize(os.path.join(OUT, f)) for f in sorted(os.listdir(OUT)) if f.startswith("ref_")})
