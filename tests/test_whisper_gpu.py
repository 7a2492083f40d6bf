"""GPU parity of the Whisper train step (through the C-ABI) against the CPU oracle on identical seeded inputs and
weights. fp32 mode: 1e-5 relative (L2, oracle in fp64; 3e-5 for 1-D bias/norm gradients, which are cancellation-prone
column sums); bf16 mode: 2e-2 relative on activations, loss and gradients, a gradient's bar lifted only to 1.5 x the error
an independent CPU emulation of bf16 storage shows for that tensor (conftest.check_bf16_grads). Includes BASELINE-size
parity (default preset, 30 s of mel frames, full 51 865 vocabulary). Dropout off.
Covers the reference quirks of App. C: anti-causal mask + uniform last row, double label shift, untied lm_head."""
import numpy as np
import pytest
import torch

from conftest import BF16_TOL, check_bf16_grads, rel_l2

pytestmark = pytest.mark.gpu


def _small_cfgs(vocab=203, d=64, heads=2, ff=128, layers=2, n_mels=16, n_ctx=64, start=200):
    from oracle import whisper_oracle as O
    from tethys_speech_b200 import whisper as W

    ocfg = O.WhisperConfig("small")
    cfg = W.WhisperConfig()
    for c in (ocfg, cfg):
        c.d_model, c.d_ff = d, ff
        c.encoder_layers = c.decoder_layers = layers
        c.encoder_attention_heads = c.decoder_attention_heads = heads
        c.vocab_size, c.n_mels, c.n_ctx, c.decoder_start_token_id = vocab, n_mels, n_ctx, start
    return O, W, ocfg, cfg


def _run(O, W, ocfg, cfg, B, Tm, S, precision, tol, gtol=None, seed=0):
    w64 = O.randomize_weights(O.init_weights(ocfg, seed=seed, dtype=torch.float64), seed=seed + 1)
    model = W.WhisperForConditionalGeneration(cfg, precision=precision, seed=seed)
    model.set_weights({k: v.float() for k, v in w64.items()})
    g = torch.Generator().manual_seed(7 + seed)
    feats = torch.randn(B, ocfg.n_mels, Tm, generator=g, dtype=torch.float64)
    labels = O.dummy_labels(np.random.default_rng(seed), B, S) if S >= 90 else torch.randint(0, min(100, ocfg.vocab_size), (B, S), generator=g, dtype=torch.int32)
    out = model(feats.float(), labels=labels, training=True, dropout=False)
    grads = model.gradient()
    torch.cuda.synchronize()
    model._prog.ctx.watchdog()
    oout, og = O.loss_and_grads(ocfg, w64, feats, labels)
    errs = {"encoder_last_hidden_state": rel_l2(out["encoder_last_hidden_state"], oout["encoder_last_hidden_state"]),
            "last_hidden_state": rel_l2(out["last_hidden_state"], oout["last_hidden_state"]),
            "logits": rel_l2(out["logits"], oout["logits"]),
            "loss": abs(float(out["loss"]) - float(oout["loss"])) / abs(float(oout["loss"]))}
    bad = {k: v for k, v in errs.items() if not v <= tol}
    assert not bad, f"forward mismatch (tol {tol}): {bad}; all {errs}"
    # App. C-1: the last decoder query sees only masked keys -> exactly uniform attention 1/S
    if precision == "fp32":
        P0 = model._prog.buffer("decoder_self_attn_probs0").float()[:, :, S - 1, :S]
        assert torch.allclose(P0, torch.full_like(P0, 1.0 / S), rtol=1e-6)
        # ... and query i attends only to keys j > i
        Pfull = model._prog.buffer("decoder_self_attn_probs0").float()[0, 0, :S, :S]
        assert float(torch.tril(Pfull[:-1], diagonal=0).abs().max()) == 0.0
    else:
        # the fused kernels keep only (row max, log row-sum): a uniform row over S keys at -1e9 has max -1e9 and sum S
        # (the mask itself is checked element-wise in tests/test_attention_gpu.py)
        st = model._prog.buffer("decoder_self_attn_stats0")[:, :, S - 1, :]
        assert float(st[..., 0].max()) <= -9.0e8
        assert torch.allclose(st[..., 1], torch.full_like(st[..., 1], float(np.log(S))), rtol=1e-5)
    gerrs = {}
    gscale = max(float(v.abs().max()) for v in og.values())
    for name, gg in zip(model.variable_names, grads):
        ref = og[name]
        if float(ref.abs().max()) < 1e-12 * max(1.0, gscale):
            assert float(gg.abs().max()) < (1e-5 if precision == "fp32" else 2e-2) * gscale, name
            continue
        gerrs[name] = rel_l2(gg, ref)
    worst = sorted(gerrs.items(), key=lambda kv: -kv[1])[:5]
    tag = f"whisper {precision} d={ocfg.d_model} B={B} Tm={Tm} S={S}"
    print(f"[{tag}] fwd {errs}; worst grads {worst}")
    if precision == "fp32":
        badg = {k: v for k, v in gerrs.items() if not v <= (gtol if og[k].dim() > 1 else 3 * gtol)}
        assert not badg, f"gradient mismatch (tol {gtol}): {len(badg)} tensors; worst {worst}"
    else:
        from oracle import tf_ops

        with tf_ops.bf16_storage():
            _, eg = O.loss_and_grads(ocfg, w64, feats, labels)
        badg = check_bf16_grads(tag, gerrs, {k: rel_l2(eg[k], og[k]) for k in gerrs})
        assert not badg, f"bf16 gradients over budget: {badg}"
    return model, w64, feats, labels


def test_whisper_small_config_fp32():
    O, W, ocfg, cfg = _small_cfgs()
    _run(O, W, ocfg, cfg, 2, 100, 12, "fp32", 1e-5, 1e-5)


def test_whisper_reference_label_layout_fp32():
    # labels exactly as W:795-809 (BOS, random 3..99, EOS, zero padding) at S=100; odd encoder length T=51 (Tp=56)
    O, W, ocfg, cfg = _small_cfgs(n_ctx=64)
    _run(O, W, ocfg, cfg, 3, 102, 100, "fp32", 1e-5, 1e-5, seed=2)


def test_whisper_small_config_bf16():
    O, W, ocfg, cfg = _small_cfgs(d=128, heads=2, ff=256)
    _run(O, W, ocfg, cfg, 2, 128, 16, "bf16", BF16_TOL)


def test_whisper_tiny_preset_bf16_full_vocab():
    # real 'tiny' preset (d384, 6 heads, 4+4 layers, vocab 51865 -> padded lm_head stride 51872)
    from oracle import whisper_oracle as O
    from tethys_speech_b200 import whisper as W

    ocfg = O.WhisperConfig("tiny")
    model_cfg = W.create_whisper_model.__globals__["WhisperConfig"]()
    model_cfg.d_model, model_cfg.encoder_layers, model_cfg.decoder_layers, model_cfg.d_ff = 384, 4, 4, 1536
    model_cfg.encoder_attention_heads = model_cfg.decoder_attention_heads = 6
    _run(O, W, ocfg, model_cfg, 2, 200, 24, "bf16", BF16_TOL)


def _default_preset():
    from oracle import whisper_oracle as O
    from tethys_speech_b200 import whisper as W

    return O, W, O.WhisperConfig("small"), W.create_whisper_model("small", precision="bf16").config


def test_whisper_default_preset_bf16_baseline_size_30s():
    # BASELINE.json configs[0] as intended: CLI-default preset d768 / 12 heads / 4+4 layers, [B,80,3000] mel
    # frames (30 s), S = 100 labels laid out as W:795-809, vocabulary 51 865: the benchmarked tcgen05 / fused-attention path
    O, W, ocfg, cfg = _default_preset()
    _run(O, W, ocfg, cfg, 1, 3000, 100, "bf16", BF16_TOL)


def test_whisper_default_preset_fp32_baseline_size_30s():
    O, W, ocfg, cfg = _default_preset()
    _run(O, W, ocfg, cfg, 1, 3000, 100, "fp32", 1e-5, 1e-5)


def test_whisper_train_steps_fp32_match_oracle_adam():
    """W:819-848 + W:901: three Adam(1e-4, eps 1e-7) steps without clipping."""
    O, W, ocfg, cfg = _small_cfgs()
    w64 = O.randomize_weights(O.init_weights(ocfg, seed=4, dtype=torch.float64), seed=5)
    model = W.WhisperForConditionalGeneration(cfg, precision="fp32", seed=4)
    model.set_weights({k: v.float() for k, v in w64.items()})
    g = torch.Generator().manual_seed(11)
    feats = torch.randn(2, ocfg.n_mels, 100, generator=g, dtype=torch.float64)
    labels = torch.randint(0, 100, (2, 12), generator=g, dtype=torch.int32)
    opt = W.Adam(learning_rate=1e-4)
    w = {k: v.clone() for k, v in w64.items()}
    m = {k: torch.zeros_like(v) for k, v in w.items()}
    v_ = {k: torch.zeros_like(v) for k, v in w.items()}
    for t in range(1, 4):
        loss = W.train_step(model, (feats.float(), labels), opt, dropout=False)
        oout = O.train_step(ocfg, w, m, v_, t, feats, labels)
        assert abs(float(loss) - float(oout["loss"])) / abs(float(oout["loss"])) < 1e-4
    got = model.get_weights()
    for k in ("lm_head.kernel", "decoder.embed_tokens.embeddings", "encoder.conv1.kernel", "decoder.layers.1.encoder_attn.k_proj.kernel",
              "encoder.layers.0.self_attn.q_proj.bias", "decoder.layer_norm.gamma"):
        d_gpu = got[k].double().cpu() - w64[k]
        d_ref = w[k] - w64[k]
        assert rel_l2(d_gpu, d_ref) < 5e-3, (k, rel_l2(d_gpu, d_ref))


def test_whisper_bf16_ragged_cross_attention_and_full_length_targets():
    # encoder T = 300 (three 128-row tiles, ragged), decoder S = 100 (W:786): cross-attention 100 x 300, anti-causal 100 x 100
    O, W, ocfg, cfg = _small_cfgs(d=128, heads=2, ff=256, n_ctx=320)
    _run(O, W, ocfg, cfg, 1, 600, 100, "bf16", BF16_TOL, seed=3)


def test_whisper_call_contract_without_labels_and_argument_errors():
    """W:547-616: model(features, decoder_input_ids=ids) returns logits and no loss; ids equal to the right-shifted labels are
    accepted next to labels; arguments without a kernel raise."""
    from oracle import whisper_oracle as O
    from tethys_speech_b200 import whisper as W

    ocfg, cfg = O.WhisperConfig("small"), W.WhisperConfig()
    for c in (ocfg, cfg):
        c.d_model, c.d_ff = 128, 256
        c.encoder_layers = c.decoder_layers = 2
        c.encoder_attention_heads = c.decoder_attention_heads = 2
        c.vocab_size, c.n_mels, c.n_ctx, c.decoder_start_token_id = 203, 16, 64, 200
    w0 = O.randomize_weights(O.init_weights(ocfg, seed=4, dtype=torch.float64), seed=5)
    model = W.WhisperForConditionalGeneration(cfg, precision="fp32", seed=4)
    model.set_weights({k: v.float() for k, v in w0.items()})
    g = torch.Generator().manual_seed(9)
    feats = torch.randn(2, ocfg.n_mels, 100, generator=g, dtype=torch.float64)
    labels = torch.randint(0, 100, (2, 24), generator=g, dtype=torch.int32)
    ids = torch.cat([torch.full((2, 1), cfg.decoder_start_token_id, dtype=torch.int32), labels[:, :-1]], dim=1)
    want = O.forward(ocfg, w0, feats, labels)
    out = model(feats.float(), decoder_input_ids=ids)
    assert out["loss"] is None and out["past_key_values"] is None
    assert rel_l2(out["logits"], want["logits"]) < 1e-5
    out = model(feats.float(), decoder_input_ids=ids, labels=labels, training=True, dropout=False)
    assert abs(float(out["loss"]) - float(want["loss"])) < 1e-5 * abs(float(want["loss"]))
    with pytest.raises(NotImplementedError):
        model(feats.float(), labels=labels, decoder_attention_mask=torch.ones(2, 24), training=True)
    with pytest.raises(NotImplementedError):
        model(feats.float(), decoder_input_ids=ids + 1)
    with pytest.raises(ValueError):
        model(feats.float())
