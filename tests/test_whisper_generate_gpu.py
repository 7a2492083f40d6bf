"""Greedy decoding — WhisperForConditionalGeneration.generate (W:636-709) against the oracle's restatement:
same token ids in fp32 mode; in bf16 mode every chosen token must be the oracle's argmax up to bf16 rounding of the logits
(teacher-forced on the GPU's own prefix). Covers decoder lengths 1..N (the anti-causal mask with a single, fully masked
position included), the cached cross-attention K/V, the all-sequences-EOS stop and the argument errors."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _cfgs(d, heads, ff, vocab=203, layers=2, n_mels=16, n_ctx=64, start=200):
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


def _model(O, W, ocfg, cfg, precision, seed, scale=0.05):
    w64 = O.randomize_weights(O.init_weights(ocfg, seed=seed, dtype=torch.float64), seed=seed + 1, scale=scale)
    model = W.WhisperForConditionalGeneration(cfg, precision=precision, seed=seed)
    model.set_weights({k: v.float() for k, v in w64.items()})
    return w64, model


def test_generate_fp32_matches_oracle_tokens():
    O, W, ocfg, cfg = _cfgs(64, 2, 128)
    w64, model = _model(O, W, ocfg, cfg, "fp32", 0)
    feats = torch.randn(3, ocfg.n_mels, 100, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    ids = model.generate(feats.float(), max_length=14, sync_every=5)
    ref = O.generate(ocfg, w64, feats, max_length=14)
    assert ids.dtype == torch.int32 and tuple(ids.shape) == tuple(ref.shape) == (3, 15)
    assert torch.equal(ids.cpu().long(), ref), (ids.cpu(), ref)
    assert len(set(ref[:, 1:].reshape(-1).tolist())) > 3          # not a degenerate constant output
    # logits of the last step against the oracle's decoder run on the same prefix
    dec = O.decoder(ocfg, w64, ref[:, :-1], O.encoder(ocfg, w64, feats))
    want = dec[:, -1, :] @ w64["lm_head.kernel"]
    got = model._prog.buffer("next_token_logits")[:, :ocfg.vocab_size]
    assert rel_l2(got, want) < 1e-5
    model._prog.ctx.watchdog()


def test_generate_bf16_fused_attention_tokens_are_oracle_argmax_up_to_rounding():
    O, W, ocfg, cfg = _cfgs(128, 2, 256)                          # head_dim 64: the tcgen05 attention kernels, lengths 1..20
    w64, model = _model(O, W, ocfg, cfg, "bf16", 2)
    feats = torch.randn(2, ocfg.n_mels, 120, generator=torch.Generator().manual_seed(3), dtype=torch.float64)
    ids = model.generate(feats.float(), max_length=20).cpu().long()
    assert tuple(ids.shape) == (2, 21) and bool((ids[:, 0] == cfg.decoder_start_token_id).all())
    enc = O.encoder(ocfg, w64, feats)
    exact = 0
    for L in range(1, 21):
        logits = O.decoder(ocfg, w64, ids[:, :L], enc)[:, -1, :] @ w64["lm_head.kernel"]
        chosen = logits.gather(1, ids[:, L:L + 1]).squeeze(1)
        gap = logits.max(dim=1).values - chosen
        assert bool((gap <= 2e-2 * logits.abs().max(dim=1).values).all()), (L, gap, ids[:, L])
        exact += int((gap == 0).sum())
    assert exact >= 30, exact                                      # most of the 40 picks are the exact argmax
    model._prog.ctx.watchdog()


def test_generate_stops_when_every_sequence_emits_eos():
    """gamma = 0 on the final decoder LayerNorm makes the decoder output the constant beta; with lm_head[:, eos] = 10 beta the
    EOS logit wins at every position, so generation stops after the first step (W:697-705) whatever `sync_every` is."""
    O, W, ocfg, cfg = _cfgs(64, 2, 128)
    w64, model = _model(O, W, ocfg, cfg, "fp32", 5)
    beta = torch.randn(64, generator=torch.Generator().manual_seed(9), dtype=torch.float64)
    head = w64["lm_head.kernel"].clone()
    head[:, cfg.eos_token_id] = 10 * beta
    model.set_weights({"decoder.layer_norm.gamma": torch.zeros(64), "decoder.layer_norm.beta": beta.float(), "lm_head.kernel": head.float()})
    feats = torch.randn(2, ocfg.n_mels, 60, generator=torch.Generator().manual_seed(1))
    for sync_every in (1, 8):
        ids = model.generate(feats, max_length=12, sync_every=sync_every).cpu()
        assert ids.tolist() == [[cfg.decoder_start_token_id, cfg.eos_token_id]] * 2, ids
    # without the rigged head the same model runs to max_length
    model.set_weights({"lm_head.kernel": w64["lm_head.kernel"].float(), "decoder.layer_norm.gamma": torch.ones(64)})
    assert tuple(model.generate(feats, max_length=12).shape) == (2, 13)


def test_generate_argument_errors_and_training_after_generation():
    from tethys_speech_b200 import TethysError

    O, W, ocfg, cfg = _cfgs(64, 2, 128)
    w64, model = _model(O, W, ocfg, cfg, "fp32", 7)
    feats = torch.randn(2, ocfg.n_mels, 60, generator=torch.Generator().manual_seed(1))
    with pytest.raises(NotImplementedError):
        model.generate(feats, max_length=4, num_beams=4)
    with pytest.raises(ValueError):
        model.generate(feats, max_length=cfg.max_target_positions + 1)
    with pytest.raises(ValueError):
        model.generate(feats, max_length=4, temperature=0.0)
    ids = model.generate(feats, max_length=1)                      # a single step: decoder length 1
    assert tuple(ids.shape) == (2, 2)
    # a training call re-plans the workspace; a decode step afterwards must be refused, not read stale encoder state
    labels = torch.randint(3, 100, (2, 10), generator=torch.Generator().manual_seed(2), dtype=torch.int32)
    out = model(feats, labels=labels, training=True, dropout=False)
    ref = O.forward(ocfg, w64, feats.double(), labels)
    assert abs(float(out["loss"]) - float(ref["loss"])) < 1e-5 * abs(float(ref["loss"]))
    p = model._prog
    tok = torch.zeros(2, 4, dtype=torch.int32, device=p.device)
    with pytest.raises(TethysError):
        p.ctx.check(p.lib.ts_whisper_decode_step(p.h, tok.data_ptr(), 4, 1, None))
