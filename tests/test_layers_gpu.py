"""the reference's sub-layer classes (tethys_speech_b200/layers.py) against the oracle's restatement of each
layer, same weights, same inputs. fp32 operands: 1e-5 (the north star's fp32 bar); bf16 operands: 2e-2."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu

BAR = {"fp32": 1e-5, "bf16": 2e-2}


def _whisper_cfgs():
    from oracle import whisper_oracle as O
    from tethys_speech_b200 import whisper as W

    ocfg, cfg = O.WhisperConfig("small"), W.WhisperConfig()
    for c in (ocfg, cfg):
        c.d_model, c.d_ff = 128, 256
        c.encoder_layers = c.decoder_layers = 2
        c.encoder_attention_heads = c.decoder_attention_heads = 2
        c.vocab_size, c.n_mels, c.n_ctx, c.decoder_start_token_id = 203, 16, 64, 200
    return ocfg, cfg


@pytest.mark.parametrize("kind", ["self", "cross", "decoder_mask"])
def test_multi_head_attention_matches_oracle(kind):
    from oracle import whisper_oracle as O
    from tethys_speech_b200 import layers as L

    precision = "bf16"          # the single-layer attention operator (K10) is bf16; fp32 attention is covered by the whole-model tests
    ocfg, cfg = _whisper_cfgs()
    with pytest.raises(NotImplementedError):
        L.MultiHeadAttention(cfg, precision="fp32")
    layer = L.MultiHeadAttention(cfg, is_decoder=kind != "self", is_cross_attention=kind == "cross", precision=precision, seed=3)
    g = torch.Generator().manual_seed(11)
    for v in layer.trainable_variables:
        if v.dim() == 1:
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
    w = {"p." + n: v.double().cpu() for n, v in zip(layer.variable_names, layer.trainable_variables)}
    x = torch.randn(2, 40, cfg.d_model, generator=g, dtype=torch.float64)
    kv = torch.randn(2, 150, cfg.d_model, generator=g, dtype=torch.float64) if kind == "cross" else None
    mask = (1.0 - torch.tril(torch.ones(40, 40))).unsqueeze(0) if kind == "decoder_mask" else None
    want = O.mha(ocfg, w, "p.", x, kv=kv, mask=mask)
    got = layer(x.float(), key_value_states=None if kv is None else kv.float(), attention_mask=mask, training=False)
    assert rel_l2(got, want) < BAR[precision], rel_l2(got, want)
    with pytest.raises(NotImplementedError):
        layer(x.float(), attention_mask=torch.rand(1, 40, 40).round())


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_feed_forward_and_projection_head_match_oracle(precision):
    from oracle import wav2vec2_oracle as OV
    from oracle import whisper_oracle as O
    from tethys_speech_b200 import layers as L
    from tethys_speech_b200 import wav2vec2 as W

    _, cfg = _whisper_cfgs()
    ff = L.FeedForward(cfg, precision=precision, seed=5)
    g = torch.Generator().manual_seed(12)
    x = torch.randn(3, 50, cfg.d_model, generator=g, dtype=torch.float64)
    w = {"p." + n: v.double().cpu() for n, v in zip(ff.variable_names, ff.trainable_variables)}
    assert rel_l2(ff(x.float()), O.ffn(w, "p.", x)) < BAR[precision]
    vcfg = W.Wav2Vec2Config("tiny")
    head = L.Wav2Vec2ProjectionHead(vcfg, precision=precision, seed=6)
    head.gamma.copy_(1 + 0.1 * torch.randn(head.gamma.shape, generator=g)); head.beta.copy_(0.1 * torch.randn(head.beta.shape, generator=g))
    w = {"h." + n: v.double().cpu() for n, v in zip(head.variable_names, head.trainable_variables)}
    xh = torch.randn(2, 30, vcfg.hidden_size, generator=g, dtype=torch.float64)
    assert rel_l2(head(xh.float()), OV.projection_head(w, "h", xh, vcfg.layer_norm_eps)) < BAR[precision]
    # training=True draws the library's dropout mask: kept elements are scaled by 1 / (1 - rate), about `rate` of them are zero
    y0, y1 = head(xh.float()), head(xh.float(), training=True)
    kept = y1 != 0
    assert abs(float((~kept).float().mean()) - vcfg.hidden_dropout) < 0.05
    assert rel_l2(y1[kept].float() * (1 - vcfg.hidden_dropout), y0[kept].float()) < 1e-2


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_group_normalization_matches_oracle(precision):
    from oracle import tf_ops as T
    from tethys_speech_b200 import layers as L

    gn = L.GroupNormalization(groups=16, axis=-1, epsilon=1e-5, precision=precision)
    g = torch.Generator().manual_seed(13)
    x = torch.randn(2, 77, 512, generator=g, dtype=torch.float64) * 2 + 0.5       # the feature encoder's shape: 512 channels, 16 groups (V:248)
    gn.build(x.shape)
    gn.gamma.copy_(1 + 0.2 * torch.randn(512, generator=g)); gn.beta.copy_(0.2 * torch.randn(512, generator=g))
    want = T.group_norm(x, gn.gamma.double().cpu(), gn.beta.double().cpu(), 16, 1e-5)
    assert rel_l2(gn(x.float()), want) < BAR[precision]
    with pytest.raises(ValueError):
        L.GroupNormalization(groups=5, precision=precision)(torch.zeros(1, 4, 512))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_whisper_encoder_decoder_model_views_match_oracle(precision):
    from oracle import whisper_oracle as O
    from tethys_speech_b200 import layers as L

    ocfg, cfg = _whisper_cfgs()
    w0 = O.randomize_weights(O.init_weights(ocfg, seed=4, dtype=torch.float64), seed=5)
    model = L.WhisperModel(cfg, precision=precision, seed=4)
    model.set_weights({k: v.float() for k, v in w0.items()})
    g = torch.Generator().manual_seed(14)
    feats = torch.randn(2, ocfg.n_mels, 100, generator=g, dtype=torch.float64)
    ids = torch.randint(0, 100, (2, 24), generator=g, dtype=torch.int32)
    ids[:, 0] = cfg.decoder_start_token_id
    enc = O.encoder(ocfg, w0, feats)
    dec = O.decoder(ocfg, w0, ids, enc)
    got_enc = model.encoder(feats.float())["last_hidden_state"]
    assert rel_l2(got_enc, enc) < BAR[precision]
    out = model(feats.float(), decoder_input_ids=ids)
    assert rel_l2(out["encoder_last_hidden_state"], enc) < BAR[precision]
    assert rel_l2(out["last_hidden_state"], dec) < BAR[precision]
    got_enc = model.encoder(feats.float())["last_hidden_state"]
    assert rel_l2(model.decoder(ids, encoder_hidden_states=got_enc)["last_hidden_state"], dec) < BAR[precision]
    assert all(n.startswith("encoder.") for n in model.encoder.variable_names) and len(model.encoder.trainable_variables) > 0
    with pytest.raises(NotImplementedError):
        model.decoder(ids, encoder_hidden_states=torch.zeros_like(got_enc))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_wav2vec2_sublayer_views_match_oracle(precision):
    from oracle import wav2vec2_oracle as O
    from tethys_speech_b200 import layers as L
    from tethys_speech_b200 import wav2vec2 as W

    ocfg = O.Wav2Vec2Config("tiny")
    w0 = O.randomize_weights(O.init_weights(ocfg, seed=0, dtype=torch.float64), seed=1)
    owner = W.Wav2Vec2ForPreTraining(W.Wav2Vec2Config("tiny"), precision=precision, seed=0)
    owner.set_weights({k: v.float() for k, v in w0.items()})
    g = torch.Generator().manual_seed(15)
    wave = torch.randn(2, 6400, generator=g, dtype=torch.float64)
    fe = L.Wav2Vec2FeatureExtractor(owner.config, _owner=owner)
    want_fe = O.feature_extractor(ocfg, w0, wave)
    assert rel_l2(fe(wave.float()), want_fe) < BAR[precision]
    enc = L.Wav2Vec2Encoder(owner.config, _owner=owner)
    Tn = O.num_frames(ocfg, 6400)
    neg = O.negative_indices_from_random(torch.randint(0, Tn, (2, Tn), generator=g), ocfg.num_negatives)
    want = O.forward(ocfg, w0, wave, neg)
    assert rel_l2(enc(wave.float())["last_hidden_state"], want["last_hidden_state"]) < BAR[precision]
    qz = L.Wav2Vec2Quantizer(owner.config, _owner=owner)
    qf, perp, idx = qz(wave.float())
    assert idx.dtype == torch.int64 and tuple(idx.shape) == (ocfg.num_codevector_groups, 2, Tn)
    if precision == "fp32":
        assert bool((idx.cpu() == want["code_indices"]).all())          # integer work: bit-exact
        assert rel_l2(qf, want["quantized_features"]) < 1e-5
    else:   # bf16 inputs may flip a near-tie: the quantised features must be the codewords of the indices the GPU chose
        want_b = O.forward(ocfg, w0, wave, neg, code_indices=idx.cpu())
        assert rel_l2(qf, want_b["quantized_features"]) < 2e-2
    assert all(n.startswith("quantizer.") for n in qz.variable_names)


@pytest.mark.parametrize("T,K", [(750, 100), (100, 100), (60, 100), (2, 100), (1500, 100)])
def test_negative_sampler_kernel_bit_exact(T, K):
    """ts_w2v_sample_negatives == tf.nn.top_k(-float(r), k) tiled to K (V:907-937) as the oracle restates it — integer work,
    bit-exact, including ties (T draws from [0, T) collide all the time) and the T - 1 < K tiling branch."""
    import ctypes as C

    from oracle import wav2vec2_oracle as O
    from tethys_speech_b200 import _lib
    from tethys_speech_b200.runtime import ptr, stream_ptr

    ctx = _lib.context(0)
    g = torch.Generator().manual_seed(T * 31 + K)
    r = torch.randint(0, T, (5, T), generator=g)
    want = O.negative_indices_from_random(r, K)
    want = want[:, 0, :] if want.dim() == 3 else want
    rd = r.to(torch.int32).cuda()
    out = torch.empty(5, K, dtype=torch.int32, device="cuda")
    ctx.check(ctx.lib.ts_w2v_sample_negatives(ctx.h, ptr(rd), 5, T, K, ptr(out), stream_ptr()))
    assert bool((out.cpu().long() == want.long()).all())
