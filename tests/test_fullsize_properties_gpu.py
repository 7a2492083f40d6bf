"""BASELINE-size runs (Wav2Vec2-base on 15 s audio, Whisper default preset on 30 s) cannot be compared with the CPU oracle in test
time; they are checked through properties that do not depend on the size (bf16 mode, dropout off):
  * the hard quantiser returns, for every frame and group, a code of minimal distance to the GPU's own quantiser input, and the
    perplexity is exp(entropy) of the histogram of the returned indices (V:627-660);
  * the contrastive logits are <projected state, projected quantised feature> / temperature at the sampled positions (V:876-888);
  * a sample's encoder output does not depend on which other samples share its batch (no op of the trunk mixes batch items);
  * Whisper's loss is the mean shifted sparse cross-entropy of its own logits over the real 51 865-entry vocabulary (W:585-600);
  * an optimiser step with learning rate 0 leaves every parameter bit-identical and the bf16 compute copy in sync (V:1243-1246)."""
import math

import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def test_w2v_base_15s_properties():
    from tethys_speech_b200 import wav2vec2 as W
    from tethys_speech_b200.runtime import Adam

    B, N = 4, 240000
    cfg = W.Wav2Vec2Config("base")
    model = W.Wav2Vec2ForPreTraining(cfg, precision="bf16", device=0, seed=1)
    g = torch.Generator().manual_seed(21)
    wave = torch.randn(B, N, generator=g).cuda()
    T = model.num_frames(N)
    assert T == 750
    neg = model._sample_negative_indices(T, B)[:, 0, :].contiguous()
    out = model(wave, training=True, neg_indices=neg, dropout=False)
    torch.cuda.synchronize()
    model._prog.ctx.watchdog()
    M, G, V = B * T, cfg.num_codevector_groups, cfg.num_codevectors_per_group
    idx = out["code_indices"].clone()                                   # [G, B, T] int64
    assert idx.dtype == torch.int64 and tuple(idx.shape) == (G, B, T) and int(idx.min()) >= 0 and int(idx.max()) < V
    # -- quantiser optimality + perplexity ---------------------------------------------------------------------------------
    z = model._prog.buffer("quantizer_input").float().reshape(M, G, -1)
    cb = model.get_weights()["quantizer.codevectors"]                    # [G, V, Dg] fp32
    perps = []
    for gi in range(G):
        d = ((z[:, gi, None, :] - cb[gi][None]) ** 2).sum(-1)            # [M, V]
        chosen = d.gather(1, idx[gi].reshape(M, 1)).squeeze(1)
        best = d.min(dim=1).values
        assert bool((chosen <= best * (1 + 1e-5) + 1e-6).all()), f"group {gi}: a returned code is not a nearest code"
        q = out["quantized_features"].float().reshape(M, G, -1)[:, gi]
        assert rel_l2(q, cb[gi][idx[gi].reshape(M)]) < 5e-3              # the gathered codeword (bf16 rounding only)
        p = torch.bincount(idx[gi].reshape(-1), minlength=V).double() / M
        p = p.clamp(1e-10, 1.0)
        perps.append(math.exp(-float((p * torch.log(p + 1e-10)).sum())))
    want = sum(perps) / G
    assert abs(float(out["codevector_perplexity"]) - want) < 1e-4 * want
    # -- contrastive logits ------------------------------------------------------------------------------------------------
    ps, pq = out["projected_states"].float(), out["projected_quantized_features"].float()
    logits = out["contrastive_logits"]
    pos = (ps * pq).sum(-1) / cfg.contrastive_logits_temperature
    assert rel_l2(logits[..., 0], pos) < 2e-3
    for k in (0, 37, 99):
        negq = torch.stack([pq[b, neg[b, k].long()] for b in range(B)])                       # [B, D] — the same position for every t
        want_k = (ps * negq[:, None, :]).sum(-1) / cfg.contrastive_logits_temperature
        assert rel_l2(logits[..., 1 + k], want_k) < 2e-3, k
    lse = torch.logsumexp(logits.double(), dim=-1)
    closs = float((lse - logits[..., 0].double()).mean())
    assert abs(float(out["contrastive_loss"]) - closs) < 1e-4 * abs(closs)
    assert abs(float(out["loss"]) - (closs - cfg.diversity_loss_weight * want)) < 1e-4 * abs(closs)
    # -- optimiser step with lr = 0 ----------------------------------------------------------------------------------------
    opt = Adam(learning_rate=0.0, epsilon=1e-8, clipnorm=1.0)
    p0 = model._prog.params.clone()
    W.train_step(model, (wave, None), opt, neg_indices=neg, dropout=False)
    torch.cuda.synchronize()
    assert torch.equal(model._prog.params, p0)
    assert torch.equal(model._prog.params_lp, p0.bfloat16())
    st = opt._bind(model)
    assert float(st["m"].abs().max()) > 0 and float(st["v"].abs().max()) > 0 and opt.iterations == 1
    # -- batch independence of the trunk -----------------------------------------------------------------------------------
    full = {k: out[k].float().clone() for k in ("extract_features", "last_hidden_state")}
    sub = model(wave[[1, 3]], training=True, neg_indices=neg[[1, 3]], dropout=False)
    for k, ref in full.items():
        assert rel_l2(sub[k].float(), ref[[1, 3]]) < 1e-2, k
    model._prog.ctx.watchdog()


def test_whisper_default_preset_30s_loss_is_the_ce_of_its_logits():
    from tethys_speech_b200 import whisper as WH

    model = WH.create_whisper_model("small", precision="bf16", device=0, seed=2)
    feats, labels = next(WH.create_dummy_dataset(4))
    assert tuple(feats.shape) == (4, 80, 3000) and tuple(labels.shape) == (4, 100)
    out = model(feats, labels=labels, training=True, dropout=False)
    torch.cuda.synchronize()
    V = model.config.vocab_size
    logits = out["logits"]
    assert tuple(logits.shape) == (4, 100, V) and V == 51865
    lab = torch.as_tensor(labels).to("cuda").long()
    lg = logits[:, :-1, :].float()
    ce = torch.nn.functional.cross_entropy(lg.reshape(-1, V), lab[:, 1:].reshape(-1), reduction="mean")   # pads included (W:585-600)
    assert abs(float(out["loss"]) - float(ce)) < 1e-3 * float(ce), (float(out["loss"]), float(ce))
    assert tuple(out["encoder_last_hidden_state"].shape) == (4, 1500, model.config.d_model)
    # backward at this size runs and fills every variable's gradient with finite numbers
    grads = model.gradient()
    torch.cuda.synchronize()
    assert all(bool(torch.isfinite(g).all()) for g in grads)
    assert float(dict(zip(model.variable_names, grads))["lm_head.kernel"].abs().max()) > 0
    model._prog.ctx.watchdog()
