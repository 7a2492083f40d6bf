"""TEST INFRASTRUCTURE ONLY — second, independent restatement (pure numpy, float64, forward only) of the two reference
models, written from the reference source without looking at oracle/*_oracle.py's torch code paths: explicit gather-based
convolutions, einsum attention, hand-rolled normalisations. tests/test_oracle_crosscheck.py requires the two restatements
to agree to ~1e-10, which is how the (otherwise unpinned) oracle is protected against transcription slips.
W = speech_jobs/whisper_dist.py, V = speech_jobs/wav2vec2_dist.py.
"""
import math

import numpy as np
from scipy.special import erf


def _same_windows(T, k, s):
    """TF SAME: T_out = ceil(T/s), pad_total = max((T_out-1)*s+k-T, 0), left = pad_total//2 (A-1).
    Returns index matrix [T_out, k] into the input (-1 = zero padding)."""
    t_out = -(-T // s)
    pad_total = max((t_out - 1) * s + k - T, 0)
    left = pad_total // 2
    idx = np.arange(t_out)[:, None] * s + np.arange(k)[None, :] - left
    idx[(idx < 0) | (idx >= T)] = -1
    return idx


def conv1d_same(x, kernel, stride=1, groups=1, bias=None):
    """x [B,T,Cin], kernel [k, Cin/groups, Cout] — gather windows, then one einsum per group."""
    B, T, Cin = x.shape
    k, cpg_in, Cout = kernel.shape
    idx = _same_windows(T, k, stride)
    xp = np.concatenate([x, np.zeros((B, 1, Cin))], axis=1)          # index -1 -> the appended zero row
    win = xp[:, idx, :]                                              # [B,T_out,k,Cin]
    out = np.zeros((B, idx.shape[0], Cout))
    cpg_out = Cout // groups
    for g in range(groups):
        wi = win[..., g * cpg_in:(g + 1) * cpg_in]
        wk = kernel[:, :, g * cpg_out:(g + 1) * cpg_out]
        out[..., g * cpg_out:(g + 1) * cpg_out] = np.einsum("btkc,kco->bto", wi, wk)
    return out if bias is None else out + bias


def gelu(x):
    return 0.5 * x * (1.0 + erf(x / math.sqrt(2.0)))


def layer_norm(x, g, b, eps=1e-5):
    mu = x.mean(-1, keepdims=True)
    var = x.var(-1, keepdims=True)          # biased
    return (x - mu) / np.sqrt(var + eps) * g + b


def group_norm(x, g, b, groups, eps=1e-5):
    """V:167-196 literally: reshape [B,T,G,C/G] -> transpose [B,T,C/G,G] -> moments over axes (1,2)."""
    B, T, C = x.shape
    r = x.reshape(B, T, groups, C // groups).transpose(0, 1, 3, 2)
    mean = r.mean(axis=(1, 2), keepdims=True)
    var = r.var(axis=(1, 2), keepdims=True)
    n = (r - mean) / np.sqrt(var + eps)
    n = n.transpose(0, 1, 3, 2).reshape(B, T, C)
    return g * n + b


def softmax(x):
    e = np.exp(x - x.max(-1, keepdims=True))
    return e / e.sum(-1, keepdims=True)


def _heads(x, nh):
    B, L, D = x.shape
    return x.reshape(B, L, nh, D // nh).transpose(0, 2, 1, 3)


# ---- Wav2Vec2 (V) ----------------------------------------------------------------------------------------------
def w2v_trunk(cfg, w, wave):
    """Wav2Vec2Model.call(training=True) up to last_hidden_state — V:768-825. Returns (extract_features, quantizer outputs, x)."""
    G = cfg.num_conv_pos_embedding_groups
    h = wave[..., None]
    for i, s in enumerate(cfg.conv_stride):                                                   # V:287-288
        h = gelu(group_norm(conv1d_same(h, w[f"fe.conv{i}.kernel"], s), w[f"fe.conv{i}.gn.gamma"], w[f"fe.conv{i}.gn.beta"], G))
    pos = conv1d_same(h, w["fe.pos_conv.kernel"], 1, groups=G, bias=w["fe.pos_conv.bias"])      # V:291
    ef = layer_norm(h + pos, w["fe.layer_norm.gamma"], w["fe.layer_norm.beta"])                 # V:294-295
    hs = layer_norm(ef @ w["feature_projection.kernel"] + w["feature_projection.bias"],
                    w["feature_projection_layer_norm.gamma"], w["feature_projection_layer_norm.beta"])   # V:777-778
    # quantiser V:581-667
    z = hs @ w["quantizer.projection.kernel"] + w["quantizer.projection.bias"]
    B, T, _ = z.shape
    Gq = cfg.num_codevector_groups
    z = z.reshape(B, T, Gq, -1)
    cb = w["quantizer.codevectors"]
    q, idxs, perps = [], [], []
    for g in range(Gq):
        dist = ((z[:, :, g, None, :] - cb[g][None, None]) ** 2).sum(-1)
        idx = dist.argmin(-1)
        onehot = np.eye(cb.shape[1])[idx]
        q.append(onehot @ cb[g])
        idxs.append(idx)
        p = np.clip(onehot.mean(axis=(0, 1)), 1e-10, 1.0)
        perps.append(np.exp(-(p * np.log(p + 1e-10)).sum()))
    qf = np.concatenate(q, -1)
    perplexity = float(np.mean(perps))
    # encoder V:419-439
    nh = cfg.num_attention_heads
    x = hs
    for l in range(cfg.num_hidden_layers):
        p = f"encoder.layers.{l}."
        a = layer_norm(x, w[p + "attention_layer_norm.gamma"], w[p + "attention_layer_norm.beta"])
        qh = _heads(a @ w[p + "attention.q_proj.kernel"] + w[p + "attention.q_proj.bias"], nh)
        kh = _heads(a @ w[p + "attention.k_proj.kernel"] + w[p + "attention.k_proj.bias"], nh)
        vh = _heads(a @ w[p + "attention.v_proj.kernel"] + w[p + "attention.v_proj.bias"], nh)
        sc = np.einsum("bhid,bhjd->bhij", qh, kh) / math.sqrt(qh.shape[-1])
        ctx = np.einsum("bhij,bhjd->bhid", softmax(sc), vh).transpose(0, 2, 1, 3).reshape(x.shape)
        x = x + ctx @ w[p + "attention.out_proj.kernel"] + w[p + "attention.out_proj.bias"]
        f = layer_norm(x, w[p + "feed_forward_layer_norm.gamma"], w[p + "feed_forward_layer_norm.beta"])
        f = gelu(f @ w[p + "feed_forward.intermediate_dense.kernel"] + w[p + "feed_forward.intermediate_dense.bias"])
        x = x + f @ w[p + "feed_forward.output_dense.kernel"] + w[p + "feed_forward.output_dense.bias"]
    return ef, (qf, np.stack(idxs, 0), perplexity), x


def _sparse_ce(logits, target):
    lse = np.log(np.exp(logits - logits.max(-1, keepdims=True)).sum(-1)) + logits.max(-1)
    return lse - np.take_along_axis(logits, np.asarray(target)[..., None].astype(np.int64), -1)[..., 0]


def w2v_head_forward(cfg, w, wave, labels, head):
    """Wav2Vec2ForCTC (V:957-1000: mean CE of every frame against class 0) / Wav2Vec2ForSequenceClassification (V:1018-1056)."""
    w = {k: np.asarray(v, dtype=np.float64) for k, v in w.items()}
    _, _, x = w2v_trunk(cfg, w, wave)
    if head == "ctc":
        logits = x @ w["lm_head.kernel"] + w["lm_head.bias"]
        loss = float(_sparse_ce(logits, np.zeros(logits.shape[:2], dtype=np.int64)).mean())
    else:
        pooled = x.mean(axis=1)
        proj = np.tanh(pooled @ w["classifier_proj.kernel"] + w["classifier_proj.bias"])
        logits = proj @ w["classifier.kernel"] + w["classifier.bias"]
        loss = float(_sparse_ce(logits, labels).mean())
    return {"last_hidden_state": x, "logits": logits, "loss": loss}


def w2v_forward(cfg, w, wave, neg_idx):
    w = {k: np.asarray(v, dtype=np.float64) for k, v in w.items()}
    ef, (qf, code_idx, perplexity), x = w2v_trunk(cfg, w, wave)
    B, T = x.shape[:2]
    ps = layer_norm(x @ w["project_hid.dense.kernel"] + w["project_hid.dense.bias"], w["project_hid.layer_norm.gamma"], w["project_hid.layer_norm.beta"])
    pq = layer_norm(qf @ w["project_q.dense.kernel"] + w["project_q.dense.bias"], w["project_q.layer_norm.gamma"], w["project_q.layer_norm.beta"])
    # contrastive V:865-899
    neg_idx = np.asarray(neg_idx)
    temp = cfg.contrastive_logits_temperature
    logits = np.zeros((B, T, 1 + neg_idx.shape[-1]))      # K = num_negatives, or T for the legacy sampler when T < num_negatives
    for b in range(B):
        for t in range(T):
            nidx = neg_idx[b] if neg_idx.ndim == 2 else neg_idx[b, t]
            logits[b, t, 0] = ps[b, t] @ pq[b, t] / temp
            logits[b, t, 1:] = pq[b, nidx] @ ps[b, t] / temp
    lse = np.log(np.exp(logits - logits.max(-1, keepdims=True)).sum(-1)) + logits.max(-1)
    closs = float((lse - logits[..., 0]).mean())
    loss = closs + cfg.diversity_loss_weight * (-perplexity)
    return {"extract_features": ef, "last_hidden_state": x, "quantized_features": qf, "code_indices": code_idx,
            "codevector_perplexity": perplexity, "projected_states": ps, "projected_quantized_features": pq,
            "contrastive_logits": logits, "contrastive_loss": closs, "loss": loss}


# ---- Whisper (W) -----------------------------------------------------------------------------------------------
def _pe(max_len, d):
    pe = np.zeros((max_len, d))
    position = np.arange(0, max_len)[:, np.newaxis]
    div_term = np.exp(np.arange(0, d, 2) * -(np.log(10000.0) / d))
    pe[:, 0::2] = np.sin(position * div_term)
    pe[:, 1::2] = np.cos(position * div_term)
    return pe.astype(np.float32).astype(np.float64)      # the table is stored in fp32 (W:63)


def _mha(w, p, x, nh, kv=None, mask=None):
    src = x if kv is None else kv
    hd = x.shape[-1] // nh
    k = _heads(src @ w[p + "k_proj.kernel"] + w[p + "k_proj.bias"], nh)
    v = _heads(src @ w[p + "v_proj.kernel"] + w[p + "v_proj.bias"], nh)
    q = _heads((x @ w[p + "q_proj.kernel"] + w[p + "q_proj.bias"]) * hd ** -0.5, nh)          # W:141
    s = np.einsum("bhid,bhjd->bhij", q, k)
    if mask is not None:
        add = ((1.0 - mask) * -1e9).astype(np.float32)
        absorbed = (s.astype(np.float32) + add).astype(np.float64)                             # fp32 add (App. C-1)
        s = np.where(add != 0, absorbed, s)
    ctx = np.einsum("bhij,bhjd->bhid", softmax(s), v).transpose(0, 2, 1, 3).reshape(x.shape)
    return ctx @ w[p + "out_proj.kernel"] + w[p + "out_proj.bias"]


def _whisper_encoder(cfg, w, feats):
    nh = cfg.encoder_attention_heads
    x = feats.transpose(0, 2, 1)
    h = gelu(conv1d_same(x, w["encoder.conv1.kernel"], 1, bias=w["encoder.conv1.bias"]))
    h = gelu(conv1d_same(h, w["encoder.conv2.kernel"], 2, bias=w["encoder.conv2.bias"]))
    h = h + _pe(cfg.n_ctx, cfg.d_model)[None, :h.shape[1]]
    for l in range(cfg.encoder_layers):
        p = f"encoder.layers.{l}."
        h = h + _mha(w, p + "self_attn.", layer_norm(h, w[p + "self_attn_layer_norm.gamma"], w[p + "self_attn_layer_norm.beta"]), nh)
        t = layer_norm(h, w[p + "final_layer_norm.gamma"], w[p + "final_layer_norm.beta"])
        h = h + gelu(t @ w[p + "feed_forward.fc1.kernel"] + w[p + "feed_forward.fc1.bias"]) @ w[p + "feed_forward.fc2.kernel"] + w[p + "feed_forward.fc2.bias"]
    return layer_norm(h, w["encoder.layer_norm.gamma"], w["encoder.layer_norm.beta"])


def _whisper_decoder(cfg, w, ids, enc):
    """WhisperDecoder.call with the anti-causal mask 1 - band_part(ones, -1, 0) of W:414-418."""
    nh = cfg.decoder_attention_heads
    S = ids.shape[1]
    g = w["decoder.embed_tokens.embeddings"][ids] + _pe(cfg.max_target_positions, cfg.d_model)[None, :S]
    mask = (1.0 - np.tril(np.ones((S, S))))[None, None]
    for l in range(cfg.decoder_layers):
        p = f"decoder.layers.{l}."
        g = g + _mha(w, p + "self_attn.", layer_norm(g, w[p + "self_attn_layer_norm.gamma"], w[p + "self_attn_layer_norm.beta"]), nh, mask=mask)
        g = g + _mha(w, p + "encoder_attn.", layer_norm(g, w[p + "encoder_attn_layer_norm.gamma"], w[p + "encoder_attn_layer_norm.beta"]), nh, kv=enc)
        t = layer_norm(g, w[p + "final_layer_norm.gamma"], w[p + "final_layer_norm.beta"])
        g = g + gelu(t @ w[p + "feed_forward.fc1.kernel"] + w[p + "feed_forward.fc1.bias"]) @ w[p + "feed_forward.fc2.kernel"] + w[p + "feed_forward.fc2.bias"]
    return layer_norm(g, w["decoder.layer_norm.gamma"], w["decoder.layer_norm.beta"])


def whisper_generate(cfg, w, feats, max_length):
    """generate() — W:636-709: greedy, the whole prefix re-decoded every step, stop when every sequence emits EOS."""
    w = {k: np.asarray(v, dtype=np.float64) for k, v in w.items()}
    enc = _whisper_encoder(cfg, w, feats)
    ids = np.full((feats.shape[0], 1), cfg.decoder_start_token_id, dtype=np.int64)
    for _ in range(max_length):
        logits = _whisper_decoder(cfg, w, ids, enc)[:, -1, :] @ w["lm_head.kernel"]
        nxt = logits.argmax(-1)                                  # numpy argmax: first maximum, like tf.argmax
        ids = np.concatenate([ids, nxt[:, None]], axis=1)
        if (nxt == cfg.eos_token_id).all():
            break
    return ids


def whisper_forward(cfg, w, feats, labels):
    w = {k: np.asarray(v, dtype=np.float64) for k, v in w.items()}
    labels = np.asarray(labels)
    enc = _whisper_encoder(cfg, w, feats)
    B, S = labels.shape
    ids = np.concatenate([np.full((B, 1), cfg.decoder_start_token_id), labels[:, :-1]], axis=1)
    dec = _whisper_decoder(cfg, w, ids, enc)
    logits = dec @ w["lm_head.kernel"]
    sl = logits[:, :-1]
    tgt = labels[:, 1:]
    lse = np.log(np.exp(sl - sl.max(-1, keepdims=True)).sum(-1)) + sl.max(-1)
    picked = np.take_along_axis(sl, tgt[..., None].astype(np.int64), -1)[..., 0]
    return {"loss": float((lse - picked).mean()), "logits": logits, "encoder_last_hidden_state": enc, "last_hidden_state": dec}
