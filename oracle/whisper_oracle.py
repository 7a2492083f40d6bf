"""TEST INFRASTRUCTURE ONLY — PyTorch-CPU restatement (autograd, any float dtype) of the reference's Whisper
encoder-decoder model and train step: speech_jobs/whisper_dist.py (W).  Structure pinned against the reference's own code on
oracle/tf_shim.py (tests/test_reference_pinning.py); TF op semantics restated — see the header of oracle/tf_ops.py.

Bug-compatible on purpose: anti-causal decoder mask with fp32 -1e9 absorption (C-1), double label
shift (C-2), un-normalised gradient SUM across replicas (C-3), query scaling after the bias (C-4), interleaved
float64 sinusoid table (C-5), untied lm_head.
"""
import math
from collections import OrderedDict

import torch

from . import tf_ops as T


# In fp64 evaluation the masked scores are given the value TF's float32 `scores + (-1e9)` would hold (absorption, App. C-1)
# while the gradient stays the identity TF's autodiff of the add uses. Tests that compare against finite differences switch
# this off: with absorption the fully-masked last decoder row is piecewise constant, so FD (= 0 there) and TF's autodiff
# (= the softmax gradient) legitimately differ.
EMULATE_FP32_ABSORPTION = True


class WhisperConfig:
    """WhisperConfig — W:10-45 with the size presets of create_whisper_model — W:852-890 ('small' = CLI default:
    d768 / 12 heads / d_ff 3072 / 4+4 layers)."""

    def __init__(self, model_type="small"):
        self.d_model, self.encoder_layers, self.decoder_layers = 768, 4, 4
        self.encoder_attention_heads = self.decoder_attention_heads = 12
        self.d_ff = 3072
        if model_type == "tiny":
            self.d_model, self.encoder_layers, self.decoder_layers, self.d_ff = 384, 4, 4, 1536
            self.encoder_attention_heads = self.decoder_attention_heads = 6
        elif model_type == "base":
            self.d_model, self.encoder_layers, self.decoder_layers, self.d_ff = 512, 6, 6, 2048
            self.encoder_attention_heads = self.decoder_attention_heads = 8
        elif model_type == "medium":
            self.d_model, self.encoder_layers, self.decoder_layers, self.d_ff = 1024, 24, 24, 4096
            self.encoder_attention_heads = self.decoder_attention_heads = 16
        elif model_type == "large":
            self.d_model, self.encoder_layers, self.decoder_layers, self.d_ff = 1280, 32, 32, 5120
            self.encoder_attention_heads = self.decoder_attention_heads = 20
        self.model_type = model_type
        self.n_mels, self.n_ctx = 80, 1500
        self.vocab_size, self.max_target_positions = 51865, 448
        self.dropout, self.attention_dropout, self.activation_dropout = 0.1, 0.1, 0.0
        self.layer_norm_eps = 1e-5
        self.pad_token_id, self.bos_token_id, self.eos_token_id = 0, 1, 2
        self.decoder_start_token_id = 50257


def init_weights(cfg, seed=0, dtype=torch.float32):
    """Keras defaults: Dense/Conv1D glorot_uniform + zero bias, LN ones/zeros, Embedding uniform(-0.05, 0.05) (A-6)."""
    g = torch.Generator().manual_seed(seed)
    w = OrderedDict()
    d, ff = cfg.d_model, cfg.d_ff

    def dense(name, i, o, bias=True):
        w[name + ".kernel"] = T.glorot_uniform(g, (i, o), i, o, dtype)
        if bias:
            w[name + ".bias"] = torch.zeros(o, dtype=dtype)

    def ln(name):
        w[name + ".gamma"] = torch.ones(d, dtype=dtype)
        w[name + ".beta"] = torch.zeros(d, dtype=dtype)

    def attn(p):
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            dense(p + n, d, d)

    w["encoder.conv1.kernel"] = T.glorot_uniform(g, (3, cfg.n_mels, d), 3 * cfg.n_mels, 3 * d, dtype)
    w["encoder.conv1.bias"] = torch.zeros(d, dtype=dtype)
    w["encoder.conv2.kernel"] = T.glorot_uniform(g, (3, d, d), 3 * d, 3 * d, dtype)
    w["encoder.conv2.bias"] = torch.zeros(d, dtype=dtype)
    for l in range(cfg.encoder_layers):
        p = f"encoder.layers.{l}."
        attn(p + "self_attn.")
        ln(p + "self_attn_layer_norm")
        dense(p + "feed_forward.fc1", d, ff)
        dense(p + "feed_forward.fc2", ff, d)
        ln(p + "final_layer_norm")
    ln("encoder.layer_norm")
    w["decoder.embed_tokens.embeddings"] = ((torch.rand((cfg.vocab_size, d), generator=g, dtype=torch.float64) * 2 - 1) * 0.05).to(dtype)
    for l in range(cfg.decoder_layers):
        p = f"decoder.layers.{l}."
        attn(p + "self_attn.")
        ln(p + "self_attn_layer_norm")
        attn(p + "encoder_attn.")
        ln(p + "encoder_attn_layer_norm")
        dense(p + "feed_forward.fc1", d, ff)
        dense(p + "feed_forward.fc2", ff, d)
        ln(p + "final_layer_norm")
    ln("decoder.layer_norm")
    dense("lm_head", d, cfg.vocab_size, bias=False)
    return w


def randomize_weights(w, seed=1, scale=0.05):
    g = torch.Generator().manual_seed(seed)
    for k, v in w.items():
        if k.endswith(".bias") or k.endswith(".beta"):
            v.copy_((torch.randn(v.shape, generator=g, dtype=torch.float64) * scale).to(v.dtype))
        elif k.endswith(".gamma"):
            v.copy_((1.0 + torch.randn(v.shape, generator=g, dtype=torch.float64) * scale).to(v.dtype))
    return w


def mha(cfg, w, p, x, kv=None, mask=None):
    """MultiHeadAttention.call — W:106-176: q = (Wq x + b) * hd^-0.5 (W:141); scores = q k^T (W:147);
    + (1 - mask) * -1e9 (W:152-154), the addition performed in float32 like TF does; softmax; @ v; out_proj."""
    B, L, d = x.shape
    nh = cfg.decoder_attention_heads
    hd = d // nh
    src = x if kv is None else kv
    Lk = src.shape[1]

    def split(t, n):
        return t.reshape(B, n, nh, hd).transpose(1, 2)

    k = split(T.dense(src, w[p + "k_proj.kernel"], w[p + "k_proj.bias"]), Lk)
    v = split(T.dense(src, w[p + "v_proj.kernel"], w[p + "v_proj.bias"]), Lk)
    q = split(T.q(T.dense(x, w[p + "q_proj.kernel"], w[p + "q_proj.bias"]) * (hd ** -0.5)), L)
    s = q @ k.transpose(-1, -2)
    if mask is not None:
        add = ((1.0 - mask) * -1e9).to(torch.float32)              # [1,L,Lk]
        # TF computes scores + mask in float32: emulate the absorption exactly, keep autograd through s
        if s.dtype == torch.float32 or not EMULATE_FP32_ABSORPTION:
            s = s + add.to(s.dtype)
        else:
            s_abs = (s.to(torch.float32) + add).to(s.dtype)        # value TF would hold at the masked entries
            s = torch.where(add != 0, s + (s_abs - s).detach(), s)
    a = T.q(torch.softmax(s, dim=-1))          # T.q: identity except under the tests' bf16-storage emulation
    ctx = T.q((a @ v).transpose(1, 2).reshape(B, L, d))
    return T.dense(ctx, w[p + "out_proj.kernel"], w[p + "out_proj.bias"])


def ffn(w, p, x):
    """FeedForward.call — W:200-206 (exact GELU; activation_dropout 0.0)."""
    h = T.gelu(T.dense(x, w[p + "fc1.kernel"], w[p + "fc1.bias"]))
    return T.dense(h, w[p + "fc2.kernel"], w[p + "fc2.bias"])


def encoder(cfg, w, feats):
    """WhisperEncoder.call — W:324-372."""
    eps = cfg.layer_norm_eps
    x = feats.transpose(1, 2)                                                     # [B,T,80]  (W:329)
    h = T.gelu(T.conv1d_same(x, w["encoder.conv1.kernel"], 1, bias=w["encoder.conv1.bias"]))
    h = T.gelu(T.conv1d_same(h, w["encoder.conv2.kernel"], 2, bias=w["encoder.conv2.bias"]))
    pe = T.sinusoid_pe(cfg.n_ctx, cfg.d_model, h.dtype)
    h = T.q(h + pe[: h.shape[1]].unsqueeze(0))
    for l in range(cfg.encoder_layers):
        p = f"encoder.layers.{l}."
        a_in = T.layer_norm(h, w[p + "self_attn_layer_norm.gamma"], w[p + "self_attn_layer_norm.beta"], eps)
        h = T.q(h + mha(cfg, w, p + "self_attn.", a_in))
        f_in = T.layer_norm(h, w[p + "final_layer_norm.gamma"], w[p + "final_layer_norm.beta"], eps)
        h = T.q(h + ffn(w, p + "feed_forward.", f_in))
    return T.layer_norm(h, w["encoder.layer_norm.gamma"], w["encoder.layer_norm.beta"], eps)


def decoder(cfg, w, ids, enc):
    """WhisperDecoder.call — W:394-466 with the anti-causal mask of W:414-418."""
    eps = cfg.layer_norm_eps
    S = ids.shape[1]
    h = w["decoder.embed_tokens.embeddings"][ids.long()]
    pe = T.sinusoid_pe(cfg.max_target_positions, cfg.d_model, h.dtype)
    h = T.q(h + pe[:S].unsqueeze(0))
    mask = (1.0 - torch.tril(torch.ones(S, S))).unsqueeze(0)                      # 1 - band_part(ones,-1,0)
    for l in range(cfg.decoder_layers):
        p = f"decoder.layers.{l}."
        x = T.layer_norm(h, w[p + "self_attn_layer_norm.gamma"], w[p + "self_attn_layer_norm.beta"], eps)
        h = T.q(h + mha(cfg, w, p + "self_attn.", x, mask=mask))
        x = T.layer_norm(h, w[p + "encoder_attn_layer_norm.gamma"], w[p + "encoder_attn_layer_norm.beta"], eps)
        h = T.q(h + mha(cfg, w, p + "encoder_attn.", x, kv=enc))
        x = T.layer_norm(h, w[p + "final_layer_norm.gamma"], w[p + "final_layer_norm.beta"], eps)
        h = T.q(h + ffn(w, p + "feed_forward.", x))
    return T.layer_norm(h, w["decoder.layer_norm.gamma"], w["decoder.layer_norm.beta"], eps)


def forward(cfg, w, feats, labels):
    """WhisperForConditionalGeneration.call(labels=…, training=True) — W:547-616."""
    B, S = labels.shape
    start = torch.full((B, 1), cfg.decoder_start_token_id, dtype=labels.dtype)
    dec_ids = torch.cat([start, labels[:, :-1]], dim=1)                            # W:559-563
    enc = encoder(cfg, w, feats)
    dec = decoder(cfg, w, dec_ids, enc)
    logits = T.q(dec @ T.qw(w["lm_head.kernel"]))                                  # W:579 (no bias, untied)
    loss = T.softmax_xent_sparse(logits[:, :-1, :], labels[:, 1:]).mean()          # W:585-600 (pads included)
    return {"loss": loss, "logits": logits, "encoder_last_hidden_state": enc, "last_hidden_state": dec}


def generate(cfg, w, feats, max_length=None):
    """WhisperForConditionalGeneration.generate — W:636-709: encoder once; each iteration runs the decoder on the whole
    prefix (no past_key_values are passed, W:664-671), takes argmax of the last position's logits (temperature and the top-k
    filter of W:676-689 do not move an argmax; tf.argmax returns the first maximum), appends it and stops once every sequence
    emitted eos_token_id in the same step (W:697-705). Returns [B, 1 + steps] int64 starting with decoder_start_token_id."""
    max_length = cfg.max_target_positions if max_length is None else max_length
    B = feats.shape[0]
    enc = encoder(cfg, w, feats)
    ids = torch.full((B, 1), cfg.decoder_start_token_id, dtype=torch.long)
    for _ in range(max_length):
        dec = decoder(cfg, w, ids, enc)
        logits = dec[:, -1, :] @ w["lm_head.kernel"]
        nxt = torch.stack([torch.nonzero(r == r.max())[0, 0] for r in logits])     # first maximum
        ids = torch.cat([ids, nxt.unsqueeze(1)], dim=1)
        if bool((nxt == cfg.eos_token_id).all()):
            break
    return ids


def loss_and_grads(cfg, w, feats, labels):
    ws = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in w.items())
    out = forward(cfg, ws, feats, labels)
    grads = torch.autograd.grad(out["loss"], list(ws.values()), allow_unused=True)
    g = OrderedDict((k, (torch.zeros_like(v) if gi is None else gi)) for (k, v), gi in zip(ws.items(), grads))
    return out, g


def train_step(cfg, w, m, v, t, feats, labels, lr=1e-4, peer_grads=None):
    """distributed_train_step — W:819-848: grads of the local mean loss, SUMMED over replicas without dividing
    (A-13 / C-3), Adam(1e-4, eps 1e-7) (W:901)."""
    out, g = loss_and_grads(cfg, w, feats, labels)
    names = list(w.keys())
    grads = [g[k] for k in names]
    if peer_grads:
        for pg in peer_grads:
            grads = [a + pg[k] for a, k in zip(grads, names)]
    T.keras_adam_step([w[k] for k in names], grads, [m[k] for k in names], [v[k] for k in names], t, lr, eps=1e-7)
    out["grads_applied"] = OrderedDict(zip(names, grads))
    return out


def dummy_labels(rng, batch, max_target_length=100):
    """create_dummy_dataset labels — W:795-809: zeros; len ~ U{50..89}; [0]=1 (BOS); [1:len-1] ~ U{3..99}; [len-1]=2."""
    import numpy as np

    lab = np.zeros((batch, max_target_length), dtype=np.int32)
    lens = rng.integers(50, 90, size=batch)
    for i in range(batch):
        n = int(lens[i])
        lab[i, 0] = 1
        lab[i, 1:n - 1] = rng.integers(3, 100, size=n - 2)
        lab[i, n - 1] = 2
    return torch.from_numpy(lab)
