"""TEST INFRASTRUCTURE ONLY — PyTorch-CPU restatement (autograd, any float dtype) of the reference's
Wav2Vec2 pre-training model and train step: speech_jobs/wav2vec2_dist.py (V), wav2vec2_single.py (VS),
whisper_single.py (WS, the legacy Wav2Vec2-base script).  Structure pinned against the reference's own code on
oracle/tf_shim.py (tests/test_reference_pinning.py); TF op semantics restated — see the header of oracle/tf_ops.py.

Weights are an ordered dict name -> tensor in Keras layouts (Dense [in,out], Conv1D [k,Cin/g,Cout]);
RNG-dependent quantities (negative indices, dropout) are explicit inputs (dropout is off: parity runs
use rate 0).
"""
import math
from collections import OrderedDict

import torch

from . import tf_ops as T


class Wav2Vec2Config:
    """Mirror of Wav2Vec2Config — V:24-128 (presets tiny/small/base) plus the extrapolated 'large'
    preset of SURVEY.md D8 (not in the reference)."""

    def __init__(self, model_size="small"):
        if model_size == "small":
            self.hidden_size, self.num_hidden_layers, self.num_attention_heads = 512, 6, 8
            self.intermediate_size = 2048
            self.conv_dim = [256] * 5
            self.conv_stride = [5, 2, 2, 2, 2]
            self.conv_kernel = [10, 3, 3, 3, 2]
            self.num_conv_pos_embeddings, self.num_conv_pos_embedding_groups = 64, 8
            self.num_codevectors_per_group, self.codevector_dim, self.proj_codevector_dim = 160, 128, 128
        elif model_size == "tiny":
            self.hidden_size, self.num_hidden_layers, self.num_attention_heads = 256, 4, 4
            self.intermediate_size = 1024
            self.conv_dim = [128] * 4
            self.conv_stride = [5, 2, 2, 2]
            self.conv_kernel = [10, 3, 3, 2]
            self.num_conv_pos_embeddings, self.num_conv_pos_embedding_groups = 32, 4
            self.num_codevectors_per_group, self.codevector_dim, self.proj_codevector_dim = 80, 64, 64
        elif model_size == "large":  # extrapolated: HF-large trunk on the base conv stack
            self.hidden_size, self.num_hidden_layers, self.num_attention_heads = 1024, 24, 16
            self.intermediate_size = 4096
            self.conv_dim = [512] * 7
            self.conv_stride = [5, 2, 2, 2, 2, 2, 2]
            self.conv_kernel = [10, 3, 3, 3, 3, 2, 2]
            self.num_conv_pos_embeddings, self.num_conv_pos_embedding_groups = 128, 16
            self.num_codevectors_per_group, self.codevector_dim, self.proj_codevector_dim = 320, 768, 768
        else:  # base
            self.hidden_size, self.num_hidden_layers, self.num_attention_heads = 768, 12, 12
            self.intermediate_size = 3072
            self.conv_dim = [512] * 7
            self.conv_stride = [5, 2, 2, 2, 2, 2, 2]
            self.conv_kernel = [10, 3, 3, 3, 3, 2, 2]
            self.num_conv_pos_embeddings, self.num_conv_pos_embedding_groups = 128, 16
            self.num_codevectors_per_group, self.codevector_dim, self.proj_codevector_dim = 320, 256, 256
        self.model_size = model_size
        self.num_codevector_groups = 2
        self.layer_norm_eps = 1e-5
        self.contrastive_logits_temperature = 0.1
        self.num_negatives = 100
        self.diversity_loss_weight = 0.1
        self.hidden_dropout = self.activation_dropout = self.attention_dropout = 0.1
        self.do_stable_layer_norm = True
        self.vocab_size = 32                                                             # V:104
        self.classifier_proj_size = {"small": 128, "tiny": 64}.get(model_size, 256)      # V:108-114
        self.num_labels = 10                                                             # VS:131


def init_weights(cfg, seed=0, dtype=torch.float32):
    """Keras default initialisers: glorot_uniform kernels, zero biases, ones/zeros norms,
    N(0,1) codebook (V:570-577)."""
    g = torch.Generator().manual_seed(seed)
    w = OrderedDict()
    cin = 1
    for i, (c, k) in enumerate(zip(cfg.conv_dim, cfg.conv_kernel)):
        w[f"fe.conv{i}.kernel"] = T.glorot_uniform(g, (k, cin, c), k * cin, k * c, dtype)
        w[f"fe.conv{i}.gn.gamma"] = torch.ones(c, dtype=dtype)
        w[f"fe.conv{i}.gn.beta"] = torch.zeros(c, dtype=dtype)
        cin = c
    C, H, G = cfg.conv_dim[-1], cfg.hidden_size, cfg.num_conv_pos_embedding_groups
    K = cfg.num_conv_pos_embeddings
    w["fe.pos_conv.kernel"] = T.glorot_uniform(g, (K, C // G, C), K * (C // G), K * C // G, dtype)
    w["fe.pos_conv.bias"] = torch.zeros(C, dtype=dtype)
    w["fe.layer_norm.gamma"] = torch.ones(C, dtype=dtype)
    w["fe.layer_norm.beta"] = torch.zeros(C, dtype=dtype)

    def dense(name, i, o):
        w[name + ".kernel"] = T.glorot_uniform(g, (i, o), i, o, dtype)
        w[name + ".bias"] = torch.zeros(o, dtype=dtype)

    def ln(name, d):
        w[name + ".gamma"] = torch.ones(d, dtype=dtype)
        w[name + ".beta"] = torch.zeros(d, dtype=dtype)

    dense("feature_projection", C, H)
    ln("feature_projection_layer_norm", H)
    for l in range(cfg.num_hidden_layers):
        p = f"encoder.layers.{l}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            dense(p + "attention." + n, H, H)
        ln(p + "attention_layer_norm", H)
        dense(p + "feed_forward.intermediate_dense", H, cfg.intermediate_size)
        dense(p + "feed_forward.output_dense", cfg.intermediate_size, H)
        ln(p + "feed_forward_layer_norm", H)
    D, P = cfg.codevector_dim, cfg.proj_codevector_dim
    w["quantizer.codevectors"] = torch.randn(
        (cfg.num_codevector_groups, cfg.num_codevectors_per_group, D // cfg.num_codevector_groups),
        generator=g, dtype=torch.float64).to(dtype)
    dense("quantizer.projection", H, D)
    dense("project_hid.dense", H, P)
    ln("project_hid.layer_norm", P)
    dense("project_q.dense", D, P)
    ln("project_q.layer_norm", P)
    return w


def randomize_weights(w, seed=1, scale=0.05):
    """Perturb biases / norm parameters away from their 0/1 init so parity tests exercise them."""
    g = torch.Generator().manual_seed(seed)
    for k, v in w.items():
        if k.endswith(".bias") or k.endswith(".beta"):
            v.copy_((torch.randn(v.shape, generator=g, dtype=torch.float64) * scale).to(v.dtype))
        elif k.endswith(".gamma"):
            v.copy_((1.0 + torch.randn(v.shape, generator=g, dtype=torch.float64) * scale).to(v.dtype))
    return w


def feature_extractor(cfg, w, wave):
    """Wav2Vec2FeatureExtractor.call — V:283-298.  GN after every conv with
    groups = num_conv_pos_embedding_groups (V:248, V:265); pos-conv inside (V:291-296)."""
    h = wave.unsqueeze(-1)
    G = cfg.num_conv_pos_embedding_groups
    for i, s in enumerate(cfg.conv_stride):
        h = T.conv1d_same(h, w[f"fe.conv{i}.kernel"], stride=s)
        h = T.group_norm(h, w[f"fe.conv{i}.gn.gamma"], w[f"fe.conv{i}.gn.beta"], G)
        h = T.gelu(h)
    pos = T.conv1d_same(h, w["fe.pos_conv.kernel"], stride=1, groups=G, bias=w["fe.pos_conv.bias"])
    h = T.layer_norm(h + pos, w["fe.layer_norm.gamma"], w["fe.layer_norm.beta"], cfg.layer_norm_eps)
    return h


def attention(cfg, w, p, x):
    """Wav2Vec2MultiHeadAttention.call — V:333-376 (scores / sqrt(hd) AFTER the matmul, V:349)."""
    B, L, H = x.shape
    nh = cfg.num_attention_heads
    hd = H // nh

    def split(t):
        return t.reshape(B, L, nh, hd).transpose(1, 2)

    q = split(T.dense(x, w[p + "q_proj.kernel"], w[p + "q_proj.bias"]))
    k = split(T.dense(x, w[p + "k_proj.kernel"], w[p + "k_proj.bias"]))
    v = split(T.dense(x, w[p + "v_proj.kernel"], w[p + "v_proj.bias"]))
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    a = T.q(torch.softmax(s, dim=-1))          # T.q: identity except under the tests' bf16-storage emulation
    ctx = T.q((a @ v).transpose(1, 2).reshape(B, L, H))
    return T.dense(ctx, w[p + "out_proj.kernel"], w[p + "out_proj.bias"])


def encoder(cfg, w, h):
    """Wav2Vec2Encoder / Wav2Vec2EncoderLayer (stable-LN path) — V:419-439, V:520-533. No final LN."""
    eps = cfg.layer_norm_eps
    for l in range(cfg.num_hidden_layers):
        p = f"encoder.layers.{l}."
        a_in = T.layer_norm(h, w[p + "attention_layer_norm.gamma"], w[p + "attention_layer_norm.beta"], eps)
        h = T.q(h + attention(cfg, w, p + "attention.", a_in))
        f_in = T.layer_norm(h, w[p + "feed_forward_layer_norm.gamma"], w[p + "feed_forward_layer_norm.beta"], eps)
        f = T.gelu(T.dense(f_in, w[p + "feed_forward.intermediate_dense.kernel"],
                           w[p + "feed_forward.intermediate_dense.bias"]))
        h = T.q(h + T.dense(f, w[p + "feed_forward.output_dense.kernel"], w[p + "feed_forward.output_dense.bias"]))
    return h


def quantizer(cfg, w, h, code_indices=None, legacy=False):
    """Wav2Vec2Quantizer.call — V:581-667: hard nearest-neighbour VQ. argmin (first minimum, int64),
    one-hot @ codebook; perplexity from the code histogram. The distance uses the literal
    sum((z-e)^2) over the group dim in the tensor's dtype, summed sequentially in index order so that
    the CUDA kernel can reproduce the fp32 rounding exactly.
    code_indices [G,B,L] int64 (tests only): replaces the argmin result. The argmin is a discontinuous integer
    decision; a reduced-precision run may legitimately pick a different near-tie code, after which every
    downstream quantity is a different (equally valid) function. Injecting the run's own indices lets the
    oracle evaluate the SAME function, so that loss and gradients stay comparable (the indices themselves are
    checked bit-exactly against an argmin on the run's own quantiser input)."""
    B, L, _ = h.shape
    G, V = cfg.num_codevector_groups, cfg.num_codevectors_per_group
    z = T.dense(h, w["quantizer.projection.kernel"], w["quantizer.projection.bias"])
    gd = cfg.codevector_dim // G
    z = z.reshape(B, L, G, gd)
    cb = w["quantizer.codevectors"]
    quantized, indices, perps = [], [], []
    for gi in range(G):
        zg = z[:, :, gi, :].detach()
        diff = zg.unsqueeze(2) - cb[gi].detach().unsqueeze(0).unsqueeze(0)   # [B,L,V,gd]
        sq = diff * diff
        dist = torch.zeros(sq.shape[:-1], dtype=sq.dtype)
        for j in range(gd):                                                 # fixed sequential order
            dist = dist + sq[..., j]
        idx = torch.argmin(dist, dim=-1)                                     # first min on ties (A-10)
        if code_indices is not None:
            idx = code_indices[gi].to(torch.long)
        onehot = torch.nn.functional.one_hot(idx, V).to(h.dtype)
        quantized.append(T.q(onehot @ cb[gi]))
        indices.append(idx)
        avg = onehot.mean(dim=(0, 1))
        if not legacy:                    # V:656 clips to [1e-10, 1]; the legacy script (WS:651-653) does not — a ~1e-9 relative
            avg = avg.clamp(1e-10, 1.0)   # difference of the loss value through the unused codes (no gradient either way)
        perps.append(torch.exp(-(avg * torch.log(avg + 1e-10)).sum()))
    q = torch.cat(quantized, dim=-1)
    return q, torch.stack(indices, 0), torch.stack(perps).mean(), z


def projection_head(w, name, x, eps):
    """Wav2Vec2ProjectionHead — V:557-561 (Dense -> LN -> dropout)."""
    return T.layer_norm(T.dense(x, w[name + ".dense.kernel"], w[name + ".dense.bias"]),
                        w[name + ".layer_norm.gamma"], w[name + ".layer_norm.beta"], eps)


def negative_indices_from_random(random_ints, num_negatives):
    """_sample_negative_indices — V:907-937 given the tf.random.uniform draw `random_ints` [B,T]:
    positions of the `actual` smallest values (top_k of the negated floats; ties: lower index first,
    A-10), tiled up to num_negatives when T-1 < num_negatives. Returns [B, num_negatives] (the
    reference then tiles the same list over every time step, V:935)."""
    B, L = random_ints.shape
    actual = max(min(num_negatives, L - 1), 1)
    vals = -random_ints.to(torch.float32)
    # stable descending sort == top_k with lower-index-first tie break
    order = torch.sort(vals, dim=1, descending=True, stable=True).indices[:, :actual]
    if actual < num_negatives:
        rep = math.ceil(num_negatives / actual)
        order = order.repeat(1, rep)[:, :num_negatives]
    return order.to(torch.int32)


def legacy_negative_indices(seq_len, perm, num_negatives):
    """Legacy sampler of whisper_single.py — WS:789-839: perm = tf.random.shuffle(range(T), seed=42)
    (injected); neg[t, k] = perm[(k - (t+1)) mod T]. Returns [T, min(num_negatives, T)] (same for all b): the reference
    slices the rolled [B, T] index row with [:, :num_negatives] (WS:821-823), so a sequence shorter than num_negatives
    yields only T negatives per step (pinned against the reference in tests/test_reference_pinning.py)."""
    t = torch.arange(seq_len).unsqueeze(1)
    k = torch.arange(min(num_negatives, seq_len)).unsqueeze(0)
    return perm[(k - (t + 1)) % seq_len].to(torch.int32)


def contrastive_loss(cfg, ps, pq, neg_idx):
    """_compute_contrastive_loss — V:865-899. neg_idx: [B,K] (shared over t, V:935) or [B,T,K]
    (legacy sampler, expanded over b by the caller). logits [B,T,1+K] / temperature; CE with label 0; mean."""
    B, L, _ = ps.shape
    temp = cfg.contrastive_logits_temperature
    pos = (ps * pq).sum(-1) / temp
    if neg_idx.dim() == 2:                                 # [B,K]: same negatives for every t
        idx = neg_idx.long().unsqueeze(1).expand(B, L, neg_idx.shape[1])
    else:                                                  # [B,T,K] (legacy sampler)
        idx = neg_idx.long()
    bi = torch.arange(B).view(B, 1, 1).expand_as(idx)
    negq = pq[bi, idx]                                   # [B,T,K,D]
    neg = (ps.unsqueeze(2) * negq).sum(-1) / temp
    logits = torch.cat([pos.unsqueeze(2), neg], dim=2)
    labels = torch.zeros(B, L, dtype=torch.long)
    loss = T.softmax_xent_sparse(logits, labels).mean()
    return logits, loss


def forward(cfg, w, wave, neg_idx, code_indices=None, legacy=False):
    """Wav2Vec2ForPreTraining.call(training=True) + the loss of the step — V:768-825, V:841-863,
    V:1202-1220. code_indices: see quantizer()."""
    out = {}
    ef = feature_extractor(cfg, w, wave)
    out["extract_features"] = ef
    hs = T.layer_norm(T.dense(ef, w["feature_projection.kernel"], w["feature_projection.bias"]),
                      w["feature_projection_layer_norm.gamma"], w["feature_projection_layer_norm.beta"],
                      cfg.layer_norm_eps)
    out["hidden_states_in"] = hs
    q, idx, perp, z = quantizer(cfg, w, hs, code_indices, legacy)
    out["quantized_features"], out["code_indices"], out["codevector_perplexity"] = q, idx, perp
    enc = encoder(cfg, w, hs)
    out["last_hidden_state"] = enc
    ps = projection_head(w, "project_hid", enc, cfg.layer_norm_eps)
    pq = projection_head(w, "project_q", q, cfg.layer_norm_eps)
    out["projected_states"], out["projected_quantized_features"] = ps, pq
    logits, closs = contrastive_loss(cfg, ps, pq, neg_idx)
    out["contrastive_logits"], out["contrastive_loss"] = logits, closs
    loss = closs + cfg.diversity_loss_weight * (-perp)        # V:1215-1220
    out["loss"] = loss
    return out


def loss_and_grads(cfg, w, wave, neg_idx, loss_div=1.0, code_indices=None, legacy=False):
    """tape.gradient(scaled_loss, trainable_variables) with None -> zeros — V:1231-1240."""
    ws = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in w.items())
    out = forward(cfg, ws, wave, neg_idx, code_indices, legacy)
    scaled = out["loss"] / loss_div
    grads = torch.autograd.grad(scaled, list(ws.values()), allow_unused=True)
    g = OrderedDict((k, (torch.zeros_like(v) if gi is None else gi)) for (k, v), gi in zip(ws.items(), grads))
    return out, g


def train_step(cfg, w, m, v, t, wave, neg_idx, lr=3e-5, eps=1e-8, legacy=False, num_replicas=1,
               peer_grads=None, code_indices=None):
    """One optimiser step, in place on (w, m, v).
    new  (V:1186-1260 / VS:1119-1176): loss/N -> grads -> clip_by_global_norm(1.0) locally ->
         [all-reduce SUM over replicas] -> per-variable clipnorm 1.0 -> Adam(lr, eps=1e-8).
    legacy (WS:1143-1180 / SV:1143-1190): plain loss, no clipping, Adam(3e-5, eps=1e-7).
    peer_grads: list of other replicas' (already locally clipped) gradient dicts, for the N>1 oracle."""
    out, g = loss_and_grads(cfg, w, wave, neg_idx, loss_div=1.0 if legacy else float(num_replicas),
                            code_indices=code_indices, legacy=legacy)
    names = list(w.keys())
    grads = [g[k] for k in names]
    if not legacy:
        grads, gnorm = T.clip_by_global_norm(grads, 1.0)
        out["global_norm"] = gnorm
    if peer_grads:
        for pg in peer_grads:
            grads = [a + pg[k] for a, k in zip(grads, names)]
    if not legacy:
        grads = T.clip_by_norm_each(grads, 1.0)
    T.keras_adam_step([w[k] for k in names], grads, [m[k] for k in names], [v[k] for k in names], t, lr,
                      eps=(1e-7 if legacy else eps))
    out["grads_applied"] = OrderedDict(zip(names, grads))
    return out


# ---- task heads on the trunk: Wav2Vec2ForCTC V:940-1001, Wav2Vec2ForSequenceClassification V:1004-1070 ----
def init_head_weights(cfg, head, seed=0, dtype=torch.float32):
    """Variables of a task-head model: the trunk's (Wav2Vec2Model incl. the quantizer, which is called in training mode;
    project_hid / project_q are never called there and so never built) + the head's Dense layers."""
    w = init_weights(cfg, seed, dtype)
    for k in [k for k in w if k.startswith("project_hid.") or k.startswith("project_q.")]:
        del w[k]
    g = torch.Generator().manual_seed(seed + 7919)
    H = cfg.hidden_size
    if head == "ctc":
        w["lm_head.kernel"] = T.glorot_uniform(g, (H, cfg.vocab_size), H, cfg.vocab_size, dtype)
        w["lm_head.bias"] = torch.zeros(cfg.vocab_size, dtype=dtype)
    elif head == "classification":
        P = cfg.classifier_proj_size
        w["classifier_proj.kernel"] = T.glorot_uniform(g, (H, P), H, P, dtype)
        w["classifier_proj.bias"] = torch.zeros(P, dtype=dtype)
        w["classifier.kernel"] = T.glorot_uniform(g, (P, cfg.num_labels), P, cfg.num_labels, dtype)
        w["classifier.bias"] = torch.zeros(cfg.num_labels, dtype=dtype)
    else:
        raise ValueError(head)
    return w


def trunk(cfg, w, wave):
    """Wav2Vec2Model.call(training=True) — V:768-825 — up to last_hidden_state (the quantizer output is unused by the heads)."""
    ef = feature_extractor(cfg, w, wave)
    hs = T.layer_norm(T.dense(ef, w["feature_projection.kernel"], w["feature_projection.bias"]),
                      w["feature_projection_layer_norm.gamma"], w["feature_projection_layer_norm.beta"], cfg.layer_norm_eps)
    return encoder(cfg, w, hs)


def forward_head(cfg, w, wave, labels, head):
    """head 'ctc': logits = lm_head(hidden) [B,T,V]; loss = mean CE of every frame against class 0 (V:994-1000).
    head 'classification': pooled = mean_t hidden (V:1043); tanh Dense; Dense; loss = mean sparse CE(labels) (V:1052-1056)."""
    h = trunk(cfg, w, wave)
    out = {"last_hidden_state": h}
    if head == "ctc":
        logits = T.dense(h, w["lm_head.kernel"], w["lm_head.bias"])
        tgt = torch.zeros(logits.shape[:2], dtype=torch.long)
        loss = T.softmax_xent_sparse(logits, tgt).mean()
    else:
        pooled = h.mean(dim=1)
        proj = torch.tanh(T.dense(pooled, w["classifier_proj.kernel"], w["classifier_proj.bias"]))
        logits = T.dense(proj, w["classifier.kernel"], w["classifier.bias"])
        loss = T.softmax_xent_sparse(logits, labels.long()).mean()
        out["pooled_output"] = pooled
    out["logits"], out["loss"] = logits, loss
    return out


def head_loss_and_grads(cfg, w, wave, labels, head, loss_div=1.0):
    """tape.gradient(loss, trainable_variables) with None -> zeros (the quantizer's variables) — VS:1163-1166."""
    ws = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in w.items())
    out = forward_head(cfg, ws, wave, labels, head)
    grads = torch.autograd.grad(out["loss"] / loss_div, list(ws.values()), allow_unused=True)
    g = OrderedDict((k, (torch.zeros_like(v) if gi is None else gi)) for (k, v), gi in zip(ws.items(), grads))
    return out, g


def head_train_step(cfg, w, m, v, t, wave, labels, head, lr=3e-5, eps=1e-8):
    """VS:1119-1176 for model_type 'asr' / 'classification': grads -> clip_by_global_norm(1.0) -> clipnorm 1.0 -> Adam."""
    out, g = head_loss_and_grads(cfg, w, wave, labels, head)
    names = list(w.keys())
    grads, gnorm = T.clip_by_global_norm([g[k] for k in names], 1.0)
    out["global_norm"] = gnorm
    grads = T.clip_by_norm_each(grads, 1.0)
    T.keras_adam_step([w[k] for k in names], grads, [m[k] for k in names], [v[k] for k in names], t, lr, eps=eps)
    return out


def ctc_loss(logits, labels, reduction="sum", zero_infinity=False):
    """Wav2Vec2ForCTC._compute_ctc_loss of the legacy file — speech_jobs/whisper_single.py:897-929: tf.nn.ctc_loss(labels [B, L] int,
    logits time-major, label_length = #(labels > 0), logit_length = T, blank_index = 0), optional inf -> 0, then "mean" / "sum" over
    the batch. tf.nn.ctc_loss is the published CTC forward algorithm (Graves et al. 2006) on log-softmax(logits); restated here with
    torch's independent implementation of the same algorithm in the dtype of `logits` (float64 in the tests), and pinned by brute-force
    enumeration of all alignments in tests/test_oracle_tf_semantics.py. Returns (loss scalar, per-sample losses)."""
    B, Tn, V = logits.shape
    lp = torch.log_softmax(logits, dim=-1).transpose(0, 1)                     # [T, B, V]: logits_time_major=True (WS:912)
    tl = (labels > 0).sum(dim=1)                                               # WS:907
    il = torch.full((B,), Tn, dtype=torch.long)                                # WS:899 (attention_mask is None on this path)
    # WS:920-921 (tf.where(is_inf(loss), 0, loss)): the replaced entries carry no gradient; torch's flag zeroes loss and gradient alike
    per = torch.nn.functional.ctc_loss(lp, labels.long(), il, tl.long(), blank=0, reduction="none", zero_infinity=bool(zero_infinity))
    return (per.mean() if reduction == "mean" else per.sum()), per


def num_frames(cfg, n_samples):
    t = n_samples
    for s in cfg.conv_stride:
        t = -(-t // s)
    return t
