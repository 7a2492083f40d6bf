"""Host-side mirror of the reference's Wav2Vec2 surface (speech_jobs/wav2vec2_dist.py = V, wav2vec2_single.py = VS,
whisper_single.py = WS): same class names, call signatures, output dict keys and step functions, driving the
native program in libtethys.so (csrc/w2v_program.cu) through ctypes. No TensorFlow, no CPU fallback.
"""
import ctypes as C
import os
import math

import numpy as np
import torch

from . import _lib
from .runtime import (Adam, GradientList, ProgramBase, ReduceOp, Strategy, clip_by_global_norm, get_strategy, ptr, stream_ptr,
                      to_device, view_from_ptr)


class Wav2Vec2Config:
    """Same attributes and presets as Wav2Vec2Config — V:24-128 (tiny / small (default) / base), plus the
    extrapolated 'large' preset of SURVEY.md D8 which the reference does not have."""

    def __init__(self, model_size="small"):
        if model_size == "small":
            self.hidden_size, self.num_hidden_layers, self.num_attention_heads, self.intermediate_size = 512, 6, 8, 2048
            self.conv_dim = [256] * 5
            self.conv_stride = [5, 2, 2, 2, 2]
            self.conv_kernel = [10, 3, 3, 3, 2]
            self.num_conv_pos_embeddings, self.num_conv_pos_embedding_groups = 64, 8
            self.num_codevectors_per_group, self.codevector_dim, self.proj_codevector_dim = 160, 128, 128
            self.classifier_proj_size = 128
        elif model_size == "tiny":
            self.hidden_size, self.num_hidden_layers, self.num_attention_heads, self.intermediate_size = 256, 4, 4, 1024
            self.conv_dim = [128] * 4
            self.conv_stride = [5, 2, 2, 2]
            self.conv_kernel = [10, 3, 3, 2]
            self.num_conv_pos_embeddings, self.num_conv_pos_embedding_groups = 32, 4
            self.num_codevectors_per_group, self.codevector_dim, self.proj_codevector_dim = 80, 64, 64
            self.classifier_proj_size = 64
        elif model_size == "large":
            self.hidden_size, self.num_hidden_layers, self.num_attention_heads, self.intermediate_size = 1024, 24, 16, 4096
            self.conv_dim = [512] * 7
            self.conv_stride = [5, 2, 2, 2, 2, 2, 2]
            self.conv_kernel = [10, 3, 3, 3, 3, 2, 2]
            self.num_conv_pos_embeddings, self.num_conv_pos_embedding_groups = 128, 16
            self.num_codevectors_per_group, self.codevector_dim, self.proj_codevector_dim = 320, 768, 768
            self.classifier_proj_size = 256
        else:  # base
            self.hidden_size, self.num_hidden_layers, self.num_attention_heads, self.intermediate_size = 768, 12, 12, 3072
            self.conv_dim = [512] * 7
            self.conv_stride = [5, 2, 2, 2, 2, 2, 2]
            self.conv_kernel = [10, 3, 3, 3, 3, 2, 2]
            self.num_conv_pos_embeddings, self.num_conv_pos_embedding_groups = 128, 16
            self.num_codevectors_per_group, self.codevector_dim, self.proj_codevector_dim = 320, 256, 256
            self.classifier_proj_size = 256
        self.model_size = model_size
        self.feat_extract_norm = "group"
        self.feat_extract_activation = "gelu"
        self.conv_bias = False
        self.hidden_act = "gelu"
        self.hidden_dropout = 0.1
        self.activation_dropout = 0.1
        self.attention_dropout = 0.1
        self.layer_norm_eps = 1e-5
        self.num_codevector_groups = 2
        self.contrastive_logits_temperature = 0.1
        self.num_negatives = 100
        self.diversity_loss_weight = 0.1
        self.mask_time_prob = 0.05      # set but never read by the reference
        self.mask_time_length = 10
        self.mask_feature_prob = 0.0
        self.mask_feature_length = 10
        self.vocab_size = 32
        self.do_stable_layer_norm = True
        self.use_weighted_layer_sum = False
        self.ctc_loss_reduction = "sum"   # read into Wav2Vec2ForCTC but unused by its stand-in loss (V:954-955, V:994-1000)
        self.ctc_zero_infinity = False
        self.num_labels = 10              # VS:131 (wav2vec2_dist.py's config lacks it: its classification model cannot be built)


# Keras trainable_variables order of Wav2Vec2ForPreTraining (attribute-tracking order, V:746-766, V:229-281)
def _keras_order(cfg, head="pretraining"):
    """head: 'pretraining' | 'ctc' | 'classification'. The task heads wrap the same Wav2Vec2Model; its project_hid / project_q
    layers are never called there, so Keras never creates their variables; the quantizer is called (training=True) and its
    variables exist with None -> zero gradients (VS:1163-1166). The head's own layers come last (V:950, V:1013-1015)."""
    names = []
    for i in range(len(cfg.conv_dim)):
        names += [f"fe.conv{i}.kernel", f"fe.conv{i}.gn.gamma", f"fe.conv{i}.gn.beta"]
    names += ["fe.pos_conv.kernel", "fe.pos_conv.bias", "fe.layer_norm.gamma", "fe.layer_norm.beta",
              "feature_projection.kernel", "feature_projection.bias",
              "feature_projection_layer_norm.gamma", "feature_projection_layer_norm.beta"]
    for l in range(cfg.num_hidden_layers):
        p = f"encoder.layers.{l}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            names += [p + f"attention.{n}.kernel", p + f"attention.{n}.bias"]
        names += [p + "attention_layer_norm.gamma", p + "attention_layer_norm.beta",
                  p + "feed_forward.intermediate_dense.kernel", p + "feed_forward.intermediate_dense.bias",
                  p + "feed_forward.output_dense.kernel", p + "feed_forward.output_dense.bias",
                  p + "feed_forward_layer_norm.gamma", p + "feed_forward_layer_norm.beta"]
    names += ["quantizer.codevectors", "quantizer.projection.kernel", "quantizer.projection.bias"]
    if head == "pretraining":
        names += ["project_hid.dense.kernel", "project_hid.dense.bias", "project_hid.layer_norm.gamma", "project_hid.layer_norm.beta",
                  "project_q.dense.kernel", "project_q.dense.bias", "project_q.layer_norm.gamma", "project_q.layer_norm.beta"]
    elif head == "ctc":
        names += ["lm_head.kernel", "lm_head.bias"]
    else:
        names += ["classifier_proj.kernel", "classifier_proj.bias", "classifier.kernel", "classifier.bias"]
    return names


_HEADS = {"pretraining": 0, "ctc": 1, "classification": 2}


class _Program(ProgramBase):
    """Arenas + the ts_w2v handle."""

    def __init__(self, cfg, precision, device, head="pretraining"):
        c = _lib.W2VConfig()
        c.head = _HEADS[head]
        c.vocab_size, c.classifier_proj, c.num_labels = cfg.vocab_size, cfg.classifier_proj_size, cfg.num_labels
        c.hidden, c.layers, c.heads, c.ffn = cfg.hidden_size, cfg.num_hidden_layers, cfg.num_attention_heads, cfg.intermediate_size
        c.n_conv = len(cfg.conv_dim)
        for i in range(c.n_conv):
            c.conv_dim[i], c.conv_kernel[i], c.conv_stride[i] = cfg.conv_dim[i], cfg.conv_kernel[i], cfg.conv_stride[i]
        c.pos_kernel, c.pos_groups = cfg.num_conv_pos_embeddings, cfg.num_conv_pos_embedding_groups
        c.cv_groups, c.cv_per_group = cfg.num_codevector_groups, cfg.num_codevectors_per_group
        c.cv_dim, c.proj_dim = cfg.codevector_dim, cfg.proj_codevector_dim
        c.num_negatives = cfg.num_negatives
        c.ln_eps, c.temperature, c.diversity_weight = cfg.layer_norm_eps, cfg.contrastive_logits_temperature, cfg.diversity_loss_weight
        c.hidden_dropout, c.activation_dropout, c.attention_dropout = cfg.hidden_dropout, cfg.activation_dropout, cfg.attention_dropout
        self.ccfg = c
        lib = _lib.load()
        super().__init__("ts_w2v", precision, device, lambda ctx_h, prec, out: lib.ts_w2v_create(ctx_h, C.byref(c), prec, out))


def _glorot_uniform(gen, shape, fan_in, fan_out, device):
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen, dtype=torch.float32, device=device) * 2 - 1) * limit


class Wav2Vec2Model:
    """Handle mirroring `Wav2Vec2Model` (V:746-825): exposes the sub-layer names the reference code touches."""

    def __init__(self, owner):
        self._owner = owner
        self.config = owner.config

    def project_hid(self, hidden_states, training=False):
        return self._owner._cached("projected_states", hidden_states)

    def project_q(self, hidden_states, training=False):
        return self._owner._cached("projected_quantized_features", hidden_states)


class _Wav2Vec2Task:
    """What the three task models share: the native program + arenas, Keras-ordered variables, weight I/O, gradients."""

    _head = "pretraining"

    def __init__(self, config, precision="bf16", device=None, seed=0):
        self.config = config
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self._prog = _Program(config, precision, device, self._head)
        self.wav2vec2 = Wav2Vec2Model(self)
        self._rng = torch.Generator(device=self._prog.device)
        self._rng.manual_seed(1234 + seed)
        self._step_seed = seed * 1000003
        self._last = {}
        self._init_weights(seed)
        names = _keras_order(config, self._head)
        assert set(names) == set(self._prog.info), "parameter table mismatch"
        self.variable_names = names
        self.trainable_variables = [self._prog.view(self._prog.params, n) for n in names]

    # -- weights ------------------------------------------------------------------------------------------
    def _init_weights(self, seed):
        """Keras defaults: glorot_uniform kernels, zero biases, ones/zeros norms, N(0,1) codebook (V:570-577)."""
        p = self._prog
        gen = torch.Generator(device=p.device)
        gen.manual_seed(seed)
        for name, (off, shp, ld) in p.info.items():
            v = p.view(p.params, name)
            if name.endswith(".kernel"):
                if len(shp) == 3:
                    k, cin, cout = shp
                    v.copy_(_glorot_uniform(gen, shp, k * cin, k * cout, p.device))
                else:
                    v.copy_(_glorot_uniform(gen, shp, shp[0], shp[1], p.device))
            elif name.endswith(".gamma"):
                v.fill_(1.0)
            elif name == "quantizer.codevectors":
                v.copy_(torch.randn(shp, generator=gen, dtype=torch.float32, device=p.device))
            else:
                v.zero_()
        p.weights_synced = False

    def set_weights(self, weights):
        """Load a {name: array} dict (Keras layouts, names as in `variable_names`)."""
        p = self._prog
        for k, w in weights.items():
            p.view(p.params, k).copy_(to_device(w, torch.float32, p.device).view(p.info[k][1]))
        p.weights_synced = False

    def get_weights(self):
        p = self._prog
        return {k: p.view(p.params, k).detach().clone() for k in self.variable_names}

    def broadcast_weights(self, strategy):
        self._prog.use_comm_buffers(strategy)      # gradient arenas into communicator-registered memory (no-op for one replica)
        strategy.broadcast_(self._prog.params)
        self._prog.weights_synced = False

    # -- shared ---------------------------------------------------------------------------------------------
    def num_frames(self, n_samples):
        t = int(n_samples)
        for s in self.config.conv_stride:
            t = -(-t // s)
        return t

    def extract_features(self, inputs):
        """Wav2Vec2FeatureExtractor.call (V:283-298) only: conv stack + GroupNorm/GELU + positional conv + LayerNorm in
        inference mode. Returns the [B, T, C] feature tensor (a view of the workspace)."""
        p = self._prog
        x = to_device(inputs, torch.float32, p.device)
        B, N = x.shape
        p.ensure_workspace(B, N)
        p.sync_weights()
        p.ctx.check(p.lib.ts_w2v_forward_features(p.h, ptr(x), B, N, stream_ptr()))
        self._last = {"x": x}
        return p.buffer("extract_features")

    def gradient(self, stage_from=0, stage_to=10 ** 6):
        """tape.gradient(loss, model.trainable_variables) (V:1234) with None→zeros (V:1237-1240): fills the
        gradient arena and returns views in `trainable_variables` order."""
        p = self._prog
        p.backward(stage_from, stage_to)
        gl = GradientList(p.view(p.grads, n) for n in self.variable_names)
        gl.owner = self
        for g in gl:
            g._ts_owner = self
        return gl

    def save_weights(self, path):
        """Keras `model.save_weights` (W:1025): variables only, in the checkpoint.py container."""
        from . import checkpoint
        return checkpoint.save(path, self)

    def load_weights(self, path, strict=True):
        """The restore the reference never calls (SURVEY f-3): variables from a file written by save_weights / Checkpoint."""
        from . import checkpoint
        return checkpoint.restore(path, self, strict=strict)


class Wav2Vec2ForPreTraining(_Wav2Vec2Task):
    """Mirror of `Wav2Vec2ForPreTraining` — V:828-937. `model(inputs, training=True)` returns the same dict keys;
    `_compute_contrastive_loss` / `_compute_diversity_loss` / `_sample_negative_indices` keep their signatures."""

    _head = "pretraining"

    def __init__(self, config, precision="bf16", device=None, seed=0):
        super().__init__(config, precision, device, seed)
        self.num_negatives = config.num_negatives
        self.contrastive_logits_temperature = config.contrastive_logits_temperature
        self.diversity_loss_weight = config.diversity_loss_weight

    # -- forward ------------------------------------------------------------------------------------------
    def _sample_negative_indices(self, sequence_length, batch_size):
        """V:907-937: per batch row, the positions of the `actual` smallest of T uniform ints (ties: lower index
        first), tiled to num_negatives, then the same list for every time step. Returns int32 [B, T, K]."""
        T, K = int(sequence_length), self.num_negatives
        actual = max(min(K, T - 1), 1)
        p = self._prog
        r = torch.randint(0, T, (int(batch_size), T), generator=self._rng, device=p.device, dtype=torch.int32)
        order = torch.empty(int(batch_size), K, dtype=torch.int32, device=p.device)
        p.ctx.check(p.lib.ts_w2v_sample_negatives(p.ctx.h, ptr(r), int(batch_size), T, K, ptr(order), stream_ptr()))   # top_k(-r), tiled
        return order.unsqueeze(1).expand(-1, T, -1)

    def __call__(self, inputs, attention_mask=None, output_attentions=False, output_hidden_states=False, training=False,
                 neg_indices=None, loss_div=1.0, dropout=True):
        """inputs: [B, N] waveform (torch / numpy / DLPack producer). With training=True the quantiser, the projection
        heads and the loss run too (V:782-789, V:852-861). `neg_indices` ([B,K] or [B,T,K] int32) injects the negative
        sample positions (parity tests); otherwise they are drawn like V:907-937. `dropout=False` disables the dropout
        layers (parity runs)."""
        if attention_mask is not None or output_attentions or output_hidden_states:
            raise NotImplementedError("attention_mask / output_attentions / output_hidden_states are not on the train path")
        p = self._prog
        x = to_device(inputs, torch.float32, p.device)
        if x.dim() != 2:
            raise ValueError("inputs must be [batch, samples]")
        B, N = x.shape
        p.ensure_workspace(B, N)
        p.sync_weights()
        T = self.num_frames(N)
        if neg_indices is None:
            neg = self._sample_negative_indices(T, B)[:, 0, :].contiguous()
        else:
            neg = to_device(neg_indices, torch.int32, p.device)
        if neg.dim() == 2:
            neg_bs, neg_ts = neg.shape[1], 0
        else:
            neg = neg.contiguous()
            neg_bs, neg_ts = neg.shape[1] * neg.shape[2], neg.shape[2]
        if neg.shape[-1] != self.num_negatives:
            raise ValueError(f"neg_indices last dim must be num_negatives={self.num_negatives}")
        self._step_seed += 1
        p.ctx.check(p.lib.ts_w2v_forward(p.h, ptr(x), B, N, ptr(neg), neg_bs, neg_ts, float(loss_div), self._step_seed,
                                         1 if (training and dropout) else 0, stream_ptr()))
        self._last = {"x": x, "neg": neg, "training": training}
        scal = p.buffer("scalars")
        out = {"last_hidden_state": p.buffer("last_hidden_state"), "extract_features": p.buffer("extract_features")}
        if training:
            out["quantized_features"] = p.buffer("quantized_features")
            out["codevector_perplexity"] = scal[2]
            out["projected_quantized_features"] = p.buffer("projected_quantized_features")
            out["projected_states"] = p.buffer("projected_states")
            out["code_indices"] = p.buffer("code_indices")
            out["loss"] = scal[0]
            out["contrastive_loss"] = scal[1]
            out["contrastive_logits"] = p.buffer("contrastive_logits")
        self._last["out"] = out
        return out

    call = __call__

    def _cached(self, key, _arg):
        return self._last["out"][key]

    def _compute_contrastive_loss(self, hidden_states, quantized_states):
        """V:865-899 → (logits [B,T,1+K], mean CE). Computed inside the fused forward of the last call."""
        o = self._last["out"]
        return o["contrastive_logits"], o["contrastive_loss"]

    def _compute_diversity_loss(self, perplexity):
        """V:901-905."""
        return -perplexity


class _Wav2Vec2HeadModel(_Wav2Vec2Task):
    """Shared call path of the two fine-tuning heads (ts_w2v_forward_head)."""

    def __call__(self, inputs, attention_mask=None, labels=None, output_attentions=False, output_hidden_states=False,
                 return_dict=True, training=False, loss_div=1.0, dropout=True):
        if attention_mask is not None or output_attentions or output_hidden_states or not return_dict:
            raise NotImplementedError("attention_mask / output_attentions / output_hidden_states / tuple outputs are not on the train path")
        p = self._prog
        x = to_device(inputs, torch.float32, p.device)
        if x.dim() != 2:
            raise ValueError("inputs must be [batch, samples]")
        B, N = x.shape
        p.ensure_workspace(B, N)
        p.sync_weights()
        with_loss = bool(training and labels is not None)          # V:982-985 / V:1052-1056: loss only when training with labels
        lab = None
        if with_loss and self._head == "classification":
            lab = to_device(labels, torch.int32, p.device).reshape(-1)
            if lab.numel() != B:
                raise ValueError(f"labels must hold one class id per clip ({B}), got {tuple(lab.shape)}")
        self._step_seed += 1
        p.ctx.check(p.lib.ts_w2v_forward_head(p.h, ptr(x), B, N, ptr(lab), float(loss_div), self._step_seed,
                                              1 if (training and dropout) else 0, 1 if with_loss else 0, stream_ptr()))
        self._last = {"x": x, "labels": lab, "training": training}
        out = {"loss": p.buffer("scalars")[0] if with_loss else None, "logits": p.buffer("head_logits"),
               "hidden_states": None, "attentions": None}
        self._last["out"] = out
        return out

    call = __call__


class Wav2Vec2ForCTC(_Wav2Vec2HeadModel):
    """Mirror of `Wav2Vec2ForCTC` — V:940-1001: dropout + lm_head(vocab_size) on the trunk. Its `_compute_ctc_loss` is the
    reference's stand-in, the mean sparse cross-entropy of every frame against class 0 (V:994-1000), not tf.nn.ctc_loss."""

    _head = "ctc"

    def __init__(self, config, precision="bf16", device=None, seed=0):
        super().__init__(config, precision, device, seed)
        self.ctc_loss_reduction = config.ctc_loss_reduction
        self.ctc_zero_infinity = config.ctc_zero_infinity

    def __call__(self, inputs, attention_mask=None, labels=None, output_attentions=False, output_hidden_states=False,
                 return_dict=True, training=False, loss_div=1.0, dropout=True):
        """labels None / per-clip scalars (the dummy dataset's, V:1139): the reference's stand-in loss (V:994-1000).
        labels [B, L] int (a transcript, 0 = padding): the real CTC loss of the legacy file's Wav2Vec2ForCTC (WS:897-929) —
        tf.nn.ctc_loss with blank 0, reduced with config.ctc_loss_reduction; its gradient replaces the stand-in's in the program."""
        lab2d = None
        if labels is not None:
            t = labels if isinstance(labels, torch.Tensor) else torch.as_tensor(np.asarray(labels))
            if t.dim() == 2 and t.shape[1] > 1:
                lab2d = t
        out = super().__call__(inputs, attention_mask, labels if lab2d is None else torch.zeros(lab2d.shape[0]), output_attentions,
                               output_hidden_states, return_dict, training, loss_div, dropout)
        if lab2d is None or not training:
            return out
        p = self._prog
        logits = out["logits"]                                         # fp32 [B, T, vocab]
        B, T, V = logits.shape
        lab = to_device(lab2d, torch.int32, p.device).contiguous()
        L = lab.shape[1]
        ws = torch.empty(int(p.lib.ts_ctc_workspace_floats(B, T, L)), dtype=torch.float32, device=p.device)
        per = torch.empty(B, dtype=torch.float32, device=p.device)
        scale = (1.0 / B if self.ctc_loss_reduction == "mean" else 1.0) / float(loss_div)
        p.ctx.check(p.lib.ts_ctc_loss(p.ctx.h, p.precision, ptr(logits), ptr(lab), B, T, V, L, 0, ptr(ws), ptr(per), ptr(p.buffer("d_head_logits")),
                                      scale, 1 if self.ctc_zero_infinity else 0, stream_ptr()))
        out["loss"] = per.mean() if self.ctc_loss_reduction == "mean" else per.sum()     # unscaled, like every model's "loss" (the step divides)
        out["ctc_loss_per_sample"] = per
        return out

    call = __call__


class Wav2Vec2ForSequenceClassification(_Wav2Vec2HeadModel):
    """Mirror of `Wav2Vec2ForSequenceClassification` — V:1004-1070: mean over time, Dense(classifier_proj_size, tanh),
    dropout, Dense(num_labels), mean sparse cross-entropy with the integer labels."""

    _head = "classification"


def create_full_model(model_type="pretraining", model_size="small", num_negatives=100, mask_time_prob=0.065,
                      mask_time_length=10, precision="bf16", device=None, seed=0):
    """V:1157-1182 / VS:1094-1115: 'pretraining' -> Wav2Vec2ForPreTraining, 'asr' -> Wav2Vec2ForCTC, 'classification' ->
    Wav2Vec2ForSequenceClassification. (The reference's fall-through, a bare Wav2Vec2Model, has no loss and cannot be trained.)"""
    config = Wav2Vec2Config(model_size=model_size)
    config.num_negatives = num_negatives
    config.mask_time_prob = mask_time_prob
    config.mask_time_length = mask_time_length
    if model_type == "pretraining":
        return Wav2Vec2ForPreTraining(config, precision=precision, device=device, seed=seed)
    if model_type == "asr":
        return Wav2Vec2ForCTC(config, precision=precision, device=device, seed=seed)
    if model_type == "classification":
        return Wav2Vec2ForSequenceClassification(config, precision=precision, device=device, seed=seed)
    raise NotImplementedError(f"model_type={model_type!r}: a bare Wav2Vec2Model has no loss to train on")


def create_dummy_dataset(batch_size, audio_length=32000, num_samples=50, seed=1234, device=None):
    """V:1123-1153: 50 × N(0,1) waveforms of `audio_length` samples (+ scalar 0 label), batched with
    drop_remainder and repeated. Seeded (the reference is not). Yields (features [B,N] pinned host fp32, labels)."""
    rng = np.random.default_rng(seed)
    data = torch.from_numpy(rng.standard_normal((num_samples, audio_length), dtype=np.float32))
    if torch.cuda.is_available():
        data = data.pin_memory()
    labels = torch.zeros(num_samples)

    def gen():
        nb = num_samples // batch_size
        if nb == 0:
            raise ValueError("batch_size larger than the dataset")
        while True:
            for i in range(nb):
                yield data[i * batch_size:(i + 1) * batch_size], labels[i * batch_size:(i + 1) * batch_size]

    return gen()


def train_step(model, inputs, optimizer, neg_indices=None, dropout=True):
    """Single-device step — VS:1119-1176: forward, loss (pre-training: contrastive + diversity; task heads: the model's own
    "loss" for (features, labels), VS:1155-1157), NaN guard, gradients (None→0), clip_by_global_norm(1.0),
    optimizer.apply_gradients (clipnorm + Adam). Returns the loss (device scalar)."""
    features, labels = inputs
    if features.shape[0] == 0:
        return torch.zeros(())
    if isinstance(model, Wav2Vec2ForPreTraining):
        outputs = model(features, training=True, neg_indices=neg_indices, dropout=dropout)
        logits, contrastive_loss = model._compute_contrastive_loss(outputs["projected_states"], outputs["projected_quantized_features"])
    else:
        outputs = model(features, labels=labels, training=True, dropout=dropout)
    loss = outputs["loss"]  # NaN→0 (V:1220-1228 / VS:1160) fused on the device
    gradients = model.gradient()
    optimizer.apply_gradients(gradients, global_clip_norm=1.0)
    return loss


def native_train_step(model, inputs, optimizer, neg_indices=None, dropout=True, strategy=None):
    """train_step (VS:1119-1176; strategy None / one replica) or distributed_train_step (V:1186-1260) as ONE C-ABI call —
    ts_w2v_step, the composite entry of SURVEY §8 b-2: forward (loss / N), backward, local clip_by_global_norm(1.0), all-reduce
    SUM, clipnorm + Adam, and the reduced loss, all enqueued on the current stream. Same arithmetic and the same kernels as the
    Python-composed step functions above; returns the step's loss (device scalar, valid until the next native step)."""
    from .runtime import native_step_args

    features, labels = inputs
    p = model._prog
    x = to_device(features, torch.float32, p.device)
    if x.dim() != 2:
        raise ValueError("inputs must be [batch, samples]")
    B, N = x.shape
    if B == 0:
        return torch.zeros((), device=p.device)
    p.ensure_workspace(B, N)
    p.sync_weights()
    neg, neg_bs, neg_ts, lab = None, 0, 0, None
    if isinstance(model, Wav2Vec2ForPreTraining):
        T = model.num_frames(N)
        if neg_indices is None:
            neg = model._sample_negative_indices(T, B)[:, 0, :].contiguous()
        else:
            neg = to_device(neg_indices, torch.int32, p.device)
        if neg.dim() == 2:
            neg_bs, neg_ts = neg.shape[1], 0
        else:
            neg = neg.contiguous()
            neg_bs, neg_ts = neg.shape[1] * neg.shape[2], neg.shape[2]
        if neg.shape[-1] != model.num_negatives:
            raise ValueError(f"neg_indices last dim must be num_negatives={model.num_negatives}")
    elif model._head == "classification" and labels is not None:
        lab = to_device(labels, torch.int32, p.device).reshape(-1)
        if lab.numel() != B:
            raise ValueError(f"labels must hold one class id per clip ({B}), got {tuple(lab.shape)}")
    elif model._head == "ctc" and labels is not None:
        t = labels if isinstance(labels, torch.Tensor) else torch.as_tensor(np.asarray(labels))
        if t.dim() == 2 and t.shape[1] > 1:
            # transcripts select the real tf.nn.ctc_loss of the legacy file (WS:897-929), which Wav2Vec2ForCTC.__call__ computes with a
            # separate operator (ts_ctc_loss) between forward and backward; the composite entry runs the reference's stand-in loss only
            raise NotImplementedError("native_train_step: transcript labels (real CTC loss) go through train_step, not the composite entry")
    model._step_seed += 1
    args, loss = native_step_args(model, optimizer, strategy, global_clip=1.0, dropout=dropout, seed=model._step_seed)
    p.ctx.check(p.lib.ts_w2v_step(p.h, ptr(x), B, N, ptr(neg), neg_bs, neg_ts, ptr(lab), C.byref(args), stream_ptr()))
    optimizer.iterations += 1
    p.weights_synced = True            # the update pass refreshed the bf16 compute copy
    model._last = {"x": x, "neg": neg, "labels": lab, "training": True}
    return loss[0]


def legacy_train_step(model, inputs, optimizer, neg_indices=None, dropout=True):
    """Legacy step of whisper_single.py — WS:1143-1180: no clipping, no NaN guard."""
    features, labels = inputs
    outputs = model(features, training=True, neg_indices=neg_indices, dropout=dropout)
    loss = outputs["loss"]
    gradients = model.gradient()
    optimizer.apply_gradients(gradients)
    return loss


def distributed_train_step(strategy, model, dist_inputs, optimizer, neg_indices=None, dropout=True):
    """V:1186-1260: per replica loss/N → grads → local clip_by_global_norm(1.0) → apply_gradients (all-reduce SUM,
    per-variable clipnorm, Adam) → strategy.reduce(SUM) of the scaled losses."""

    def step(inputs):
        features, labels = inputs
        if features.shape[0] == 0:
            return torch.zeros((), device=model._prog.device)
        n = float(strategy.num_replicas_in_sync)
        if isinstance(model, Wav2Vec2ForPreTraining):
            outputs = model(features, training=True, neg_indices=neg_indices, loss_div=n, dropout=dropout)
        else:   # V:1222-1224: outputs = model(features, labels=labels, training=True); loss = outputs["loss"]
            outputs = model(features, labels=labels, training=True, loss_div=n, dropout=dropout)
        scaled_loss = outputs["loss"] / n
        gradients = model.gradient()
        optimizer.apply_gradients(gradients, strategy=strategy, global_clip_norm=1.0)
        return scaled_loss

    per_replica_losses = strategy.run(step, args=(dist_inputs,))
    return strategy.reduce(ReduceOp.SUM, per_replica_losses, axis=None)


def _span_mask(hidden_states, start_mask, mask_prob, mask_length, axis):
    dev = torch.device("cuda", torch.cuda.current_device())
    x = hidden_states if isinstance(hidden_states, torch.Tensor) and hidden_states.is_cuda else to_device(hidden_states, torch.float32, dev)
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    x = x.contiguous()
    B, T, H = x.shape
    L = T if axis == 1 else H
    if start_mask is None:
        start_mask = torch.rand(B, L, device=x.device) < mask_prob          # tf.random.uniform(shape) < mask_prob (V:1078 / V:1103)
    start = to_device(start_mask, torch.uint8, x.device)
    y = torch.empty_like(x)
    expanded = torch.empty(B, L, dtype=torch.float32, device=x.device)
    ctx = _lib.context(x.device.index)
    dt = _lib.TS_F32 if x.dtype == torch.float32 else _lib.TS_BF16
    ctx.check(ctx.lib.ts_span_mask_apply(ctx.h, dt, ptr(x), ptr(start), ptr(y), ptr(expanded), B, T, H, axis, int(mask_length),
                                         stream_ptr()))
    return y, (expanded.unsqueeze(-1) if axis == 1 else expanded.unsqueeze(1))


def apply_time_mask(hidden_states, mask_prob=0.05, mask_length=10, start_mask=None):
    """V:1073-1095 — returns (masked_hidden_states, expanded_mask [B, T, 1]). `start_mask` ([B, T] bool) injects the span
    starts (parity tests); otherwise they are drawn on the device."""
    return _span_mask(hidden_states, start_mask, mask_prob, mask_length, 1)


def apply_feature_mask(hidden_states, mask_prob=0.05, mask_length=10, start_mask=None):
    """V:1098-1120 — returns (masked_hidden_states, expanded_mask [B, 1, H])."""
    return _span_mask(hidden_states, start_mask, mask_prob, mask_length, 2)


def make_graphed_distributed_step(strategy, model, optimizer, example_features, dropout=True, warmup=3):
    """distributed_train_step (V:1186-1260) replayed from CUDA graphs. With the native communicator (the default on GPUs) the whole
    step is ONE graph: [advance state, forward (loss / N), backward, local clip factor, pack] -> NCCL SUM of the gradient arena
    (captured) -> [per-variable clipnorm + Adam] -> NCCL SUM of the scaled losses (captured). With torch.distributed collectives
    (TETHYS_NATIVE_COMM=0) the two compute parts are graph segments around the eager all-reduce and the eager strategy.reduce.
    Returns (step(features) -> reduced loss, segments)."""
    from .runtime import GraphedSegments

    prog = model._prog
    n = float(strategy.num_replicas_in_sync)
    B, N = example_features.shape
    feats = example_features.to(prog.device).clone()
    T = model.num_frames(N)
    neg = model._sample_negative_indices(T, B)[:, 0, :].contiguous().clone()
    state = {}
    # local clip_by_global_norm (V:1243) folded into the collective: sum_r scale_r * g_r in one NCCL pre-multiplied sum
    premul = (strategy.dist is not None and strategy.dist.get_backend() == "nccl"
              and (getattr(strategy, "comm", None) is not None or hasattr(strategy.dist, "_make_nccl_premul_sum"))
              and not os.environ.get("TETHYS_NO_PREMUL"))
    clip_scale = torch.ones(1, device=prog.device)
    # bf16 compute: the gradient arena crosses NVLink as bf16 (half the bytes); the local clip factor rides on the pack
    lp = strategy.dist is not None and prog.ar_bf16()

    def seg_fwd_bwd():
        prog.ctx.check(prog.lib.ts_step_state_advance(prog.ctx.h, stream_ptr()))
        out = model(feats, training=True, neg_indices=neg, loss_div=n, dropout=dropout)
        state["scaled_loss"] = out["loss"] / n
        model.gradient()
        if lp:
            optimizer.local_clip_scale(model, 1.0, clip_scale)
            prog.pack_grads(scale=clip_scale)
        elif premul:
            optimizer.local_clip_scale(model, 1.0, clip_scale)     # factor only; applied inside the all-reduce
        else:
            optimizer.local_clip(model, 1.0)

    def seg_reduce():
        if lp:
            strategy.all_reduce_sum_(prog.grads_lp())
        elif premul:
            strategy.all_reduce_premul_sum_(prog.grads, clip_scale)
        else:
            strategy.all_reduce_sum_(prog.grads, bucket_elems=int(os.environ.get("TETHYS_AR_BUCKET_ELEMS", 0)))

    def seg_update():
        # bf16 buckets: clipnorm + Adam read the reduced bucket directly (no unpack pass; prog.grads keeps the LOCAL gradients)
        optimizer.update(model, grads_lp=prog.grads_lp() if lp else None)

    if getattr(strategy, "comm", None) is not None:
        # native communicator: the all-reduce is a stream-ordered NCCL call on the compute stream — the whole step, collective and
        # loss reduce included, is ONE CUDA graph (no host round trip at the fwd/bwd -> reduce -> Adam boundaries)
        # (measured at N = 2, profiles/r02_bench_n2_*: splitting the arena into buckets whose clipnorm + Adam run underneath the next
        # bucket's all-reduce LOSES 0.2 ms — the ring kernels and the update compete for HBM — so: one reduce, one update)
        def seg_all():
            seg_fwd_bwd()
            seg_reduce()
            seg_update()
            state["loss_red"] = strategy.reduce(ReduceOp.SUM, state["scaled_loss"], axis=None)

        segs = GraphedSegments([("graph", seg_all)], model, optimizer, warmup=warmup)

        def step_native(features):
            feats.copy_(features, non_blocking=True)
            neg.copy_(model._sample_negative_indices(T, B)[:, 0, :])      # V:907-937, drawn outside the graph
            segs()
            return state["loss_red"]

        return step_native, segs
    segs = GraphedSegments([("graph", seg_fwd_bwd), ("eager", seg_reduce), ("graph", seg_update)], model, optimizer, warmup=warmup)

    def step(features):
        feats.copy_(features, non_blocking=True)
        neg.copy_(model._sample_negative_indices(T, B)[:, 0, :])      # V:907-937, drawn outside the graph
        segs()
        return strategy.reduce(ReduceOp.SUM, state["scaled_loss"], axis=None)

    return step, segs
