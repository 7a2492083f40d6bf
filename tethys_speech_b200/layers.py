"""The reference's sub-layer classes as thin fronts over the library.

Two kinds:
  * operator-level layers that own their variables and call the single-operator C-ABI entries (ts_gemm, ts_attn_fwd,
    ts_layernorm_fwd, ts_groupnorm_fwd):
        MultiHeadAttention(config, is_decoder, is_cross_attention)   W:73-176
        FeedForward(config, is_decoder)                              W:180-206
        GroupNormalization(groups, axis, epsilon)                    V:140-196
        Wav2Vec2ProjectionHead(config)                               V:550-561
  * program-level views that run the corresponding slice of a whole-model program (the fused kernels the train step uses)
    and hand back the tensor the reference's layer returns:
        WhisperEncoder / WhisperDecoder / WhisperModel               W:305-372, W:376-466, W:470-532
        Wav2Vec2FeatureExtractor / Wav2Vec2Encoder / Wav2Vec2Quantizer   V:229-298, V:464-546, V:564-667
Forward (`call`) only — gradients on the hot path come from the whole-step programs (model.gradient()). Same constructor
arguments, attribute names (`q_proj.kernel`-style variable paths, `.trainable_variables`, `.config`) and argument meaning as
the reference; arguments that would need a kernel this library does not have raise NotImplementedError (no fallback)."""
import ctypes as C
import math

import torch

from . import _lib
from .runtime import ptr, stream_ptr, to_device

_TD = {"fp32": torch.float32, "float32": torch.float32, "bf16": torch.bfloat16, "bfloat16": torch.bfloat16}


def _dev(device):
    return torch.device("cuda", torch.cuda.current_device() if device is None else device)


class _Dense:
    """tf.keras.layers.Dense: kernel [in, out] (glorot_uniform), bias [out] (zeros) — fp32 masters, bf16 compute copy on demand."""

    def __init__(self, gen, n_in, n_out, device, use_bias=True):
        lim = math.sqrt(6.0 / (n_in + n_out))
        self.kernel = (torch.rand(n_in, n_out, generator=gen, device=device) * 2 - 1) * lim
        self.bias = torch.zeros(n_out, device=device) if use_bias else None

    def variables(self, prefix):
        out = [(prefix + ".kernel", self.kernel)]
        if self.bias is not None:
            out.append((prefix + ".bias", self.bias))
        return out


class _Layer:
    def __init__(self, precision, device):
        self.precision = precision
        self.dtype = _TD[precision]
        self.ts_dtype = _lib.TS_F32 if self.dtype == torch.float32 else _lib.TS_BF16
        self.device = _dev(device)
        self.ctx = _lib.context(self.device.index)
        self._named = []

    @property
    def trainable_variables(self):
        return [v for _, v in self._named]

    @property
    def variable_names(self):
        return [n for n, _ in self._named]

    def set_weights(self, weights):
        for n, v in self._named:
            if n in weights:
                v.copy_(to_device(weights[n], torch.float32, self.device).view(v.shape))

    def _gemm(self, x2d, dense, act=0, drop=0.0, seed=0, residual=None):
        """y = act(x @ kernel + bias) [dropout] [+ residual] through ts_gemm (tcgen05 engine for bf16 operands)."""
        m, k = x2d.shape
        n = dense.kernel.shape[1]
        w = dense.kernel.to(self.dtype).contiguous()
        y = torch.empty(m, n, device=self.device, dtype=self.dtype)
        d = _lib.GemmDesc()
        d.a, d.b, d.c = x2d.data_ptr(), w.data_ptr(), y.data_ptr()
        d.m, d.n, d.k, d.a_major, d.b_major = m, n, k, 0, 1
        d.lda, d.ldb, d.ldc = k, n, n
        d.batch1 = d.batch2 = 1
        d.in_dtype = d.out_dtype = self.ts_dtype
        d.alpha, d.act, d.drop, d.seed = 1.0, act, float(drop), int(seed)
        if dense.bias is not None:
            d.bias = dense.bias.data_ptr()
        if residual is not None:
            d.residual, d.ldr = residual.data_ptr(), n
        self.ctx.check(self.ctx.lib.ts_gemm(self.ctx.h, C.byref(d), stream_ptr()))
        return y

    def _ln(self, x2d, gamma, beta, eps):
        rows, cols = x2d.shape
        y = torch.empty_like(x2d)
        mean = torch.empty(rows, device=self.device); rstd = torch.empty(rows, device=self.device)
        self.ctx.check(self.ctx.lib.ts_layernorm_fwd(self.ctx.h, self.ts_dtype, ptr(x2d), ptr(gamma), ptr(beta), ptr(y), ptr(mean), ptr(rstd),
                                                     rows, cols, float(eps), stream_ptr()))
        return y


# ---------------------------------------------------------------------------------------------------------------------
# operator-level layers
# ---------------------------------------------------------------------------------------------------------------------
class MultiHeadAttention(_Layer):
    """W:73-176. call(hidden_states, key_value_states=None, attention_mask=None, training=False) -> [B, L, d_model].
    attention_mask: None, an all-ones mask (adds 0) or the decoder's `1 - band_part(ones, -1, 0)` mask of W:416-418 (the only
    masks the reference ever builds); q is scaled by head_dim^-0.5 (W:141), dropout acts on the probabilities (W:160)."""

    def __init__(self, config, is_decoder=False, is_cross_attention=False, precision="bf16", device=None, seed=0):
        super().__init__(precision, device)
        if self.dtype != torch.bfloat16:
            raise NotImplementedError("MultiHeadAttention as a single layer runs on the fused bf16 attention operator (K10, ts_attn_fwd); "
                                      "fp32 attention exists inside the whole-model programs (parity mode) only")
        self.config, self.is_decoder, self.is_cross_attention = config, is_decoder, is_cross_attention
        self.num_heads = config.decoder_attention_heads if is_decoder else config.encoder_attention_heads
        self.d_model = config.d_model
        self.head_dim = self.d_model // self.num_heads
        if self.head_dim != 64:
            raise NotImplementedError("the fused attention operator is built for head_dim 64 (every preset of the reference)")
        self.scaling = self.head_dim ** -0.5
        self.attention_dropout = float(config.attention_dropout)
        gen = torch.Generator(device=self.device); gen.manual_seed(seed)
        self.k_proj, self.v_proj, self.q_proj, self.out_proj = (_Dense(gen, self.d_model, self.d_model, self.device) for _ in range(4))
        self._named = (self.k_proj.variables("k_proj") + self.v_proj.variables("v_proj") + self.q_proj.variables("q_proj")
                       + self.out_proj.variables("out_proj"))
        self._seed = seed * 7919

    def _mask_mode(self, mask, L, Lk):
        if mask is None:
            return 0
        m = to_device(mask, torch.float32, self.device).reshape(-1, L, Lk)[0]
        if bool((m == 1).all()):
            return 0
        band = 1.0 - torch.tril(torch.ones(L, Lk, device=self.device))
        if L == Lk and bool((m == band).all()):
            return 1
        raise NotImplementedError("attention_mask: only None / all-ones / the decoder mask of W:416-418 have a kernel")

    def __call__(self, hidden_states, key_value_states=None, attention_mask=None, training=False):
        x = to_device(hidden_states, self.dtype, self.device)
        kv = x if key_value_states is None else to_device(key_value_states, self.dtype, self.device)
        B, L, d = x.shape
        Lk = kv.shape[1]
        q = self._gemm(x.reshape(B * L, d), self.q_proj)
        k = self._gemm(kv.reshape(B * Lk, d), self.k_proj)
        v = self._gemm(kv.reshape(B * Lk, d), self.v_proj)
        o = torch.empty(B, L, d, device=self.device, dtype=self.dtype)
        stats = torch.empty(B, self.num_heads, L, 2, device=self.device)
        a = _lib.AttnDesc()
        a.q, a.k, a.v, a.o = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr()
        a.q_ld = a.kv_ld = a.o_ld = d
        a.q_bs, a.kv_bs, a.o_bs = L * d, Lk * d, L * d
        a.stats = stats.data_ptr()
        a.batch, a.heads, a.tq, a.tk, a.head_dim = B, self.num_heads, L, Lk, self.head_dim
        a.scale, a.mask_mode = self.scaling, self._mask_mode(attention_mask, L, Lk)
        self._seed += 1
        a.drop, a.seed = (self.attention_dropout if training else 0.0), self._seed
        self.ctx.check(self.ctx.lib.ts_attn_fwd(self.ctx.h, C.byref(a), stream_ptr()))
        return self._gemm(o.reshape(B * L, d), self.out_proj).reshape(B, L, d)

    call = __call__


class FeedForward(_Layer):
    """W:180-206: fc1 -> exact GELU -> dropout(activation_dropout) -> fc2 -> dropout(config.dropout)."""

    def __init__(self, config, is_decoder=False, precision="bf16", device=None, seed=0):
        super().__init__(precision, device)
        self.config, self.is_decoder = config, is_decoder
        gen = torch.Generator(device=self.device); gen.manual_seed(seed)
        self.fc1 = _Dense(gen, config.d_model, config.d_ff, self.device)
        self.fc2 = _Dense(gen, config.d_ff, config.d_model, self.device)
        self._named = self.fc1.variables("fc1") + self.fc2.variables("fc2")
        self._seed = seed * 104729

    def __call__(self, hidden_states, training=False):
        x = to_device(hidden_states, self.dtype, self.device)
        shp = x.shape
        self._seed += 2
        h = self._gemm(x.reshape(-1, shp[-1]), self.fc1, act=1, drop=float(self.config.activation_dropout) if training else 0.0, seed=self._seed)
        y = self._gemm(h, self.fc2, drop=float(self.config.dropout) if training else 0.0, seed=self._seed + 1)
        return y.reshape(shp)

    call = __call__


class GroupNormalization(_Layer):
    """V:140-196: statistics over (time, channels-in-group) per (batch, group), then the per-channel gamma / beta.
    build(input_shape) is implicit at the first call, as in Keras."""

    def __init__(self, groups=32, axis=-1, epsilon=1e-5, precision="bf16", device=None, **kwargs):
        super().__init__(precision, device)
        if axis != -1:
            raise NotImplementedError("GroupNormalization: the reference only normalises the last axis")
        self.groups, self.axis, self.epsilon = groups, axis, epsilon
        self.gamma = self.beta = None

    def build(self, input_shape):
        c = int(input_shape[-1])
        if c % self.groups:
            raise ValueError(f"channels ({c}) must be divisible by groups ({self.groups})")   # V:150-156
        self.gamma = torch.ones(c, device=self.device)
        self.beta = torch.zeros(c, device=self.device)
        self._named = [("gamma", self.gamma), ("beta", self.beta)]

    def __call__(self, inputs):
        x = to_device(inputs, self.dtype, self.device).contiguous()
        if self.gamma is None:
            self.build(x.shape)
        B, T, Cc = x.shape
        y = torch.empty_like(x)
        mean = torch.empty(B, self.groups, device=self.device); rstd = torch.empty_like(mean)
        acc = torch.zeros(2 * B * self.groups, dtype=torch.float64, device=self.device)
        self.ctx.check(self.ctx.lib.ts_groupnorm_fwd(self.ctx.h, self.ts_dtype, ptr(x), ptr(self.gamma), ptr(self.beta), ptr(y), ptr(mean),
                                                     ptr(rstd), ptr(acc), B, T, Cc, self.groups, float(self.epsilon), stream_ptr()))
        return y

    call = __call__


class Wav2Vec2ProjectionHead(_Layer):
    """V:550-561: Dense(proj_codevector_dim) -> LayerNormalization(layer_norm_eps) -> Dropout(hidden_dropout)."""

    def __init__(self, config, precision="bf16", device=None, seed=0, in_dim=None):
        super().__init__(precision, device)
        self.config = config
        gen = torch.Generator(device=self.device); gen.manual_seed(seed)
        self.dense = _Dense(gen, in_dim or config.hidden_size, config.proj_codevector_dim, self.device)
        self.gamma = torch.ones(config.proj_codevector_dim, device=self.device)
        self.beta = torch.zeros(config.proj_codevector_dim, device=self.device)
        self._named = self.dense.variables("dense") + [("layer_norm.gamma", self.gamma), ("layer_norm.beta", self.beta)]
        self._seed = seed * 15485863

    def __call__(self, hidden_states, training=False):
        x = to_device(hidden_states, self.dtype, self.device)
        shp = x.shape
        y = self._ln(self._gemm(x.reshape(-1, shp[-1]), self.dense), self.gamma, self.beta, self.config.layer_norm_eps)
        if training and self.config.hidden_dropout > 0:
            self._seed += 1
            self.ctx.check(self.ctx.lib.ts_dropout(self.ctx.h, self.ts_dtype, ptr(y), ptr(y), y.numel(), float(self.config.hidden_dropout),
                                                   self._seed, stream_ptr()))
        return y.reshape(*shp[:-1], -1)

    call = __call__


# ---------------------------------------------------------------------------------------------------------------------
# program-level views
# ---------------------------------------------------------------------------------------------------------------------
class _WhisperView:
    def __init__(self, config, precision="bf16", device=None, seed=0, _owner=None):
        from .whisper import WhisperForConditionalGeneration

        self.config = config
        self._owner = _owner or WhisperForConditionalGeneration(config, precision=precision, device=device, seed=seed)
        self._prefix = ""

    @property
    def variable_names(self):
        return [n for n in self._owner.variable_names if n.startswith(self._prefix)]

    @property
    def trainable_variables(self):
        p = self._owner._prog
        return [p.view(p.params, n) for n in self.variable_names]

    def set_weights(self, weights):
        self._owner.set_weights(weights)
        self._owner._prog.weights_synced = False


class WhisperEncoder(_WhisperView):
    """W:305-372: conv stem, positional encoding, the encoder layers, final LayerNorm. call(input_features [B, n_mels, T]) ->
    dict(last_hidden_state [B, T/2, d_model])."""

    def __init__(self, config, precision="bf16", device=None, seed=0, _owner=None):
        super().__init__(config, precision, device, seed, _owner)
        self._prefix = "encoder."

    def __call__(self, input_features, attention_mask=None, output_attentions=False, output_hidden_states=False, training=False):
        if attention_mask is not None or output_attentions or output_hidden_states:
            raise NotImplementedError("WhisperEncoder: masks / attention / hidden-state outputs are not on the hot path")
        if training:
            raise NotImplementedError("WhisperEncoder alone runs in inference mode (dropout off); the training pass is the whole-model step")
        p = self._owner._prog
        x = to_device(input_features, torch.float32, p.device)
        B, _, Tm = x.shape
        p.ensure_workspace(B, Tm, int(self.config.max_target_positions))
        p.sync_weights()
        p.ctx.check(p.lib.ts_whisper_encode(p.h, ptr(x), B, Tm, int(self.config.max_target_positions), stream_ptr()))
        out = p.buffer("encoder_last_hidden_state")
        self._owner._last_encoder = (x, out)
        return {"last_hidden_state": out, "hidden_states": None, "attentions": None}

    call = __call__


class WhisperModel(_WhisperView):
    """W:470-532: encoder + decoder without the lm_head. call(input_features, decoder_input_ids) -> dict(last_hidden_state,
    encoder_last_hidden_state, ...). decoder_input_ids must start with config.decoder_start_token_id (that is how the program
    builds them from labels, W:557-563)."""

    def __init__(self, config, precision="bf16", device=None, seed=0, _owner=None):
        super().__init__(config, precision, device, seed, _owner)
        self.encoder = WhisperEncoder(config, _owner=self._owner)
        self.decoder = WhisperDecoder(config, _owner=self._owner)

    def __call__(self, input_features, decoder_input_ids=None, attention_mask=None, decoder_attention_mask=None, encoder_outputs=None,
                 past_key_values=None, use_cache=None, training=False, **kwargs):
        if any(a is not None for a in (attention_mask, decoder_attention_mask, encoder_outputs, past_key_values)):
            raise NotImplementedError("WhisperModel: masks / precomputed encoder outputs / caches are not on the hot path")
        if decoder_input_ids is None:
            raise ValueError("decoder_input_ids are required")
        ids = to_device(decoder_input_ids, torch.int32, self._owner._prog.device)
        if not bool((ids[:, 0] == int(self.config.decoder_start_token_id)).all()):
            raise NotImplementedError("decoder_input_ids must begin with decoder_start_token_id")
        labels = torch.cat([ids[:, 1:], torch.zeros_like(ids[:, :1])], dim=1)      # pad(labels[:, :-1], start) == ids (W:559-563)
        out = self._owner(input_features, labels=labels, training=training, dropout=training)
        return {"last_hidden_state": out["last_hidden_state"], "past_key_values": None,
                "encoder_last_hidden_state": out["encoder_last_hidden_state"], "decoder_hidden_states": None,
                "decoder_attentions": None, "cross_attentions": None, "encoder_hidden_states": None, "encoder_attentions": None}

    call = __call__


class WhisperDecoder(_WhisperView):
    """W:376-466: embedding + positional encoding, the decoder layers (self-attention under the mask of W:416-418, cross-attention,
    FFN), final LayerNorm. call(input_ids, encoder_hidden_states) -> dict(last_hidden_state [B, S, d_model]); the encoder states
    must be the ones this object's sibling WhisperEncoder produced last (the program keeps them on the device)."""

    def __init__(self, config, precision="bf16", device=None, seed=0, _owner=None):
        super().__init__(config, precision, device, seed, _owner)
        self._prefix = "decoder."

    def __call__(self, input_ids, encoder_hidden_states=None, attention_mask=None, encoder_attention_mask=None, past_key_values=None,
                 use_cache=None, training=False, **kwargs):
        if any(a is not None for a in (attention_mask, encoder_attention_mask, past_key_values)):
            raise NotImplementedError("WhisperDecoder: masks / caches are not on the hot path")
        last = getattr(self._owner, "_last_encoder", None)
        if last is None or encoder_hidden_states is None or encoder_hidden_states.data_ptr() != last[1].data_ptr():
            raise NotImplementedError("encoder_hidden_states must come from the sibling WhisperEncoder's last call")
        return {"last_hidden_state": WhisperModel(self.config, _owner=self._owner)(last[0], decoder_input_ids=input_ids,
                                                                                  training=training)["last_hidden_state"]}

    call = __call__


class _W2VView:
    def __init__(self, config, precision="bf16", device=None, seed=0, _owner=None):
        from .wav2vec2 import Wav2Vec2ForPreTraining

        self.config = config
        self._owner = _owner or Wav2Vec2ForPreTraining(config, precision=precision, device=device, seed=seed)
        self._prefixes = ("",)

    @property
    def variable_names(self):
        return [n for n in self._owner.variable_names if n.startswith(self._prefixes)]

    @property
    def trainable_variables(self):
        p = self._owner._prog
        return [p.view(p.params, n) for n in self.variable_names]

    def set_weights(self, weights):
        self._owner.set_weights(weights)


class Wav2Vec2FeatureExtractor(_W2VView):
    """V:229-298: the conv stack with GroupNorm + GELU, positional conv embedding, LayerNorm. call(waveform [B, N]) ->
    [B, N/320, conv_dim[-1]] (inference mode, like `extract_features` of the task models)."""

    def __init__(self, config, precision="bf16", device=None, seed=0, _owner=None):
        super().__init__(config, precision, device, seed, _owner)
        self._prefixes = ("fe.",)

    def __call__(self, inputs, training=False):
        if training:
            raise NotImplementedError("Wav2Vec2FeatureExtractor alone runs in inference mode; the training pass is the whole-model step")
        return self._owner.extract_features(inputs)

    call = __call__


class Wav2Vec2Encoder(_W2VView):
    """V:464-546 as it sits in the model: call(waveform) runs the trunk in inference mode and returns
    dict(last_hidden_state [B, T, hidden]) — the encoder consumes the projected features of the same program."""

    def __init__(self, config, precision="bf16", device=None, seed=0, _owner=None):
        super().__init__(config, precision, device, seed, _owner)
        self._prefixes = ("encoder.",)

    def __call__(self, inputs, attention_mask=None, output_attentions=False, output_hidden_states=False, training=False):
        if attention_mask is not None or output_attentions or output_hidden_states or training:
            raise NotImplementedError("Wav2Vec2Encoder view: inference call on a waveform only")
        return {"last_hidden_state": self._owner(inputs, training=False)["last_hidden_state"], "hidden_states": None, "attentions": None}

    call = __call__


class Wav2Vec2Quantizer(_W2VView):
    """V:564-667: hard nearest-codeword quantiser. call(waveform) -> (quantized_features [B, T, codevector_dim], perplexity,
    code indices int64 [groups, B, T]) — the quantiser input is the feature-extractor output of the same program (V:784-789)."""

    def __init__(self, config, precision="bf16", device=None, seed=0, _owner=None):
        super().__init__(config, precision, device, seed, _owner)
        self._prefixes = ("quantizer.",)

    def __call__(self, inputs, training=True):
        out = self._owner(inputs, training=True, dropout=False)
        return out["quantized_features"], out["codevector_perplexity"], out["code_indices"]

    call = __call__
