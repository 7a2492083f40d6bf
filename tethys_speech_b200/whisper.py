"""Host-side mirror of the reference's Whisper surface (speech_jobs/whisper_dist.py = W): WhisperConfig,
WhisperForConditionalGeneration, create_whisper_model, create_dummy_dataset, distributed_train_step — driving the native
program in libtethys.so (csrc/whisper_program.cu) through ctypes. No TensorFlow, no CPU fallback.
"""
import ctypes as C
import math
import os

import numpy as np
import torch

from . import _lib
from .runtime import Adam, GradientList, ProgramBase, ReduceOp, Strategy, get_strategy, ptr, stream_ptr, to_device  # noqa: F401


class WhisperConfig:
    """Same attributes as WhisperConfig — W:10-45 (defaults = the CLI's 'small' preset: d768, 12 heads, 4+4 layers)."""

    def __init__(self):
        self.d_model = 768
        self.encoder_layers = 4
        self.encoder_attention_heads = 12
        self.decoder_layers = 4
        self.decoder_attention_heads = 12
        self.d_ff = 3072
        self.n_mels = 80
        self.n_ctx = 1500
        self.vocab_size = 51865
        self.max_target_positions = 448
        self.dropout = 0.1
        self.attention_dropout = 0.1
        self.activation_dropout = 0.0
        self.activation_function = "gelu"
        self.layer_norm_eps = 1e-5
        self.init_std = 0.02
        self.pad_token_id = 0
        self.bos_token_id = 1
        self.eos_token_id = 2
        self.use_cache = True
        self.decoder_start_token_id = 50257


def _keras_order(cfg):
    """trainable_variables order of WhisperForConditionalGeneration (attribute-tracking order, W:73-94, W:210-216,
    W:240-253, W:305-322, W:376-392, W:536-545)."""
    names = ["encoder.conv1.kernel", "encoder.conv1.bias", "encoder.conv2.kernel", "encoder.conv2.bias"]

    def attn(p):
        out = []
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            out += [p + n + ".kernel", p + n + ".bias"]
        return out

    def ln(p):
        return [p + ".gamma", p + ".beta"]

    def ffn(p):
        return [p + "fc1.kernel", p + "fc1.bias", p + "fc2.kernel", p + "fc2.bias"]

    for l in range(cfg.encoder_layers):
        p = f"encoder.layers.{l}."
        names += attn(p + "self_attn.") + ln(p + "self_attn_layer_norm") + ffn(p + "feed_forward.") + ln(p + "final_layer_norm")
    names += ln("encoder.layer_norm")
    names += ["decoder.embed_tokens.embeddings"]
    for l in range(cfg.decoder_layers):
        p = f"decoder.layers.{l}."
        names += (attn(p + "self_attn.") + ln(p + "self_attn_layer_norm") + attn(p + "encoder_attn.") + ln(p + "encoder_attn_layer_norm")
                  + ffn(p + "feed_forward.") + ln(p + "final_layer_norm"))
    names += ln("decoder.layer_norm")
    names += ["lm_head.kernel"]
    return names


class _Program(ProgramBase):
    def __init__(self, cfg, precision, device):
        c = _lib.WhisperCfg()
        c.d_model, c.enc_layers, c.dec_layers, c.heads, c.d_ff = cfg.d_model, cfg.encoder_layers, cfg.decoder_layers, cfg.encoder_attention_heads, cfg.d_ff
        if cfg.decoder_attention_heads != cfg.encoder_attention_heads:
            raise ValueError("encoder and decoder head counts must match (they do in every reference preset)")
        c.n_mels, c.n_ctx, c.vocab, c.max_target = cfg.n_mels, cfg.n_ctx, cfg.vocab_size, cfg.max_target_positions
        c.start_token = cfg.decoder_start_token_id
        c.ln_eps, c.dropout, c.attention_dropout, c.activation_dropout = cfg.layer_norm_eps, cfg.dropout, cfg.attention_dropout, cfg.activation_dropout
        self.ccfg = c
        lib = _lib.load()
        super().__init__("ts_whisper", precision, device, lambda ctx_h, prec, out: lib.ts_whisper_create(ctx_h, C.byref(c), prec, out))


def _glorot_uniform(gen, shape, fan_in, fan_out, device):
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen, dtype=torch.float32, device=device) * 2 - 1) * limit


class WhisperForConditionalGeneration:
    """Mirror of `WhisperForConditionalGeneration` — W:536-616. `model(features, labels=labels, training=True)` returns a
    dict with "loss", "logits", "encoder_last_hidden_state", "last_hidden_state" (the remaining reference keys —
    attentions, hidden state tuples, past_key_values — are inference/debug outputs off the train path and are None)."""

    def __init__(self, config, precision="bf16", device=None, seed=0):
        self.config = config
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self._prog = _Program(config, precision, device)
        self._step_seed = seed * 1000003
        self._last = {}
        self._init_weights(seed)
        names = _keras_order(config)
        assert set(names) == set(self._prog.info), "parameter table mismatch"
        self.variable_names = names
        self.trainable_variables = [self._prog.view(self._prog.params, n) for n in names]

    def _init_weights(self, seed):
        """Keras defaults: Dense/Conv1D glorot_uniform + zeros, LayerNorm ones/zeros, Embedding uniform(-0.05, 0.05)."""
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
            elif name.endswith(".embeddings"):
                v.copy_((torch.rand(shp, generator=gen, dtype=torch.float32, device=p.device) * 2 - 1) * 0.05)
            elif name.endswith(".gamma"):
                v.fill_(1.0)
            else:
                v.zero_()
        p.weights_synced = False

    def set_weights(self, weights):
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

    def __call__(self, input_features, decoder_input_ids=None, attention_mask=None, decoder_attention_mask=None,
                 encoder_outputs=None, past_key_values=None, labels=None, use_cache=None, return_dict=True, training=False,
                 dropout=True):
        """input_features [B, n_mels, T_mel]; labels [B, S] int. The train call model(features, labels=..., training=True) (W:826) and
        the inference calls model(features, labels=...) / model(features, decoder_input_ids=...) are supported; masks / caches are
        not (generate() is the cached decode)."""
        if decoder_attention_mask is not None:
            # W:594-597 weights the loss with the mask, but W:575 also hands the same [B, S] tensor to the decoder as ITS attention mask,
            # where it replaces the causal mask and is added to [B, heads, S, S] scores (W:150-154): that broadcast only type-checks for
            # B == 1 or B == S. The reference's own loops never pass it (W:826, W:1003); there is no kernel for that attention pattern.
            raise NotImplementedError("decoder_attention_mask: not on the train path (see the comment above; the reference never passes it)")
        if any(a is not None for a in (attention_mask, encoder_outputs, past_key_values)):
            raise NotImplementedError("attention_mask / encoder_outputs / past_key_values: use generate() for cached decoding")
        p = self._prog
        if labels is None:
            # inference call model(features, decoder_input_ids=ids): logits only (W:579), loss None (W:585). The program builds its
            # decoder inputs as pad(labels[:, :-1], start) (W:559-563), so ids must begin with decoder_start_token_id.
            if decoder_input_ids is None:
                raise ValueError("either labels or decoder_input_ids are required (W:557-563)")
            if training:
                raise ValueError("training=True needs labels (the loss of W:585-600)")
            ids = to_device(decoder_input_ids, torch.int32, p.device)
            if not bool((ids[:, 0] == int(self.config.decoder_start_token_id)).all()):
                raise NotImplementedError("decoder_input_ids must begin with decoder_start_token_id")
            labels = torch.cat([ids[:, 1:], torch.zeros_like(ids[:, :1])], dim=1)
        elif decoder_input_ids is not None:
            ids = to_device(decoder_input_ids, torch.int32, p.device)
            lab_ = to_device(labels, torch.int32, p.device)
            want = torch.cat([torch.full_like(lab_[:, :1], int(self.config.decoder_start_token_id)), lab_[:, :-1]], dim=1)
            if ids.shape != want.shape or not bool((ids == want).all()):
                raise NotImplementedError("decoder_input_ids other than the right-shifted labels (W:559-563) are not on the train path")
        x = to_device(input_features, torch.float32, p.device)
        lab = to_device(labels, torch.int32, p.device)
        B, nm, Tm = x.shape
        S = lab.shape[1]
        p.ensure_workspace(B, Tm, S)
        p.sync_weights()
        self._step_seed += 1
        p.ctx.check(p.lib.ts_whisper_forward(p.h, ptr(x), B, Tm, ptr(lab), S, self._step_seed, 1 if (training and dropout) else 0,
                                             1 if training else 0, stream_ptr()))
        self._last = {"x": x, "labels": lab}
        scal = p.buffer("scalars")
        out = {"loss": scal[0] if training else None,
               "logits": p.buffer("logits")[:, :, :self.config.vocab_size],
               "past_key_values": None,
               "encoder_last_hidden_state": p.buffer("encoder_last_hidden_state"),
               "last_hidden_state": p.buffer("last_hidden_state"),
               "encoder_hidden_states": None, "encoder_attentions": None, "decoder_hidden_states": None,
               "decoder_attentions": None, "cross_attentions": None}
        return out

    call = __call__

    def generate(self, input_features, max_length=None, min_length=None, num_beams=None, temperature=1.0, top_k=None, top_p=None,
                 repetition_penalty=None, attention_mask=None, sync_every=8, **kwargs):
        """W:636-709: encoder once, then up to `max_length` (default max_target_positions) greedy steps, each re-running the
        decoder on the whole prefix (the reference passes no past_key_values, and under its anti-causal mask a self-attention
        cache would not be valid anyway); stops after the first step at which EVERY sequence emits eos_token_id (W:700-705).
        Returns int32 [B, 1 + steps] starting with decoder_start_token_id.

        temperature > 0 and the top-k filter (W:676-689) cannot change an argmax, min_length / top_p / repetition_penalty are
        read and ignored by the reference, and num_beams > 1 is a `pass` there that leaves next_tokens undefined (W:692-694):
        greedy is the only behaviour to reproduce. The host looks at the tokens every `sync_every` steps instead of after each
        one; steps computed past the stopping point are discarded, so the result is the same."""
        if attention_mask is not None:
            raise NotImplementedError("attention_mask is not supported by generate()")
        if num_beams not in (None, 1):
            raise NotImplementedError("num_beams > 1: the reference's beam-search branch is empty (W:692-694)")
        if temperature is not None and not temperature > 0:
            raise ValueError("temperature must be > 0")
        cfg = self.config
        max_length = int(max_length) if max_length is not None else cfg.max_target_positions
        if not 1 <= max_length <= cfg.max_target_positions:
            raise ValueError(f"max_length must be in [1, {cfg.max_target_positions}] (size of the positional table, W:383)")
        p = self._prog
        x = to_device(input_features, torch.float32, p.device)
        B, nm, Tm = x.shape
        p.ensure_workspace(B, Tm, max(2, max_length))
        p.sync_weights()
        p.ctx.check(p.lib.ts_whisper_encode(p.h, ptr(x), B, Tm, max_length, stream_ptr()))
        tokens = torch.zeros(B, max_length, dtype=torch.int32, device=p.device)
        steps, checked = max_length, 0
        for L in range(1, max_length + 1):
            p.ctx.check(p.lib.ts_whisper_decode_step(p.h, ptr(tokens), max_length, L, stream_ptr()))
            if L % max(1, int(sync_every)) == 0 or L == max_length:
                all_eos = (tokens[:, checked:L] == cfg.eos_token_id).all(dim=0)
                hit = torch.nonzero(all_eos)
                if hit.numel():
                    steps = checked + int(hit[0]) + 1
                    break
                checked = L
        self._last = {"x": x}
        start = torch.full((B, 1), cfg.decoder_start_token_id, dtype=torch.int32, device=p.device)
        return torch.cat([start, tokens[:, :steps]], dim=1)

    def gradient(self, stage_from=0, stage_to=10 ** 6):
        """tape.gradient(loss, model.trainable_variables) — W:833."""
        p = self._prog
        p.backward(stage_from, stage_to)
        gl = GradientList(p.view(p.grads, n) for n in self.variable_names)
        gl.owner = self
        for g in gl:
            g._ts_owner = self
        return gl

    def gradient_allreduced(self, strategy):
        """tape.gradient + the cross-replica SUM that apply_gradients performs (W:833-834), with the all-reduce of each
        arena bucket overlapped with the remaining backward stages. The returned list is marked as already reduced."""
        p = self._prog
        p.backward_allreduce_overlapped(strategy)
        gl = GradientList(p.view(p.grads, n) for n in self.variable_names)
        gl.owner = self
        gl.reduced = True
        return gl

    def save_weights(self, path):
        """Keras `model.save_weights` (W:1025): variables only, in the checkpoint.py container."""
        from . import checkpoint
        return checkpoint.save(path, self)

    def load_weights(self, path, strict=True):
        """The restore the reference never calls (SURVEY f-3): variables from a file written by save_weights / Checkpoint."""
        from . import checkpoint
        return checkpoint.restore(path, self, strict=strict)


def create_whisper_model(model_type="small", precision="bf16", device=None, seed=0):
    """W:852-890."""
    config = WhisperConfig()
    if model_type == "tiny":
        config.d_model, config.encoder_layers, config.decoder_layers, config.d_ff = 384, 4, 4, 1536
        config.encoder_attention_heads = config.decoder_attention_heads = 6
    elif model_type == "base":
        config.d_model, config.encoder_layers, config.decoder_layers, config.d_ff = 512, 6, 6, 2048
        config.encoder_attention_heads = config.decoder_attention_heads = 8
    elif model_type == "medium":
        config.d_model, config.encoder_layers, config.decoder_layers, config.d_ff = 1024, 24, 24, 4096
        config.encoder_attention_heads = config.decoder_attention_heads = 16
    elif model_type == "large":
        config.d_model, config.encoder_layers, config.decoder_layers, config.d_ff = 1280, 32, 32, 5120
        config.encoder_attention_heads = config.decoder_attention_heads = 20
    return WhisperForConditionalGeneration(config, precision=precision, device=device, seed=seed)


def create_dummy_dataset(batch_size, n_mels=80, seq_len=3000, max_target_length=100, num_samples=50, seed=1234):
    """W:784-815: 50 samples of N(0,1) mel [n_mels, seq_len] + labels (zeros; len ~ U{50..89}; [0]=1; [1:len-1] ~ U{3..99};
    [len-1]=2), batched and repeated. Seeded (the reference is not). Yields (features pinned fp32, labels int32)."""
    rng = np.random.default_rng(seed)
    feats = torch.from_numpy(rng.standard_normal((num_samples, n_mels, seq_len), dtype=np.float32))
    labels = np.zeros((num_samples, max_target_length), dtype=np.int32)
    lens = rng.integers(50, 90, size=num_samples)
    for i in range(num_samples):
        n = int(lens[i])
        labels[i, 0] = 1
        labels[i, 1:n - 1] = rng.integers(3, 100, size=n - 2)
        labels[i, n - 1] = 2
    labels = torch.from_numpy(labels)
    if torch.cuda.is_available():
        feats, labels = feats.pin_memory(), labels.pin_memory()

    def gen():
        while True:
            for i in range(0, num_samples, batch_size):   # dataset.batch() keeps the ragged last batch (W:815)
                yield feats[i:i + batch_size], labels[i:i + batch_size]

    return gen()


def create_dummy_waveform_dataset(batch_size, audio_seconds=30.0, sample_rate=16000, max_target_length=100, num_samples=50, seed=1234):
    """f-1: the dummy dataset of W:784-815 one stage earlier — N(0,1) WAVEFORMS [num_samples, seconds * 16000]
    instead of ready-made mel features; `waveform_to_features` turns a batch into the [B, 80, frames] model input with the
    fused log-mel kernel (extract_fbank_features, W:739-766). Labels as in create_dummy_dataset."""
    rng = np.random.default_rng(seed)
    n = int(round(audio_seconds * sample_rate))
    waves = torch.from_numpy(rng.standard_normal((num_samples, n), dtype=np.float32))
    labels = np.zeros((num_samples, max_target_length), dtype=np.int32)
    lens = rng.integers(50, 90, size=num_samples)
    for i in range(num_samples):
        k = int(lens[i])
        labels[i, 0] = 1
        labels[i, 1:k - 1] = rng.integers(3, 100, size=k - 2)
        labels[i, k - 1] = 2
    labels = torch.from_numpy(labels)
    if torch.cuda.is_available():
        waves, labels = waves.pin_memory(), labels.pin_memory()

    def gen():
        while True:
            for i in range(0, num_samples, batch_size):
                yield waves[i:i + batch_size], labels[i:i + batch_size]

    return gen()


def waveform_to_features(waveform, device=None):
    """[B, N] waveform -> [B, 80, F] log-mel in the layout WhisperEncoder.call consumes (W:326-329); F is made even by
    dropping the last frame if necessary (the conv stem halves it)."""
    from .frontend import extract_fbank_features

    mel = extract_fbank_features(waveform, mel_major=True, device=device)
    if mel.shape[-1] % 2:
        mel = mel[..., :-1].contiguous()
    return mel


def distributed_train_step(strategy, model, dist_inputs, optimizer, dropout=True):
    """W:819-848: per replica forward + mean CE loss, gradients, optimizer.apply_gradients (all-reduce SUM across replicas
    WITHOUT dividing — App. C-3 — then Adam), and strategy.reduce(SUM) of the per-replica losses."""

    def train_step(inputs):
        features, labels = inputs
        outputs = model(features, labels=labels, training=True, dropout=dropout)
        loss = outputs["loss"]
        if strategy.num_replicas_in_sync > 1 and not os.environ.get("TETHYS_NO_OVERLAP"):
            gradients = model.gradient_allreduced(strategy)      # bucketed NCCL all-reduce overlapped with backward
            optimizer.apply_gradients(gradients, strategy=strategy, already_reduced=True)
        else:
            gradients = model.gradient()
            optimizer.apply_gradients(gradients, strategy=strategy)
        return loss

    per_replica_losses = strategy.run(train_step, args=(dist_inputs,))
    return strategy.reduce(ReduceOp.SUM, per_replica_losses, axis=None)


def native_train_step(model, inputs, optimizer, dropout=True, strategy=None):
    """The per-replica body and the reduce of distributed_train_step (W:819-848) as ONE C-ABI call — ts_whisper_step, the composite
    entry of SURVEY §8 b-2: forward + shifted CE, backward, all-reduce SUM of the gradients (not divided by N), Adam, and the SUM of the
    replicas' losses, all enqueued on the current stream (strategy None / one replica: no collective). Same kernels as train_step;
    the all-reduce is one message after backward (the Python distributed step overlaps its buckets with backward instead)."""
    from .runtime import native_step_args

    features, labels = inputs
    p = model._prog
    x = to_device(features, torch.float32, p.device)
    lab = to_device(labels, torch.int32, p.device)
    B, nm, Tm = x.shape
    S = lab.shape[1]
    p.ensure_workspace(B, Tm, S)
    p.sync_weights()
    model._step_seed += 1
    args, loss = native_step_args(model, optimizer, strategy, global_clip=0.0, dropout=dropout, seed=model._step_seed)
    p.ctx.check(p.lib.ts_whisper_step(p.h, ptr(x), B, Tm, ptr(lab), S, C.byref(args), stream_ptr()))
    optimizer.iterations += 1
    p.weights_synced = True
    model._last = {"x": x, "labels": lab}
    return loss[0]


def _adam_buckets(model, bucket_elems=16 * 1024 * 1024):
    """Arena ranges that become final together (groups of backward stages, ts_whisper_stage_end), their per-range ts_optim and a
    side stream; built once per model."""
    prog = model._prog
    st = getattr(prog, "_adam_buckets", None)
    if st is None:
        groups, start, first = [], 0, 0
        ends = prog.stage_ends
        for s_, end in enumerate(ends):
            if end - start >= bucket_elems or s_ == len(ends) - 1:
                if end > start:
                    groups.append((first, s_, start, end))
                start, first = end, s_ + 1
        st = {"groups": groups, "optims": [prog.make_optim_range(a0, a1) for (_, _, a0, a1) in groups],
              "side": torch.cuda.Stream(device=prog.device)}
        prog._adam_buckets = st
    return st


def train_step(model, inputs, optimizer, dropout=True):
    """Single-device form of W:823-836 (the reference has no single-GPU Whisper script — SURVEY D1).

    TETHYS_OVERLAP_ADAM=1 (opt-in): the step has no gradient clipping (Adam 1e-4, W:901), so a variable can be updated as soon as its
    gradient is final — backward runs in stage groups (the arena is laid out in backward-completion order) and the Adam pass of each
    finished group goes to a side stream underneath the remaining backward stages (a group's parameters are no longer read by later
    stages; the arithmetic per element is the one of the single launch). MEASURED SLOWER on a B200 and therefore off by default:
    5.37 vs 5.10 ms/step (default preset), 5.68 vs 5.44 ms (base) — the HBM-bound Adam passes take bandwidth and SM slots from the
    backward GEMMs they run under and lengthen them by more than the 0.75 ms the update costs at the end (the same finding as for
    per-bucket updates under the all-reduce at N = 2)."""
    features, labels = inputs
    outputs = model(features, labels=labels, training=True, dropout=dropout)
    overlap = (os.environ.get("TETHYS_OVERLAP_ADAM", "0") == "1" and not getattr(optimizer, "clipnorm", None)
               and model._prog.device.type == "cuda")
    if not overlap:
        gradients = model.gradient()
        optimizer.apply_gradients(gradients)
        return outputs["loss"]
    prog = model._prog
    st = _adam_buckets(model)
    cur = torch.cuda.current_stream(prog.device)
    side = st["side"]
    for g, (s0, s1, a0, a1) in enumerate(st["groups"]):
        prog.backward(s0, s1)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            optimizer.update_range(model, st["optims"][g])
    cur.wait_stream(side)
    optimizer.iterations += 1
    return outputs["loss"]


def make_graphed_distributed_step(strategy, model, optimizer, example_features, example_labels, dropout=True, warmup=3,
                                  bucket_elems=16 * 1024 * 1024):
    """distributed_train_step (W:819-848) replayed from CUDA graphs. With the native communicator (default on GPUs) the whole step is
    ONE graph in which every bucket's NCCL all-reduce is a node on the communicator's side stream (a fork / join the capture
    records); with torch.distributed collectives (TETHYS_NATIVE_COMM=0) the compute parts are graph segments between eager
    all-reduces. The backward stages are grouped into buckets (arena prefixes that are final, ts_whisper_stage_end). Per bucket g:
        main stream : graph [backward stages of g] -> async all-reduce SUM of bucket g (un-normalised — App. C-3)
    so the reductions of earlier buckets run underneath the remaining backward graphs; after the last bucket the step waits
    for all reductions and replays the Adam graph. With TETHYS_SIDE_ADAM=1 each bucket is instead updated on a side stream as
    soon as it is reduced (a bucket's parameters are no longer read by later backward stages). Returns (step(features, labels) -> reduced loss, segments)."""
    from .runtime import GraphedSegments

    prog = model._prog
    dev = prog.device
    feats = example_features.to(dev).clone()
    labels = example_labels.to(dev).clone()
    state = {}
    ends = prog.stage_ends
    groups, start, first = [], 0, 0
    for s_, end in enumerate(ends):
        if end - start >= bucket_elems or s_ == len(ends) - 1:
            if end > start:
                groups.append((first, s_, start, end))
            start, first = end, s_ + 1
    # measured at N=2: updating reduced buckets on a side stream underneath backward does not pay (7.18 vs 7.10 ms: the Adam
    # passes compete with backward for HBM) — opt-in only
    side_adam = bool(os.environ.get("TETHYS_SIDE_ADAM"))
    lp = strategy.dist is not None and prog.ar_bf16() and not side_adam     # bf16 gradient buckets (bf16 compute only)
    optims = [prog.make_optim_range(a0, a1) for (_, _, a0, a1) in groups] if side_adam else None
    side = torch.cuda.Stream(device=dev)
    plan = []
    if getattr(strategy, "comm", None) is not None and not side_adam:
        # native communicator: the collectives are stream-ordered NCCL calls, so the WHOLE step is one CUDA graph — every bucket's
        # all-reduce forks onto the communicator's side stream inside the capture and runs underneath the remaining backward
        # stages; no host round trip at any boundary
        pack_on_main = bool(os.environ.get("TETHYS_PACK_ON_MAIN"))   # A/B switch

        def seg_all():
            prog.ctx.check(prog.lib.ts_step_state_advance(prog.ctx.h, stream_ptr()))
            out = model(feats, labels=labels, training=True, dropout=dropout)
            state["loss"] = out["loss"]
            for (s0, s1, a0, a1) in groups:
                prog.backward(s0, s1)
                if lp and pack_on_main:
                    prog.pack_grads(a0, a1)
                    strategy.all_reduce_async_(prog.grads_lp()[a0:a1])
                elif lp:  # the bucket is packed to bf16 on the communicator's stream too, underneath the next backward stages
                    strategy.all_reduce_async_(prog.grads_lp()[a0:a1], pre=lambda a0=a0, a1=a1: prog.pack_grads(a0, a1))
                else:
                    strategy.all_reduce_async_(prog.grads[a0:a1])
            strategy.join_async()
            optimizer.update(model, grads_lp=prog.grads_lp() if lp else None)     # bf16 buckets are read as they are; (per-bucket updates underneath the last all-reduce were measured slower at N = 2: 7.61 vs 7.21 ms)
            state["loss_red"] = strategy.reduce(ReduceOp.SUM, state["loss"], axis=None)     # W:848, inside the graph as well

        segs = GraphedSegments([("graph", seg_all)], model, optimizer, warmup=warmup)

        def step_native(features, lab):
            feats.copy_(features, non_blocking=True)
            labels.copy_(lab, non_blocking=True)
            segs()
            return state["loss_red"]

        return step_native, segs

    def seg_forward():
        prog.ctx.check(prog.lib.ts_step_state_advance(prog.ctx.h, stream_ptr()))
        out = model(feats, labels=labels, training=True, dropout=dropout)
        state["loss"] = out["loss"]

    works = []
    for gi, (s0, s1, a0, a1) in enumerate(groups):
        def seg_bwd(s0=s0, s1=s1, gi=gi, a0=a0, a1=a1):
            if gi == 0:
                seg_forward()
            prog.backward(s0, s1)
            if lp:
                prog.pack_grads(a0, a1)

        def seg_reduce(a0=a0, a1=a1):
            bucket = prog.grads_lp()[a0:a1] if lp else prog.grads[a0:a1]
            works.append(strategy.dist.all_reduce(bucket, op=strategy.dist.ReduceOp.SUM, async_op=True))

        plan += [("graph", seg_bwd), ("eager", seg_reduce)]
        if side_adam:
            def seg_side_wait():
                side.wait_stream(torch.cuda.current_stream(dev))    # Adam's state / step counter are ordered after this step's start
                with torch.cuda.stream(side):
                    works.pop(0).wait()

            def seg_side_update(gi=gi):
                optimizer.update_range(model, optims[gi])

            plan += [("eager", seg_side_wait), ("side_graph", seg_side_update)]

    def seg_join():
        if side_adam:
            torch.cuda.current_stream(dev).wait_stream(side)
        else:
            while works:
                works.pop(0).wait()

    plan += [("eager", seg_join)]
    def seg_update():
        if lp:
            prog.unpack_grads()
        optimizer.update(model)

    if not side_adam:
        plan += [("graph", seg_update)]
    segs = GraphedSegments(plan, model, optimizer, warmup=warmup, side_stream=side, bump_iterations=side_adam)

    def step(features, lab):
        feats.copy_(features, non_blocking=True)
        labels.copy_(lab, non_blocking=True)
        segs()
        return strategy.reduce(ReduceOp.SUM, state["loss"], axis=None)

    return step, segs
