"""TEST INFRASTRUCTURE ONLY — runs the UNMODIFIED reference scripts (/root/reference/speech_jobs/*.py) on oracle/tf_shim.py and
exposes them next to the oracle: same weights in, same seeded inputs in, outputs / losses / gradients / post-step weights out.

    ref = ref_runner.load("wav2vec2_dist")                    # the reference module object, imported as-is
    model = ref_runner.build_w2v(ref, "tiny", example_wave)   # ref.Wav2Vec2ForPreTraining(ref.Wav2Vec2Config("tiny")), built
    ref_runner.set_w2v_weights(model, oracle_weights)         # oracle name -> the reference's own tf.Variable
    out = model(wave, training=True)                          # the reference's call()

Only available where /root/reference exists (this container); never on the GPU box — tests that use it skip there and the GPU
suite relies on the golden vectors generated from it (tests/golden/make_ref_golden.py)."""
import importlib.util
import os
import sys
from collections import OrderedDict

import torch

from . import tf_shim

REF_ROOT = os.environ.get("TETHYS_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "speech_jobs", "wav2vec2_dist.py"))


def load(name, floatx=torch.float64):
    """Import speech_jobs/<name>.py from the reference tree, unmodified, with `tensorflow` = the shim."""
    tf_shim.install()
    tf_shim.set_floatx(floatx)
    key = f"_tethys_reference_{name}"
    if key in sys.modules:
        return sys.modules[key]
    path = os.path.join(REF_ROOT, "speech_jobs", name + ".py")
    spec = importlib.util.spec_from_file_location(key, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[key] = mod
    spec.loader.exec_module(mod)
    return mod


def zero_dropout(cfg):
    """Parity runs use dropout rate 0 (TF's RNG stream is not reproducible; the configs expose the rates: V:69-71, W:29-31)."""
    for k in ("hidden_dropout", "activation_dropout", "attention_dropout", "dropout"):
        if hasattr(cfg, k):
            setattr(cfg, k, 0.0)
    return cfg


# ---- Wav2Vec2 ---------------------------------------------------------------------------------------------------------
def w2v_variable_map(model, head="pretraining"):
    """oracle weight name -> the reference model's tf.Variable, by walking the reference's own attribute structure
    (V:229-281 feature extractor, V:302-460 encoder layers, V:550-667 heads/quantiser, V:746-766 model)."""
    base = model.wav2vec2
    fe = base.feature_extractor
    m = OrderedDict()
    for i, seq in enumerate(fe.conv_layers):
        conv, gn = seq._seq[0], seq._seq[1]
        m[f"fe.conv{i}.kernel"] = conv.kernel
        m[f"fe.conv{i}.gn.gamma"] = gn.gamma
        m[f"fe.conv{i}.gn.beta"] = gn.beta
    m["fe.pos_conv.kernel"], m["fe.pos_conv.bias"] = fe.pos_conv_embed.kernel, fe.pos_conv_embed.bias
    m["fe.layer_norm.gamma"], m["fe.layer_norm.beta"] = fe.layer_norm.gamma, fe.layer_norm.beta
    m["feature_projection.kernel"], m["feature_projection.bias"] = base.feature_projection.kernel, base.feature_projection.bias
    m["feature_projection_layer_norm.gamma"] = base.feature_projection_layer_norm.gamma
    m["feature_projection_layer_norm.beta"] = base.feature_projection_layer_norm.beta
    for l, layer in enumerate(base.encoder.layers):
        p = f"encoder.layers.{l}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            d = getattr(layer.attention, n)
            m[p + f"attention.{n}.kernel"], m[p + f"attention.{n}.bias"] = d.kernel, d.bias
        m[p + "attention_layer_norm.gamma"], m[p + "attention_layer_norm.beta"] = layer.attention_layer_norm.gamma, layer.attention_layer_norm.beta
        ff = layer.feed_forward
        m[p + "feed_forward.intermediate_dense.kernel"], m[p + "feed_forward.intermediate_dense.bias"] = ff.intermediate_dense.kernel, ff.intermediate_dense.bias
        m[p + "feed_forward.output_dense.kernel"], m[p + "feed_forward.output_dense.bias"] = ff.output_dense.kernel, ff.output_dense.bias
        m[p + "feed_forward_layer_norm.gamma"], m[p + "feed_forward_layer_norm.beta"] = layer.feed_forward_layer_norm.gamma, layer.feed_forward_layer_norm.beta
    q = base.quantizer
    m["quantizer.codevectors"] = q.codevectors
    m["quantizer.projection.kernel"], m["quantizer.projection.bias"] = q.projection.kernel, q.projection.bias
    if head == "pretraining":
        for n in ("project_hid", "project_q"):
            h = getattr(base, n)
            m[f"{n}.dense.kernel"], m[f"{n}.dense.bias"] = h.dense.kernel, h.dense.bias
            m[f"{n}.layer_norm.gamma"], m[f"{n}.layer_norm.beta"] = h.layer_norm.gamma, h.layer_norm.beta
    elif head == "ctc":
        m["lm_head.kernel"], m["lm_head.bias"] = model.lm_head.kernel, model.lm_head.bias
    else:
        m["classifier_proj.kernel"], m["classifier_proj.bias"] = model.projector.kernel, model.projector.bias
        m["classifier.kernel"], m["classifier.bias"] = model.classifier.kernel, model.classifier.bias
    return m


def build_w2v(ref, size, example_wave, model_type="pretraining", labels=None, dropout_off=True, cfg_edit=None):
    """The reference's own constructor path (create_full_model's body, V:1157-1182) + one call to create the variables."""
    cfg = ref.Wav2Vec2Config(size)
    cfg.num_negatives = 100                       # create_full_model sets it (V:1167)
    if dropout_off:
        zero_dropout(cfg)
    if cfg_edit:
        cfg_edit(cfg)
    cls = {"pretraining": "Wav2Vec2ForPreTraining", "asr": "Wav2Vec2ForCTC", "ctc": "Wav2Vec2ForCTC",
           "classification": "Wav2Vec2ForSequenceClassification"}[model_type]
    model = getattr(ref, cls)(cfg)
    if model_type == "pretraining":
        model(example_wave, training=True)
    else:
        model(example_wave, labels=labels, training=True)
    return model


def set_weights(var_map, weights):
    missing = set(var_map) ^ set(weights)
    assert not missing, f"weight names differ: {sorted(missing)[:6]}"
    for k, v in var_map.items():
        assert tuple(v.shape) == tuple(weights[k].shape), (k, tuple(v.shape), tuple(weights[k].shape))
        v.assign(weights[k])


def grads_by_name(var_map, model, loss):
    """tape.gradient(loss, model.trainable_variables) keyed by oracle name; None -> zeros as V:1237-1240 does."""
    tv = model.trainable_variables
    gs = torch.autograd.grad(loss, tv, allow_unused=True, retain_graph=True)
    by_id = {id(v): (torch.zeros_like(v) if g is None else g) for v, g in zip(tv, gs)}
    assert {id(v) for v in var_map.values()} == set(by_id), "trainable_variables and the name map cover different variables"
    return OrderedDict((k, by_id[id(v)].detach()) for k, v in var_map.items())


# ---- Whisper ----------------------------------------------------------------------------------------------------------
def whisper_variable_map(model):
    """oracle weight name -> the reference's variable (W:305-323 encoder, W:376-392 decoder, W:210-216 / W:240-253 layers)."""
    m = OrderedDict()
    enc, dec = model.model.encoder, model.model.decoder

    def attn(p, a):
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            d = getattr(a, n)
            m[p + n + ".kernel"], m[p + n + ".bias"] = d.kernel, d.bias

    def ln(p, l):
        m[p + ".gamma"], m[p + ".beta"] = l.gamma, l.beta

    def ffn(p, f):
        m[p + "fc1.kernel"], m[p + "fc1.bias"] = f.fc1.kernel, f.fc1.bias
        m[p + "fc2.kernel"], m[p + "fc2.bias"] = f.fc2.kernel, f.fc2.bias

    m["encoder.conv1.kernel"], m["encoder.conv1.bias"] = enc.conv1.kernel, enc.conv1.bias
    m["encoder.conv2.kernel"], m["encoder.conv2.bias"] = enc.conv2.kernel, enc.conv2.bias
    for l, layer in enumerate(enc.layers):
        p = f"encoder.layers.{l}."
        attn(p + "self_attn.", layer.self_attn)
        ln(p + "self_attn_layer_norm", layer.self_attn_layer_norm)
        ffn(p + "feed_forward.", layer.feed_forward)
        ln(p + "final_layer_norm", layer.final_layer_norm)
    ln("encoder.layer_norm", enc.layer_norm)
    m["decoder.embed_tokens.embeddings"] = dec.embed_tokens.embeddings
    for l, layer in enumerate(dec.layers):
        p = f"decoder.layers.{l}."
        attn(p + "self_attn.", layer.self_attn)
        ln(p + "self_attn_layer_norm", layer.self_attn_layer_norm)
        attn(p + "encoder_attn.", layer.encoder_attn)
        ln(p + "encoder_attn_layer_norm", layer.encoder_attn_layer_norm)
        ffn(p + "feed_forward.", layer.feed_forward)
        ln(p + "final_layer_norm", layer.final_layer_norm)
    ln("decoder.layer_norm", dec.layer_norm)
    m["lm_head.kernel"] = model.lm_head.kernel
    return m


def build_whisper(ref, cfg_edit, example_feats, example_labels, dropout_off=True):
    cfg = ref.WhisperConfig()
    if cfg_edit:
        cfg_edit(cfg)
    if dropout_off:
        zero_dropout(cfg)
    model = ref.WhisperForConditionalGeneration(cfg)
    model(example_feats, labels=example_labels, training=True)
    return model
