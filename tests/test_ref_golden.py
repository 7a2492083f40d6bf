"""CPU: the oracle reproduces, to 1e-10, the golden vectors that the REFERENCE'S OWN CODE produced (tests/golden/ref_*.npz, written
by tests/golden/make_ref_golden.py from the unmodified /root/reference scripts running on oracle/tf_shim.py). Unlike
tests/test_reference_pinning.py this needs no /root/reference: it also runs on the GPU box, and the same fixtures are the bar
of the CUDA path in tests/test_ref_golden_gpu.py."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MAX_ELEMS = 1500
W2V_GRADS = ("fe.conv0.kernel", "fe.conv2.gn.gamma", "encoder.layers.1.attention.q_proj.kernel", "encoder.layers.3.feed_forward.output_dense.bias",
             "quantizer.codevectors", "project_q.dense.kernel")
WH_GRADS = ("encoder.conv1.kernel", "encoder.layers.1.self_attn.q_proj.kernel", "decoder.embed_tokens.embeddings",
            "decoder.layers.0.self_attn.k_proj.kernel", "decoder.layers.1.encoder_attn.v_proj.bias", "lm_head.kernel")


def sub(t):
    a = t.detach().double().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
    flat = a.reshape(-1)
    return flat[::-(-flat.size // MAX_ELEMS)].copy()


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64).reshape(-1), np.asarray(b, dtype=np.float64).reshape(-1)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


def whisper_edit(c):
    c.d_model, c.d_ff = 64, 128
    c.encoder_layers = c.decoder_layers = 2
    c.encoder_attention_heads = c.decoder_attention_heads = 2
    c.vocab_size, c.n_mels, c.n_ctx, c.decoder_start_token_id = 203, 16, 64, 200


def w2v_case():
    from oracle import wav2vec2_oracle as O

    z = np.load(os.path.join(GOLD, "ref_w2v_tiny.npz"))
    seed = int(z["seed"])
    ocfg = O.Wav2Vec2Config(str(z["size"]))
    w = O.randomize_weights(O.init_weights(ocfg, seed, torch.float64), seed + 1)
    wave = torch.from_numpy(z["wave"]).double()
    neg = O.negative_indices_from_random(torch.from_numpy(z["random_ints"]), ocfg.num_negatives)
    step_negs = [O.negative_indices_from_random(torch.from_numpy(r), ocfg.num_negatives) for r in z["step_random_ints"]]
    return z, O, ocfg, w, wave, neg, step_negs


def whisper_case(which="ref_whisper_small_cfg.npz", dtype=torch.float64):
    from oracle import whisper_oracle as O

    z = np.load(os.path.join(GOLD, which))
    seed = int(z["seed"])
    ocfg = O.WhisperConfig("small")
    whisper_edit(ocfg)
    w = O.randomize_weights(O.init_weights(ocfg, seed, dtype), seed + 1)
    return z, O, ocfg, w, torch.from_numpy(z["feats"]).to(dtype), torch.from_numpy(z["labels"])


def test_oracle_reproduces_the_reference_w2v_vectors():
    z, O, ocfg, w, wave, neg, step_negs = w2v_case()
    out, g = O.loss_and_grads(ocfg, w, wave, neg)
    assert np.array_equal(out["code_indices"].numpy(), z["code_indices"])                      # integer work: bit-exact
    assert abs(float(out["loss"]) - float(z["loss"])) < 1e-10 * abs(float(z["loss"]))
    assert abs(float(out["codevector_perplexity"]) - float(z["perplexity"])) < 1e-10 * float(z["perplexity"])
    for key, okey in (("logits_sub", "contrastive_logits"), ("last_hidden_sub", "last_hidden_state"), ("extract_features_sub", "extract_features")):
        assert rel(sub(out[okey]), z[key]) < 1e-10, key
    for k in W2V_GRADS:
        assert rel(sub(g[k]), z["grad::" + k]) < 1e-8, k
    w0 = {k: v.clone() for k, v in w.items()}
    m = {k: torch.zeros_like(v) for k, v in w.items()}
    v_ = {k: torch.zeros_like(v) for k, v in w.items()}
    for t in (1, 2):
        o = O.train_step(ocfg, w, m, v_, t, wave, step_negs[t - 1], lr=3e-5, eps=1e-8)
        assert abs(float(o["loss"]) - float(z["step_losses"][t - 1])) < 1e-10 * abs(float(z["step_losses"][t - 1]))
    for k in W2V_GRADS:
        assert rel(sub(w[k] - w0[k]), z["delta2::" + k]) < 1e-8, k


def test_oracle_reproduces_the_reference_whisper_vectors():
    from oracle import whisper_oracle

    z, O, ocfg, w, feats, labels = whisper_case()
    old = whisper_oracle.EMULATE_FP32_ABSORPTION
    try:
        whisper_oracle.EMULATE_FP32_ABSORPTION = False       # the float64 reference run performs score + (-1e9) exactly
        out, g = O.loss_and_grads(ocfg, w, feats, labels)
        assert abs(float(out["loss"]) - float(z["loss"])) < 1e-10 * abs(float(z["loss"]))
        assert rel(sub(out["logits"]), z["logits_sub"]) < 1e-10 and rel(sub(out["encoder_last_hidden_state"]), z["encoder_sub"]) < 1e-10
        for k in WH_GRADS:
            assert rel(sub(g[k]), z["grad::" + k]) < 1e-8, k
        w0 = {k: v.clone() for k, v in w.items()}
        m = {k: torch.zeros_like(v) for k, v in w.items()}
        v_ = {k: torch.zeros_like(v) for k, v in w.items()}
        for t in (1, 2):
            o = O.train_step(ocfg, w, m, v_, t, feats, labels)
            assert abs(float(o["loss"]) - float(z["step_losses"][t - 1])) < 1e-10 * abs(float(z["step_losses"][t - 1]))
        for k in WH_GRADS:
            assert rel(sub(w[k] - w0[k]), z["delta2::" + k]) < 1e-8, k
        # the float32 reference run (TF's own arithmetic: absorption of the score by -1e9, App. C-1) vs the fp64 oracle WITH emulation
        z32, _, ocfg32, w32, feats32, labels32 = whisper_case("ref_whisper_small_cfg_f32.npz", torch.float32)
        whisper_oracle.EMULATE_FP32_ABSORPTION = True
        o64 = O.forward(ocfg32, {k: v.double() for k, v in w32.items()}, feats32.double(), labels32)
        assert abs(float(o64["loss"]) - float(z32["loss"])) < 1e-6 * float(z32["loss"]) and rel(sub(o64["logits"]), z32["logits_sub"]) < 5e-6
    finally:
        whisper_oracle.EMULATE_FP32_ABSORPTION = old
