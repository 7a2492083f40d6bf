#!/usr/bin/env python
"""Per-tensor parity report of the bf16 (benchmarked) CUDA path against the CPU oracle, written as JSON.

    python tools/parity_report.py [--out gpurun_out/parity_report.json] [--cases w2v_tiny,w2v_base_2s,...]

For every case: forward activations, loss and EVERY gradient as relative L2 error ||gpu - oracle|| / ||oracle||
(oracle in fp64 on identical weights/inputs; for Wav2Vec2 the GPU's own VQ code indices are injected into the oracle so
both sides evaluate the same function — the indices themselves are checked bit-exactly elsewhere). This is the data the
per-tensor error budgets in tests/ are derived from; it is a diagnostic, not a test (TEST INFRASTRUCTURE: imports oracle/).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch


def rel_l2(a, b):
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-300))


def w2v_case(size, B, N, precision, seed=0, odtype=torch.float64):
    from oracle import wav2vec2_oracle as O
    from tethys_speech_b200 import wav2vec2 as W

    ocfg = O.Wav2Vec2Config(size)
    w64 = O.randomize_weights(O.init_weights(ocfg, seed=seed, dtype=odtype), seed=seed + 1)
    model = W.Wav2Vec2ForPreTraining(W.Wav2Vec2Config(size), precision=precision, seed=seed)
    model.set_weights({k: v.float() for k, v in w64.items()})
    g = torch.Generator().manual_seed(100 + seed)
    wave = torch.randn(B, N, generator=g, dtype=odtype)
    T = O.num_frames(ocfg, N)
    neg = O.negative_indices_from_random(torch.randint(0, T, (B, T), generator=g), ocfg.num_negatives)
    out = model(wave.float(), training=True, neg_indices=neg, dropout=False)
    grads = model.gradient()
    torch.cuda.synchronize()
    model._prog.ctx.watchdog()
    idx = out["code_indices"].cpu()
    t0 = time.time()
    oout, og = O.loss_and_grads(ocfg, w64, wave, neg, code_indices=idx)
    t_oracle = time.time() - t0
    oout_free = None
    rep = {"oracle_s": t_oracle, "T": T, "fwd": {}, "grads": {}}
    with torch.no_grad():
        free = O.forward(ocfg, w64, wave, neg)
    rep["vq_index_mismatch_vs_fp64_argmin"] = float((free["code_indices"] != idx).double().mean())
    for key in ("extract_features", "last_hidden_state", "projected_states", "quantized_features", "projected_quantized_features",
                "contrastive_logits"):
        rep["fwd"][key] = rel_l2(out[key], oout[key])
    rep["fwd"]["loss"] = abs(float(out["loss"]) - float(oout["loss"])) / abs(float(oout["loss"]))
    rep["fwd"]["perplexity"] = abs(float(out["codevector_perplexity"]) - float(oout["codevector_perplexity"])) / float(oout["codevector_perplexity"])
    gscale = max(float(v.abs().max()) for v in og.values())
    eg = None
    if precision == "bf16":
        from oracle import tf_ops
        with tf_ops.bf16_storage():
            _, eg = O.loss_and_grads(ocfg, w64, wave, neg, code_indices=idx)
    for name, gg in zip(model.variable_names, grads):
        ref = og[name]
        if float(ref.abs().max()) < 1e-12 * max(1.0, gscale):
            rep["grads"][name] = {"zero_ref": True, "gpu_absmax_over_gscale": float(gg.abs().max()) / gscale}
        else:
            rep["grads"][name] = {"rel": rel_l2(gg, ref), "ndim": ref.dim(), "ref_norm": float(ref.norm())}
            if eg is not None:
                rep["grads"][name]["emu"] = rel_l2(eg[name], ref)
    del model
    torch.cuda.empty_cache()
    return rep


def whisper_case(preset, B, Tm, S, precision, seed=0, small=None, odtype=torch.float64):
    from oracle import whisper_oracle as O
    from tethys_speech_b200 import whisper as W

    if small:
        ocfg = O.WhisperConfig("small")
        cfg = W.WhisperConfig()
        for c in (ocfg, cfg):
            c.d_model, c.d_ff = small["d"], small["ff"]
            c.encoder_layers = c.decoder_layers = small["layers"]
            c.encoder_attention_heads = c.decoder_attention_heads = small["heads"]
            c.vocab_size, c.n_mels, c.n_ctx, c.decoder_start_token_id = small["vocab"], small["n_mels"], small["n_ctx"], small["start"]
        model = W.WhisperForConditionalGeneration(cfg, precision=precision, seed=seed)
    else:
        ocfg = O.WhisperConfig(preset)
        model = W.create_whisper_model(preset, precision=precision, seed=seed)
    w64 = O.randomize_weights(O.init_weights(ocfg, seed=seed, dtype=odtype), seed=seed + 1)
    model.set_weights({k: v.float() for k, v in w64.items()})
    g = torch.Generator().manual_seed(7 + seed)
    feats = torch.randn(B, ocfg.n_mels, Tm, generator=g, dtype=odtype)
    labels = (O.dummy_labels(np.random.default_rng(seed), B, S) if S >= 90
              else torch.randint(0, min(100, ocfg.vocab_size), (B, S), generator=g, dtype=torch.int32))
    out = model(feats.float(), labels=labels, training=True, dropout=False)
    grads = model.gradient()
    torch.cuda.synchronize()
    model._prog.ctx.watchdog()
    t0 = time.time()
    oout, og = O.loss_and_grads(ocfg, w64, feats, labels)
    rep = {"oracle_s": time.time() - t0, "fwd": {}, "grads": {}}
    for key in ("encoder_last_hidden_state", "last_hidden_state", "logits"):
        rep["fwd"][key] = rel_l2(out[key], oout[key])
    rep["fwd"]["loss"] = abs(float(out["loss"]) - float(oout["loss"])) / abs(float(oout["loss"]))
    gscale = max(float(v.abs().max()) for v in og.values())
    eg = None
    if precision == "bf16":
        from oracle import tf_ops
        with tf_ops.bf16_storage():
            _, eg = O.loss_and_grads(ocfg, w64, feats, labels)
    for name, gg in zip(model.variable_names, grads):
        ref = og[name]
        if float(ref.abs().max()) < 1e-12 * max(1.0, gscale):
            rep["grads"][name] = {"zero_ref": True, "gpu_absmax_over_gscale": float(gg.abs().max()) / gscale}
        else:
            rep["grads"][name] = {"rel": rel_l2(gg, ref), "ndim": ref.dim(), "ref_norm": float(ref.norm())}
            if eg is not None:
                rep["grads"][name]["emu"] = rel_l2(eg[name], ref)
    del model
    torch.cuda.empty_cache()
    return rep


SMALL = dict(vocab=203, d=128, heads=2, ff=256, layers=2, n_mels=16, n_ctx=64, start=200)
CASES = {
    "w2v_tiny_bf16": lambda: w2v_case("tiny", 2, 3200, "bf16"),
    "w2v_tiny_bf16_seed3": lambda: w2v_case("tiny", 2, 3200, "bf16", seed=3),
    "w2v_tiny_bf16_b4_1s": lambda: w2v_case("tiny", 4, 16000, "bf16", seed=5),
    "w2v_tiny_bf16_T200": lambda: w2v_case("tiny", 1, 8000, "bf16"),
    "w2v_small_bf16_2s": lambda: w2v_case("small", 2, 32000, "bf16"),
    "w2v_base_bf16_2s": lambda: w2v_case("base", 2, 32000, "bf16"),
    "w2v_base_bf16_15s_b1": lambda: w2v_case("base", 1, 240000, "bf16"),
    "w2v_base_fp32_15s_b1": lambda: w2v_case("base", 1, 240000, "fp32"),
    "whisper_smallcfg_bf16": lambda: whisper_case(None, 2, 128, 16, "bf16", small=SMALL),
    "whisper_tiny_bf16": lambda: whisper_case("tiny", 2, 200, 24, "bf16"),
    "whisper_default_bf16_30s_b1": lambda: whisper_case("small", 1, 3000, 100, "bf16"),
    "whisper_default_fp32_30s_b1": lambda: whisper_case("small", 1, 3000, 100, "fp32"),
}


def summarize(rep):
    rels = sorted(((v["rel"], k) for k, v in rep["grads"].items() if "rel" in v), reverse=True)
    ratios = sorted(((v["rel"] / v["emu"], v["rel"], v["emu"], k) for k, v in rep["grads"].items() if "emu" in v), reverse=True)
    over = [(round(r, 2), round(g, 4), round(e, 4), k) for r, g, e, k in ratios if g > 2e-2]
    return {"fwd": rep["fwd"], "worst_grads": rels[:4], "gpu_over_emu_max": ratios[0][:3] if ratios else None,
            "gpu_over_emu_median": ratios[len(ratios) // 2][0] if ratios else None, "over_2e-2_with_ratio": over[:12], "n_over_2e-2": sum(1 for r, _ in rels if r > 2e-2),
            "n_over_1e-2": sum(1 for r, _ in rels if r > 1e-2), "median_grad": rels[len(rels) // 2][0], "n": len(rels),
            "oracle_s": rep["oracle_s"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_report.json"))
    ap.add_argument("--cases", default=",".join(CASES))
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    full = {}
    for name in args.cases.split(","):
        t0 = time.time()
        try:
            full[name] = CASES[name]()
            print(name, json.dumps(summarize(full[name])), f"({time.time() - t0:.1f} s)", flush=True)
        except Exception as e:  # keep going: this is a report
            full[name] = {"error": repr(e)}
            print(name, "ERROR", repr(e), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(full, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
