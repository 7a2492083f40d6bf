#!/usr/bin/env python
"""Generates the committed golden vectors (tests/golden/*.npz) from the CPU oracle in fp64.

The reference (hyunnnchoi/tethys-speech) has no tests, fixtures or seeds and TensorFlow cannot be installed here, so these
vectors pin the *oracle* (against silent edits), not the reference: parity stays "unpinned" in the sense of SURVEY §8c.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import wav2vec2_oracle as WO  # noqa: E402
from oracle import whisper_oracle as HO  # noqa: E402
from test_oracle_crosscheck import _generate_case, _head_case, _w2v_case, _whisper_case  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    seed = 3
    cfg, w, wave, neg = _w2v_case(seed=seed)
    out, g = WO.loss_and_grads(cfg, w, wave, neg)
    np.savez_compressed(
        os.path.join(OUT, "w2v_tiny.npz"), seed=seed, wave=wave.numpy(), neg=neg.numpy(), loss=float(out["loss"]),
        code_indices=out["code_indices"].numpy(), logits_sub=out["contrastive_logits"].detach().numpy()[:, ::7, ::9],
        **{"grad::" + k: g[k].numpy() for k in ("fe.conv0.kernel", "encoder.layers.3.feed_forward.output_dense.bias", "quantizer.codevectors")})
    cfg, w, feats, labels = _whisper_case(seed=seed)
    out, g = HO.loss_and_grads(cfg, w, feats, labels)
    np.savez_compressed(
        os.path.join(OUT, "whisper_small_cfg.npz"), seed=seed, feats=feats.numpy(), labels=labels.numpy(), loss=float(out["loss"]),
        logits_sub=out["logits"].detach().numpy()[:, ::2, ::5],
        **{"grad::" + k: g[k].numpy() for k in ("encoder.conv1.kernel", "lm_head.kernel", "decoder.layers.0.self_attn.k_proj.kernel")})
    # f-2: the task heads on the Wav2Vec2 trunk and greedy generate()
    rec = {"seed": seed}
    for head in ("ctc", "classification"):
        cfg, w, wave, labels = _head_case(head, seed=seed)
        out, g = WO.head_loss_and_grads(cfg, w, wave, labels, head)
        rec["wave"], rec["labels"] = wave.numpy(), labels.numpy()
        rec[head + "::loss"] = float(out["loss"])
        rec[head + "::logits_sub"] = out["logits"].detach().numpy().reshape(-1)[::3]
        k = "lm_head.kernel" if head == "ctc" else "classifier_proj.kernel"
        rec[head + "::grad::" + k] = g[k].numpy()
        rec[head + "::grad::fe.conv0.kernel"] = g["fe.conv0.kernel"].numpy()
    rec["gen_seed"] = 28
    cfg, w, feats = _generate_case(seed=28)
    rec["feats"], rec["max_length"] = feats.numpy(), 12
    with torch.no_grad():
        rec["generate::ids"] = HO.generate(cfg, w, feats, max_length=12).numpy()
    np.savez_compressed(os.path.join(OUT, "heads_generate.npz"), **rec)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
