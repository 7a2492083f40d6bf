"""The CUDA path (fp32 parity mode, through the C-ABI) against the COMMITTED golden vectors of tests/golden/*.npz — loss, integer
outputs bit-exact, logits and a few gradients within 1e-5 relative — without evaluating the oracle's model code at test time
(only its seeded weight initialisers, which the fixtures' generator used too: tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-5


def test_w2v_pretraining_step_matches_golden():
    from oracle import wav2vec2_oracle as WO
    from tethys_speech_b200 import wav2vec2 as W

    z = np.load(os.path.join(GOLD, "w2v_tiny.npz"))
    seed = int(z["seed"])
    w64 = WO.randomize_weights(WO.init_weights(WO.Wav2Vec2Config("tiny"), seed=seed, dtype=torch.float64), seed=seed + 1)
    model = W.Wav2Vec2ForPreTraining(W.Wav2Vec2Config("tiny"), precision="fp32", device=0)
    model.set_weights({k: v.float() for k, v in w64.items()})
    out = model(torch.from_numpy(z["wave"]).float(), training=True, neg_indices=torch.from_numpy(z["neg"]), dropout=False)
    grads = dict(zip(model.variable_names, model.gradient()))
    assert np.array_equal(out["code_indices"].cpu().numpy(), z["code_indices"])            # integer work: bit-exact
    assert abs(float(out["loss"]) - float(z["loss"])) <= TOL * abs(float(z["loss"]))
    assert rel_l2(out["contrastive_logits"][:, ::7, ::9], torch.from_numpy(z["logits_sub"])) < TOL
    for k in ("fe.conv0.kernel", "encoder.layers.3.feed_forward.output_dense.bias", "quantizer.codevectors"):
        assert rel_l2(grads[k], torch.from_numpy(z["grad::" + k])) < 3 * TOL, k
    model._prog.ctx.watchdog()


def test_whisper_step_matches_golden():
    from test_oracle_crosscheck import _generate_case, _whisper_case
    from tethys_speech_b200 import whisper as WH

    z = np.load(os.path.join(GOLD, "whisper_small_cfg.npz"))
    ocfg, w64, _, _ = _whisper_case(seed=int(z["seed"]))
    cfg = WH.WhisperConfig()
    for name in ("d_model", "d_ff", "encoder_layers", "decoder_layers", "encoder_attention_heads", "decoder_attention_heads",
                 "vocab_size", "n_mels", "n_ctx", "decoder_start_token_id"):
        setattr(cfg, name, getattr(ocfg, name))
    model = WH.WhisperForConditionalGeneration(cfg, precision="fp32", device=0)
    model.set_weights({k: v.float() for k, v in w64.items()})
    out = model(torch.from_numpy(z["feats"]).float(), labels=torch.from_numpy(z["labels"]), training=True, dropout=False)
    grads = dict(zip(model.variable_names, model.gradient()))
    assert abs(float(out["loss"]) - float(z["loss"])) <= TOL * abs(float(z["loss"]))
    assert rel_l2(out["logits"][:, ::2, ::5], torch.from_numpy(z["logits_sub"])) < TOL
    for k in ("encoder.conv1.kernel", "lm_head.kernel", "decoder.layers.0.self_attn.k_proj.kernel"):
        assert rel_l2(grads[k], torch.from_numpy(z["grad::" + k])) < 3 * TOL, k
    # generate(): the fixture's greedy token ids, bit-exact
    z2 = np.load(os.path.join(GOLD, "heads_generate.npz"))
    _, w64, _ = _generate_case(seed=int(z2["gen_seed"]))
    model.set_weights({k: v.detach().float() for k, v in w64.items()})
    ids = model.generate(torch.from_numpy(z2["feats"]).float(), max_length=int(z2["max_length"]))
    assert np.array_equal(ids.cpu().numpy().astype(np.int64), z2["generate::ids"])
    model._prog.ctx.watchdog()


@pytest.mark.parametrize("model_type,head", [("asr", "ctc"), ("classification", "classification")])
def test_task_heads_match_golden(model_type, head):
    from oracle import wav2vec2_oracle as WO
    from tethys_speech_b200 import wav2vec2 as W

    z = np.load(os.path.join(GOLD, "heads_generate.npz"))
    seed = int(z["seed"])
    w64 = WO.randomize_weights(WO.init_head_weights(WO.Wav2Vec2Config("tiny"), head, seed=seed, dtype=torch.float64), seed=seed + 1)
    model = W.create_full_model(model_type, "tiny", precision="fp32", device=0)
    model.set_weights({k: v.float() for k, v in w64.items()})
    out = model(torch.from_numpy(z["wave"]).float(), labels=torch.from_numpy(z["labels"]), training=True, dropout=False)
    grads = dict(zip(model.variable_names, model.gradient()))
    assert abs(float(out["loss"]) - float(z[head + "::loss"])) <= TOL * abs(float(z[head + "::loss"]))
    assert rel_l2(out["logits"].reshape(-1)[::3], torch.from_numpy(z[head + "::logits_sub"])) < TOL
    k = "lm_head.kernel" if head == "ctc" else "classifier_proj.kernel"
    for name in (k, "fe.conv0.kernel"):
        assert rel_l2(grads[name], torch.from_numpy(z[head + "::grad::" + name])) < 3 * TOL, name
    model._prog.ctx.watchdog()
