"""GPU: the CUDA path (through the C-ABI) against the golden vectors produced by the REFERENCE'S OWN CODE (tests/golden/ref_*.npz,
see tests/golden/make_ref_golden.py): fp32 parity mode within 1e-5 relative of the reference's float64 run (3e-5 for 1-D
column-sum gradients, 2e-3 on the two-step weight CHANGE, which Adam's sign-like first updates amplify), VQ code indices
bit-exact. Nothing of the oracle's model code runs here — only its seeded weight initialisers, which the generator used too."""
import numpy as np
import pytest
import torch

from test_ref_golden import W2V_GRADS, WH_GRADS, rel, sub, w2v_case, whisper_case, whisper_edit

pytestmark = pytest.mark.gpu
TOL = 1e-5


def test_cuda_w2v_step_matches_the_reference_vectors():
    from tethys_speech_b200 import wav2vec2 as W

    z, O, ocfg, w, wave, neg, step_negs = w2v_case()
    model = W.Wav2Vec2ForPreTraining(W.Wav2Vec2Config(str(z["size"])), precision="fp32", device=0)
    model.set_weights({k: v.float() for k, v in w.items()})
    out = model(wave.float(), training=True, neg_indices=neg, dropout=False)
    grads = dict(zip(model.variable_names, model.gradient()))
    assert np.array_equal(out["code_indices"].cpu().numpy(), z["code_indices"])                # integer work: bit-exact
    assert abs(float(out["loss"]) - float(z["loss"])) <= TOL * abs(float(z["loss"]))
    assert abs(float(out["codevector_perplexity"]) - float(z["perplexity"])) <= TOL * float(z["perplexity"])
    for key, okey in (("logits_sub", "contrastive_logits"), ("last_hidden_sub", "last_hidden_state"), ("extract_features_sub", "extract_features")):
        assert rel(sub(out[okey]), z[key]) < TOL, (key, rel(sub(out[okey]), z[key]))
    for k in W2V_GRADS:
        assert rel(sub(grads[k]), z["grad::" + k]) < (TOL if grads[k].dim() > 1 else 3 * TOL), (k, rel(sub(grads[k]), z["grad::" + k]))
    opt = W.Adam(learning_rate=3e-5, epsilon=1e-8, clipnorm=1.0)
    before = model.get_weights()
    for t in (1, 2):
        loss = W.train_step(model, (wave.float(), None), opt, neg_indices=step_negs[t - 1], dropout=False)
        assert abs(float(loss) - float(z["step_losses"][t - 1])) <= 1e-4 * abs(float(z["step_losses"][t - 1]))
    after = model.get_weights()
    for k in W2V_GRADS:
        assert rel(sub(after[k].double() - before[k].double()), z["delta2::" + k]) < 2e-3, k
    model._prog.ctx.watchdog()


def test_cuda_whisper_step_matches_the_reference_float32_vectors():
    """The reference's float32 run carries TF's absorption quirk (App. C-1) natively; the CUDA kernels add the -1e9 mask
    literally in fp32, so they must reproduce it without any emulation."""
    from tethys_speech_b200 import whisper as W

    z, O, ocfg, w, feats, labels = whisper_case("ref_whisper_small_cfg_f32.npz", torch.float32)
    cfg = W.WhisperConfig()
    whisper_edit(cfg)
    model = W.WhisperForConditionalGeneration(cfg, precision="fp32", device=0)
    model.set_weights({k: v.float() for k, v in w.items()})
    out = model(feats.float(), labels=labels, training=True, dropout=False)
    torch.cuda.synchronize()
    assert abs(float(out["loss"]) - float(z["loss"])) <= TOL * abs(float(z["loss"]))
    assert rel(sub(out["logits"]), z["logits_sub"]) < TOL
    model._prog.ctx.watchdog()
