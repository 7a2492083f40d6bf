"""GPU parity of the two fine-tuning heads on the Wav2Vec2 trunk against the CPU oracle:
Wav2Vec2ForCTC (V:940-1001, stand-in loss = mean CE against class 0) and Wav2Vec2ForSequenceClassification
(V:1004-1070), reached through create_full_model(model_type='asr' | 'classification') and the VS:1119-1176 train step.
fp32 mode 1e-5 relative, bf16 mode 2e-2 (logits, loss, gradients; a gradient's bar is lifted only to 1.5 x the error an
independent CPU emulation of bf16 storage shows for that tensor, conftest.check_bf16_grads); dropout off in parity runs."""
import pytest
import torch

from conftest import BF16_TOL, check_bf16_grads, rel_l2

pytestmark = pytest.mark.gpu

HEAD = {"asr": "ctc", "classification": "classification"}


def _setup(model_type, size, B, N, precision, seed=0):
    from oracle import wav2vec2_oracle as O
    from tethys_speech_b200 import wav2vec2 as W

    ocfg = O.Wav2Vec2Config(size)
    w64 = O.randomize_weights(O.init_head_weights(ocfg, HEAD[model_type], seed=seed, dtype=torch.float64), seed=seed + 1)
    model = W.create_full_model(model_type, size, precision=precision, device=0, seed=seed)
    assert set(model.variable_names) == set(w64), set(model.variable_names) ^ set(w64)
    model.set_weights({k: v.float() for k, v in w64.items()})
    g = torch.Generator().manual_seed(200 + seed)
    wave = torch.randn(B, N, generator=g, dtype=torch.float64)
    labels = torch.randint(0, ocfg.num_labels, (B,), generator=g)
    return O, ocfg, w64, model, wave, labels


def _check(model_type, size, B, N, precision, tol, grad_tol=None):
    O, ocfg, w64, model, wave, labels = _setup(model_type, size, B, N, precision)
    out = model(wave.float(), labels=labels, training=True, dropout=False)
    grads = model.gradient()
    torch.cuda.synchronize()
    model._prog.ctx.watchdog()
    oout, og = O.head_loss_and_grads(ocfg, w64, wave, labels, HEAD[model_type])
    assert tuple(out["logits"].shape) == tuple(oout["logits"].shape)
    errs = {"logits": rel_l2(out["logits"], oout["logits"]),
            "loss": abs(float(out["loss"]) - float(oout["loss"])) / abs(float(oout["loss"]))}
    bad = {k: v for k, v in errs.items() if not v <= tol}
    assert not bad, f"forward mismatch (tol {tol}): {errs}"
    gerrs = {}
    gscale = max(float(v.abs().max()) for v in og.values())
    for name, g in zip(model.variable_names, grads):
        ref = og[name]
        if float(ref.abs().max()) == 0.0:                      # the quantizer's variables: None -> zeros (VS:1163-1166)
            assert name.startswith("quantizer.") and float(g.abs().max()) == 0.0, name
            continue
        if float(ref.abs().max()) < 1e-12 * max(gscale, 1.0):  # mathematically zero (key bias): rounding noise only
            assert float(g.abs().max()) < (1e-5 if precision == "fp32" else 2e-2) * gscale, name
            continue
        gerrs[name] = rel_l2(g, ref)
    worst = sorted(gerrs.items(), key=lambda kv: -kv[1])[:4]
    print(f"[{model_type} {size} {precision}] fwd {errs}; worst grads {worst}")
    if precision == "fp32":
        badg = {k: v for k, v in gerrs.items() if not v <= (grad_tol if og[k].dim() > 1 else 3 * grad_tol)}
        assert not badg, f"gradient mismatch (tol {grad_tol}): {len(badg)} tensors; worst {worst}"
    else:
        from oracle import tf_ops

        with tf_ops.bf16_storage():
            _, eg = O.head_loss_and_grads(ocfg, w64, wave, labels, HEAD[model_type])
        badg = check_bf16_grads(f"{model_type} {size} bf16", gerrs, {k: rel_l2(eg[k], og[k]) for k in gerrs})
        assert not badg, f"bf16 gradients over budget: {badg}"


def test_ctc_head_tiny_fp32():
    _check("asr", "tiny", 2, 3200, "fp32", 1e-5, 1e-5)


def test_classification_head_tiny_fp32():
    _check("classification", "tiny", 3, 3333, "fp32", 1e-5, 1e-5)


def test_ctc_head_tiny_bf16():
    _check("asr", "tiny", 2, 3200, "bf16", BF16_TOL)


def test_classification_head_base_bf16():
    _check("classification", "base", 2, 16000, "bf16", BF16_TOL)


@pytest.mark.parametrize("model_type", ["asr", "classification"])
def test_head_train_steps_fp32_match_oracle_adam(model_type):
    """VS:1119-1176 with a task head: 2 steps of clip_by_global_norm(1.0) + clipnorm 1.0 + Adam(3e-5, eps 1e-8); the
    quantizer's variables get zero gradients and must not move."""
    from tethys_speech_b200 import wav2vec2 as W
    from tethys_speech_b200.runtime import Adam

    O, ocfg, w64, model, wave, labels = _setup(model_type, "tiny", 2, 3200, "fp32", seed=4)
    opt = Adam(learning_rate=3e-5, epsilon=1e-8, clipnorm=1.0)
    w = {k: v.clone() for k, v in w64.items()}
    m = {k: torch.zeros_like(v) for k, v in w.items()}
    v = {k: torch.zeros_like(x) for k, x in w.items()}
    cb0 = w64["quantizer.codevectors"].clone()
    for t in (1, 2, 3):
        loss = W.train_step(model, (wave.float(), labels), opt, dropout=False)
        oout = O.head_train_step(ocfg, w, m, v, t, wave, labels, HEAD[model_type], lr=3e-5, eps=1e-8)
        assert abs(float(loss) - float(oout["loss"])) <= 1e-4 * abs(float(oout["loss"])), (t, float(loss), float(oout["loss"]))
    got = model.get_weights()
    # an Adam step moves every element by ~lr whatever the gradient's size: compare the accumulated change (tensors whose
    # gradient is pure rounding noise, e.g. the key bias, are left out)
    head_vars = ["lm_head.kernel", "lm_head.bias"] if model_type == "asr" else \
        ["classifier_proj.kernel", "classifier_proj.bias", "classifier.kernel", "classifier.bias"]
    for k in head_vars + ["encoder.layers.0.attention.q_proj.kernel", "fe.conv1.kernel", "fe.conv0.gn.gamma",
                          "encoder.layers.1.feed_forward.output_dense.bias"]:
        d_gpu = got[k].double().cpu() - w64[k]
        d_ref = w[k] - w64[k]
        assert rel_l2(d_gpu, d_ref) < 2e-3, (k, rel_l2(d_gpu, d_ref))
    for k in ("quantizer.projection.kernel", "quantizer.projection.bias"):
        assert torch.equal(got[k].cpu(), w64[k].float()), k
    assert torch.equal(got["quantizer.codevectors"].cpu(), cb0.float())
    model._prog.ctx.watchdog()


def test_head_inference_call_and_errors():
    """model(x, training=False): logits only, no loss (V:982-985); labels of the wrong length and pre-training-only buffers fail loudly."""
    from tethys_speech_b200 import TethysError
    from tethys_speech_b200 import wav2vec2 as W

    model = W.create_full_model("classification", "tiny", precision="bf16", device=0)
    x = torch.randn(2, 3200)
    out = model(x, training=False)
    assert out["loss"] is None and tuple(out["logits"].shape) == (2, model.config.num_labels)
    a = out["logits"].clone()
    b = model(x, labels=torch.zeros(2), training=False)["logits"]     # labels without training: still no loss, same logits
    assert torch.equal(a, b)
    with pytest.raises(ValueError):
        model(x, labels=torch.zeros(3), training=True)
    with pytest.raises(TethysError):
        model._prog.buffer("projected_states")
    with pytest.raises(TethysError):
        model.gradient()                                               # no training forward -> no backward state
    ctc = W.create_full_model("asr", "tiny", precision="bf16", device=0)
    o = ctc(x, labels=torch.zeros(2), training=True)
    T = ctc.num_frames(3200)
    assert tuple(o["logits"].shape) == (2, T, ctc.config.vocab_size) and float(o["loss"]) > 0
    with pytest.raises(NotImplementedError):
        W.create_full_model("something_else", "tiny")


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_ctc_loss_operator_matches_oracle(dtype):
    """ts_ctc_loss (tf.nn.ctc_loss of WS:897-929) against the oracle: per-sample losses and d loss / d logits, with a repeated label,
    padded label rows, an empty transcript, an impossible alignment (inf, zero gradient; 0 under zero_infinity) and T = 750."""
    import ctypes as C  # noqa: F401

    from oracle import wav2vec2_oracle as O
    from tethys_speech_b200 import _lib
    from tethys_speech_b200.runtime import ptr, stream_ptr

    ctx = _lib.context(0)
    g = torch.Generator().manual_seed(21)
    for (B, Tn, V, L) in ((4, 50, 32, 12), (2, 750, 32, 100), (3, 6, 5, 4)):
        logits = torch.randn(B, Tn, V, generator=g, dtype=torch.float64) * 2
        labels = torch.randint(1, V, (B, L), generator=g)
        for b in range(B):
            n = int(torch.randint(0, L + 1, (1,), generator=g)) if b else L
            labels[b, n:] = 0
        if B > 2:
            labels[1, :2] = labels[1, 0]                      # a repeat
            labels[2, :] = 0                                  # empty transcript
        if Tn == 6:
            labels[0] = torch.tensor([1, 1, 2, 2])            # needs 6 frames exactly; row 1 below cannot be aligned in 6
            labels[1] = torch.tensor([3, 3, 3, 3])
        lg = logits.clone().requires_grad_(True)
        loss, per = O.ctc_loss(lg, labels, reduction="sum", zero_infinity=True)
        loss.backward()
        want_g = lg.grad
        _, per_raw = O.ctc_loss(logits, labels)
        d_lg = logits.float().cuda().contiguous()
        d_lab = labels.to(torch.int32).cuda().contiguous()
        ws = torch.empty(int(ctx.lib.ts_ctc_workspace_floats(B, Tn, L)), device="cuda")
        per_gpu = torch.empty(B, device="cuda")
        dl = torch.empty(B, Tn, V, device="cuda", dtype=torch.float32 if dtype == "fp32" else torch.bfloat16)
        for zi in (0, 1):
            ctx.check(ctx.lib.ts_ctc_loss(ctx.h, _lib.TS_F32 if dtype == "fp32" else _lib.TS_BF16, ptr(d_lg), ptr(d_lab), B, Tn, V, L, 0, ptr(ws),
                                          ptr(per_gpu), ptr(dl), 1.0, zi, stream_ptr()))
            ref = (per if zi else per_raw).detach()
            got = per_gpu.cpu().double()
            fin = torch.isfinite(ref)
            assert bool((torch.isinf(got) == torch.isinf(ref)).all())
            assert float(((got[fin] - ref[fin]).abs() / ref[fin].abs().clamp_min(1e-6)).max()) < 2e-5, (B, Tn, got, ref)
        assert rel_l2(dl, want_g) < (2e-5 if dtype == "fp32" else 4e-3)      # bf16: one rounding of each gradient element


def test_ctc_model_real_loss_and_gradients_fp32():
    """Wav2Vec2ForCTC with a transcript (labels [B, L]) takes the real CTC loss of the legacy file (WS:897-929, reduction "sum"):
    loss and every gradient of the train-step body against the oracle (autograd through the restated trunk + ctc_loss)."""
    from collections import OrderedDict

    from oracle import wav2vec2_oracle as O
    from tethys_speech_b200 import wav2vec2 as W

    ocfg = O.Wav2Vec2Config("tiny")
    w0 = O.randomize_weights(O.init_head_weights(ocfg, "ctc", seed=0, dtype=torch.float64), seed=1)
    model = W.Wav2Vec2ForCTC(W.Wav2Vec2Config("tiny"), precision="fp32", seed=0)
    model.set_weights({k: v.float() for k, v in w0.items()})
    g = torch.Generator().manual_seed(31)
    wave = torch.randn(2, 6400, generator=g, dtype=torch.float64)
    labels = torch.randint(1, ocfg.vocab_size, (2, 6), generator=g)
    labels[1, 4:] = 0
    out = model(wave.float(), labels=labels, training=True, dropout=False)
    model.gradient()
    ws = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in w0.items())
    oo = O.forward_head(ocfg, ws, wave, None, "ctc")
    loss, per = O.ctc_loss(oo["logits"], labels, reduction="sum")
    grads = torch.autograd.grad(loss, list(ws.values()), allow_unused=True)
    assert abs(float(out["loss"]) - float(loss)) < 1e-5 * abs(float(loss))
    prog = model._prog
    worst = 0.0
    for (k, v), gi in zip(ws.items(), grads):
        if gi is None or float(gi.abs().max()) < 1e-12:
            continue
        e = rel_l2(prog.view(prog.grads, k), gi)
        worst = max(worst, e / (1.0 if gi.dim() > 1 else 3.0))
    assert worst < 1e-5, worst
    # the dummy dataset's per-clip label keeps the reference's stand-in loss (V:994-1000)
    out2 = model(wave.float(), labels=torch.zeros(2), training=True, dropout=False)
    assert abs(float(out2["loss"]) - float(oo["loss"])) < 1e-5 * abs(float(oo["loss"]))
