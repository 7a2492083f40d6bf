"""Dropout (tf.keras.layers.Dropout, W:160, W:203-205, W:342; V:281, V:393-396, V:431) is a counter-based mask that every kernel
regenerates from (seed, flat element index) instead of storing it. What has to hold:
  * the GEMM epilogues (tcgen05 engine: whole-chunk path and element path; CUDA-core engine) and the element-wise kernel
    produce the SAME mask for the same (seed, index) — otherwise forward and backward of a layer disagree silently;
  * the keep rate is 1 - rate at every position of the 32-element generator chunk, and seeds decorrelate;
  * with dropout ON the analytic gradient is the derivative of the dropped loss (central differences along the gradient
    direction, fp32 mode) — this walks every dropout site of a model, forward and backward;
  * bf16 mode (tcgen05 epilogues) draws the same masks as fp32 mode (CUDA-core kernels): same loss / gradients up to bf16."""
import ctypes as C

import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _ctx():
    from tethys_speech_b200 import _lib
    return _lib, _lib.context(0)


def _dropout(x, rate, seed):
    from tethys_speech_b200.runtime import stream_ptr
    _lib, ctx = _ctx()
    y = torch.empty_like(x)
    dt = _lib.TS_BF16 if x.dtype == torch.bfloat16 else _lib.TS_F32
    ctx.check(ctx.lib.ts_dropout(ctx.h, dt, x.data_ptr(), y.data_ptr(), x.numel(), rate, seed, stream_ptr()))
    return y


def _gemm_identity_dropout(x, rate, seed, engine):
    """y = dropout(x @ I): the epilogue's mask over the flat [m, n] output."""
    from tethys_speech_b200.runtime import stream_ptr
    _lib, ctx = _ctx()
    m, n = x.shape
    eye = torch.eye(n, dtype=x.dtype, device=x.device)
    y = torch.empty_like(x)
    d = _lib.GemmDesc()
    d.a, d.b, d.c = x.data_ptr(), eye.data_ptr(), y.data_ptr()
    d.m, d.n, d.k, d.a_major, d.b_major = m, n, n, 0, 1
    d.lda = d.ldb = d.ldc = n
    d.batch1 = d.batch2 = 1
    d.in_dtype = d.out_dtype = _lib.TS_BF16 if x.dtype == torch.bfloat16 else _lib.TS_F32
    d.alpha, d.drop, d.seed, d.force_engine = 1.0, rate, seed, engine
    ctx.check(ctx.lib.ts_gemm(ctx.h, C.byref(d), stream_ptr()))
    torch.cuda.synchronize()
    ctx.watchdog()
    return y


@pytest.mark.parametrize("m,n,dtype,engine", [
    (300, 256, torch.bfloat16, 2),    # tcgen05 engine, row stride a multiple of 32: whole-chunk path of the epilogue
    (300, 3072, torch.bfloat16, 2),   # the FFN width (256-column tiles)
    (200, 72, torch.bfloat16, 2),     # row stride 72: rows start mid-chunk -> element path
    (130, 96, torch.float32, 1),      # CUDA-core engine (fp32 parity mode)
    (130, 96, torch.bfloat16, 1),
])
def test_gemm_epilogue_mask_is_the_elementwise_mask(m, n, dtype, engine):
    g = torch.Generator().manual_seed(m + n)
    x = (torch.randn(m, n, generator=g) + 3.0).to(dtype).cuda()      # no zeros in the input: zeros in the output = dropped
    for seed in (7, 0x123456789ABCDEF):
        a = _gemm_identity_dropout(x, 0.1, seed, engine)
        b = _dropout(x, 0.1, seed)
        assert torch.equal(a == 0, b == 0), f"masks differ at {int(((a == 0) != (b == 0)).sum())} of {a.numel()} positions"
        assert torch.equal(a, b)                                       # and the kept values carry the same 1/(1-rate)
        kept = float((b != 0).float().mean())
        assert abs(kept - 0.9) < 0.01, kept


def test_mask_statistics_per_chunk_position_and_seed_independence():
    n = 32 * 40000
    x = torch.ones(n, device="cuda")
    y1 = _dropout(x, 0.1, 11)
    y2 = _dropout(x, 0.1, 12)
    assert torch.equal(y1, _dropout(x, 0.1, 11))                       # a pure function of (seed, index)
    k1, k2 = (y1 != 0).float(), (y2 != 0).float()
    assert abs(float(y1.mean()) - 1.0) < 5e-3                          # unbiased: E[mask / (1 - rate)] = 1
    assert bool(((y1 == 0) | ((y1 - 1 / 0.9).abs() < 1e-6)).all())
    per_pos = k1.view(-1, 32).mean(dim=0)                              # keep rate at each of the 32 positions of a chunk
    assert float((per_pos - 0.9).abs().max()) < 0.008, per_pos
    # neighbours inside a chunk are uncorrelated (the generator is an LCG within a chunk) and so are two seeds
    a, b = k1.view(-1, 32)[:, :-1].reshape(-1), k1.view(-1, 32)[:, 1:].reshape(-1)
    for u, v in ((a, b), (k1, k2)):
        corr = float(((u - u.mean()) * (v - v.mean())).mean() / (u.std() * v.std()))
        assert abs(corr) < 5e-3, corr
    for rate in (0.05, 0.5):
        assert abs(float((_dropout(x, rate, 3) != 0).float().mean()) - (1 - rate)) < 3e-3
    assert torch.equal(_dropout(x, 0.0, 3), x)


def _directional_check(loss_at, params, grads, eps, tol):
    """central difference of the loss along the (per-tensor normalised) gradient direction vs the analytic g . d"""
    d = {k: g / (g.norm() + 1e-30) for k, g in grads.items()}
    analytic = sum(float((grads[k].double() * d[k].double()).sum()) for k in grads)
    lp = loss_at({k: params[k] + eps * d[k] for k in params})
    lm = loss_at({k: params[k] - eps * d[k] for k in params})
    numeric = (lp - lm) / (2 * eps)
    assert abs(numeric - analytic) <= tol * abs(analytic), (numeric, analytic)
    return numeric, analytic


def test_w2v_ctc_fp32_gradient_is_the_derivative_of_the_dropped_loss():
    """CTC-head model (no hard quantiser on the loss path): every dropout site of the trunk + head, forward and backward."""
    from tethys_speech_b200 import wav2vec2 as W

    model = W.create_full_model("asr", "tiny", precision="fp32", device=0, seed=3)
    g = torch.Generator().manual_seed(8)
    wave = torch.randn(2, 3200, generator=g).cuda()
    step = 4242

    def forward(weights=None):
        if weights is not None:
            model.set_weights(weights)
        model._step_seed = step - 1                      # the same dropout masks at every evaluation
        return model(wave, labels=torch.zeros(2), training=True, dropout=True)

    w0 = model.get_weights()
    out = forward()
    loss_drop = float(out["loss"])
    grads = {k: v.clone() for k, v in zip(model.variable_names, model.gradient())}
    loss_nodrop = float(model(wave, labels=torch.zeros(2), training=True, dropout=False)["loss"])
    assert abs(loss_drop - loss_nodrop) > 1e-4 * abs(loss_nodrop)          # dropout really was on
    grads = {k: v for k, v in grads.items() if float(v.abs().max()) > 0}   # the quantiser's variables get none
    params = {k: w0[k] for k in grads}
    num, ana = _directional_check(lambda w: float(forward(w)["loss"]), params, grads, eps=2e-3, tol=2e-2)
    print(f"w2v-ctc fp32 dropout on: numeric {num:.6f} analytic {ana:.6f}")
    model._prog.ctx.watchdog()


def test_whisper_fp32_gradient_is_the_derivative_of_the_dropped_loss():
    from tethys_speech_b200 import whisper as WH

    cfg = WH.WhisperConfig()
    cfg.d_model, cfg.encoder_layers, cfg.decoder_layers, cfg.d_ff = 64, 2, 2, 128
    cfg.encoder_attention_heads = cfg.decoder_attention_heads = 2
    cfg.vocab_size, cfg.n_mels, cfg.n_ctx, cfg.decoder_start_token_id = 203, 16, 64, 200
    cfg.activation_dropout = 0.1                          # the reference has 0.0 here (W:31): exercise that site too
    model = WH.WhisperForConditionalGeneration(cfg, precision="fp32", device=0, seed=2)
    g = torch.Generator().manual_seed(5)
    feats = torch.randn(2, cfg.n_mels, 100, generator=g).cuda()
    labels = torch.randint(3, 100, (2, 12), generator=g, dtype=torch.int32).cuda()
    step = 777

    def forward(weights=None):
        if weights is not None:
            model.set_weights(weights)
        model._step_seed = step - 1
        return model(feats, labels=labels, training=True, dropout=True)

    w0 = model.get_weights()
    loss_drop = float(forward()["loss"])
    grads = {k: v.clone() for k, v in zip(model.variable_names, model.gradient())}
    loss_nodrop = float(model(feats, labels=labels, training=True, dropout=False)["loss"])
    assert abs(loss_drop - loss_nodrop) > 1e-4 * abs(loss_nodrop)
    grads = {k: v for k, v in grads.items() if float(v.abs().max()) > 0}
    params = {k: w0[k] for k in grads}
    num, ana = _directional_check(lambda w: float(forward(w)["loss"]), params, grads, eps=2e-3, tol=2e-2)
    print(f"whisper fp32 dropout on: numeric {num:.6f} analytic {ana:.6f}")
    model._prog.ctx.watchdog()


def test_bf16_mode_draws_the_same_masks_as_fp32_mode():
    """Same weights, same step seed, dropout on everywhere but inside attention (the fused bf16 attention kernels draw their
    probabilities' mask per (batch, head) stream, the unfused fp32 path per flat index): the tcgen05 epilogues and the bf16
    element-wise kernels must reproduce the fp32 run up to bf16 rounding — a different mask would move the loss by percents."""
    from tethys_speech_b200 import wav2vec2 as W

    outs = {}
    for precision in ("fp32", "bf16"):
        cfg = W.Wav2Vec2Config("tiny")
        cfg.attention_dropout = 0.0
        cfg.hidden_dropout, cfg.activation_dropout = 0.2, 0.2
        model = W.Wav2Vec2ForCTC(cfg, precision=precision, device=0, seed=6)
        if "w" in outs:
            model.set_weights(outs["w"])
        else:
            outs["w"] = model.get_weights()
        wave = torch.randn(2, 3200, generator=torch.Generator().manual_seed(9)).cuda()
        model._step_seed = 99
        o = model(wave, labels=torch.zeros(2), training=True, dropout=True)
        grads = {k: v.clone() for k, v in zip(model.variable_names, model.gradient())}
        outs[precision] = (float(o["loss"]), o["logits"].float().clone(), grads)
        if precision == "fp32":
            model._step_seed = 99
            nodrop = float(model(wave, labels=torch.zeros(2), training=True, dropout=False)["loss"])
            assert abs(nodrop - outs["fp32"][0]) > 1e-3 * abs(nodrop)
    (l32, lg32, g32), (l16, lg16, g16) = outs["fp32"], outs["bf16"]
    assert abs(l16 - l32) < 2e-2 * abs(l32), (l16, l32)
    assert rel_l2(lg16, lg32) < 3e-2
    for k in ("lm_head.kernel", "encoder.layers.0.feed_forward.intermediate_dense.kernel", "encoder.layers.3.attention.out_proj.kernel",
              "feature_projection.kernel", "fe.conv1.kernel"):
        assert rel_l2(g16[k], g32[k]) < 8e-2, (k, rel_l2(g16[k], g32[k]))
