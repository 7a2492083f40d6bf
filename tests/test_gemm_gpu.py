"""K9 — ts_gemm (tcgen05 engine for bf16 operands, CUDA-core engine for fp32) against plain PyTorch fp32 on the same
(bf16-rounded) inputs: Dense forward with the fused epilogue (W:194-205 / V:383-398), strided-conv window rows
(V:254-268, no im2col), weight gradient with split-K accumulation, and the GELU-backward epilogue."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm(a, b, c, m, n, k, a_major, b_major, lda, ldb, in_dt, out_dt, **kw):
    from tethys_speech_b200 import _lib
    from tethys_speech_b200.runtime import stream_ptr

    ctx = _lib.context(0)
    d = _lib.GemmDesc()
    d.a, d.b, d.c = a.data_ptr(), b.data_ptr(), c.data_ptr()
    d.m, d.n, d.k, d.a_major, d.b_major = m, n, k, a_major, b_major
    d.lda, d.ldb, d.ldc = lda, ldb, n
    d.batch1 = d.batch2 = 1
    d.in_dtype, d.out_dtype, d.alpha = in_dt, out_dt, kw.get("alpha", 1.0)
    for key in ("bias", "residual", "c_preact", "act_aux"):
        if kw.get(key) is not None:
            setattr(d, key, kw[key].data_ptr())
    d.ldr = d.ld_aux = n
    d.act, d.accumulate, d.drop, d.seed = kw.get("act", 0), kw.get("accumulate", 0), kw.get("drop", 0.0), 5
    ctx.check(ctx.lib.ts_gemm(ctx.h, C.byref(d), stream_ptr()))
    torch.cuda.synchronize()
    ctx.watchdog()


def _rel(x, ref):
    return float((x.float() - ref).norm() / ref.norm())


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 1e-2), (torch.float32, 1e-5)])
def test_dense_forward_bias_gelu_preact_residual(dtype, tol):
    from tethys_speech_b200 import _lib

    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(1)
    m, n, k = 1000, 640, 320
    x = torch.randn(m, k, generator=g).to(dtype).to(dev)
    w = (torch.randn(k, n, generator=g) * 0.1).to(dtype).to(dev)        # Keras Dense kernel [in, out] -> b_major = 1
    bias = torch.randn(n, generator=g).to(dev)
    res = torch.randn(m, n, generator=g).to(dtype).to(dev)
    y = torch.empty(m, n, dtype=dtype, device=dev); pre = torch.empty_like(y)
    dt = _lib.TS_BF16 if dtype == torch.bfloat16 else _lib.TS_F32
    _gemm(x, w, y, m, n, k, 0, 1, k, n, dt, dt, bias=bias, residual=res, c_preact=pre, act=1)
    u = x.float() @ w.float() + bias
    assert _rel(pre, u) < tol
    assert _rel(y, torch.nn.functional.gelu(u) + res.float()) < tol


def test_gelu_backward_epilogue():
    from tethys_speech_b200 import _lib

    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(2)
    m, n, k = 700, 512, 256
    dy = torch.randn(m, k, generator=g).bfloat16().to(dev)
    w = (torch.randn(n, k, generator=g) * 0.1).bfloat16().to(dev)       # dX = dY W^T: W stored [n][k] -> b_major = 0
    u = (torch.randn(m, n, generator=g) * 2).bfloat16().to(dev)
    dx = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
    _gemm(dy, w, dx, m, n, k, 0, 0, k, k, _lib.TS_BF16, _lib.TS_BF16, act=2, act_aux=u)
    uf = u.float().requires_grad_(True)
    torch.nn.functional.gelu(uf).backward(dy.float() @ w.float().t())
    assert _rel(dx, uf.grad) < 1e-2


def test_weight_gradient_split_k_accumulates():
    from tethys_speech_b200 import _lib

    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(3)
    rows, kin, nout = 9000, 256, 384                                       # few output tiles, long reduce dim -> split-K
    x = torch.randn(rows, kin, generator=g).bfloat16().to(dev)
    dy = torch.randn(rows, nout, generator=g).bfloat16().to(dev)
    dw = torch.full((kin, nout), 0.5, device=dev)                          # C += A^T B on top of existing content
    _gemm(x, dy, dw, kin, nout, rows, 1, 1, kin, nout, _lib.TS_BF16, _lib.TS_F32, accumulate=1)
    assert _rel(dw, x.float().t() @ dy.float() + 0.5) < 1e-5


def test_strided_conv_as_window_gemm_without_im2col():
    """Conv1D k=3, s=2 over [T, C] stored with its SAME-padding row: output row t reads input rows 2t..2t+2 = a 3C-long
    window starting at row stride 2C (lda < k): the A operand's rows overlap."""
    from tethys_speech_b200 import _lib

    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(4)
    T, Cin, Cout = 400, 64, 128
    x = torch.randn(T + 1, Cin, generator=g).bfloat16().to(dev)
    x[T] = 0                                                                # right SAME-padding row (k3 s2: pad 0 / 1)
    w = (torch.randn(3, Cin, Cout, generator=g) * 0.1).bfloat16().to(dev)
    To = T // 2
    y = torch.empty(To, Cout, dtype=torch.bfloat16, device=dev)
    _gemm(x, w, y, To, Cout, 3 * Cin, 0, 1, 2 * Cin, Cout, _lib.TS_BF16, _lib.TS_BF16)
    ref = torch.nn.functional.conv1d(x.float().t().unsqueeze(0), w.float().permute(2, 1, 0), stride=2).squeeze(0).t()
    assert _rel(y, ref[:To]) < 1e-2


def test_misaligned_operands_fall_to_cuda_core_engine_not_cpu():
    from tethys_speech_b200 import _lib

    dev = torch.device("cuda", 0)
    a = torch.randn(33, 30, device=dev).bfloat16()                         # row stride 60 B: not TMA-describable
    b = torch.randn(30, 20, device=dev).bfloat16()
    c = torch.empty(33, 20, dtype=torch.bfloat16, device=dev)
    _gemm(a, b, c, 33, 20, 30, 0, 1, 30, 20, _lib.TS_BF16, _lib.TS_BF16)
    assert _rel(c, a.float() @ b.float()) < 1e-2


@pytest.mark.parametrize("force", [2, 3, 4])   # 1-CTA tiles / CTA-pair tiles / 4-CTA cluster (two pairs, B multicast)
def test_groupnorm_statistics_in_the_epilogue(force):
    """ts_gemm_desc.gn_accum: the conv GEMM takes the GroupNormalization moments (V:167-176, over time x channels of a group, per
    batch element) of its own output, skipping the window-slack rows behind every batch block; warps that straddle a batch
    boundary and the clipped last tile are part of the shape."""
    from tethys_speech_b200 import _lib
    from tethys_speech_b200.runtime import stream_ptr

    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(11)
    B, rpb, valid, k, n, G = 3, 237, 229, 192, 256, 4                       # 64 channels per group; 711 rows = 5.55 tiles of 128
    m = B * rpb
    x = torch.randn(m, k, generator=g).bfloat16().to(dev)
    w = (torch.randn(k, n, generator=g) * 0.1).bfloat16().to(dev)
    y = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
    accum = torch.zeros(B, G, 2, dtype=torch.float64, device=dev)
    ctx = _lib.context(0)
    d = _lib.GemmDesc()
    d.a, d.b, d.c = x.data_ptr(), w.data_ptr(), y.data_ptr()
    d.m, d.n, d.k, d.a_major, d.b_major = m, n, k, 0, 1
    d.lda, d.ldb, d.ldc = k, n, n
    d.batch1 = d.batch2 = 1
    d.in_dtype = d.out_dtype = _lib.TS_BF16
    d.alpha, d.force_engine = 1.0, force
    d.gn_accum, d.gn_rows_per_batch, d.gn_valid_rows, d.gn_groups = accum.data_ptr(), rpb, valid, G
    ctx.check(ctx.lib.ts_gemm(ctx.h, C.byref(d), stream_ptr()))
    torch.cuda.synchronize()
    ctx.watchdog()
    u = (x.float() @ w.float()).double().view(B, rpb, G, n // G)[:, :valid]
    ref = torch.stack([u.sum(dim=(1, 3)), (u * u).sum(dim=(1, 3))], dim=-1)
    assert _rel(y, x.float() @ w.float()) < 1e-2
    assert float(((accum - ref).abs() / ref.abs().clamp_min(1.0)).max()) < 1e-5, (accum, ref)
    # an epilogue that cannot take them says so
    d.act = 1
    with pytest.raises(_lib.TethysError):
        ctx.check(ctx.lib.ts_gemm(ctx.h, C.byref(d), stream_ptr()))


@pytest.mark.parametrize("amaj,bmaj", [(0, 0), (0, 1), (1, 1)])
def test_four_cta_multicast_engine(amaj, bmaj):
    """force_engine = 4: a cluster of two CTA pairs on vertically adjacent 256-row tiles; each CTA fetches a quarter of the B operand
    and TMA-multicasts it to its counterpart in the other pair. Ragged m (clipped last cluster tile), several tiles per cluster,
    K-major and MN-major operands, bias + residual epilogue, against fp32 torch on the same bf16 inputs."""
    from tethys_speech_b200 import _lib
    from tethys_speech_b200.runtime import stream_ptr

    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(21 + amaj * 2 + bmaj)
    m, n, k = 2896, 768, 448      # 5.66 cluster tiles of 512 rows; m % 8 == 0 so that an MN-major A has 16-byte row strides
    a = torch.randn((m, k) if amaj == 0 else (k, m), generator=g).bfloat16().to(dev)
    b = (torch.randn((n, k) if bmaj == 0 else (k, n), generator=g) * 0.1).bfloat16().to(dev)
    bias = torch.randn(n, generator=g).to(dev)
    res = torch.randn(m, n, generator=g).bfloat16().to(dev)
    y = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
    ctx = _lib.context(0)
    d = _lib.GemmDesc()
    d.a, d.b, d.c = a.data_ptr(), b.data_ptr(), y.data_ptr()
    d.m, d.n, d.k, d.a_major, d.b_major = m, n, k, amaj, bmaj
    d.lda, d.ldb, d.ldc = a.shape[1], b.shape[1], n
    d.batch1 = d.batch2 = 1
    d.in_dtype = d.out_dtype = _lib.TS_BF16
    d.alpha, d.force_engine = 1.0, 4
    d.bias, d.residual, d.ldr = bias.data_ptr(), res.data_ptr(), n
    ctx.check(ctx.lib.ts_gemm(ctx.h, C.byref(d), stream_ptr()))
    torch.cuda.synchronize()
    ctx.watchdog()
    af = a.float() if amaj == 0 else a.float().t()
    bf = b.float().t() if bmaj == 0 else b.float()
    assert _rel(y, af @ bf + bias + res.float()) < 1e-2
