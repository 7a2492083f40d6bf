"""K10 — fused tcgen05 attention (ts_attn_fwd / ts_attn_bwd) against a plain PyTorch fp32 reference of the same op
on the same bf16-rounded inputs (W:147-167, V:348-362). Tolerance: 2e-2 relative (the bf16 bar of BASELINE.json)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 2e-2


def _run(B, nh, Tq, Tk, mask, cross, seed=0):
    from tethys_speech_b200 import _lib
    from tethys_speech_b200.runtime import stream_ptr

    ctx = _lib.context(0)
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(seed)
    H = nh * 64
    if cross:
        qb = torch.randn(B, Tq, H, generator=g).bfloat16().to(dev)
        kvb = torch.randn(B, Tk, 2 * H, generator=g).bfloat16().to(dev)
        q, k, v = qb, kvb[..., :H], kvb[..., H:]
        gq = torch.zeros_like(qb); gkv = torch.zeros_like(kvb)
        dq, dk, dv = gq, gkv[..., :H], gkv[..., H:]
    else:
        qkv = torch.randn(B, Tq, 3 * H, generator=g).bfloat16().to(dev)
        q, k, v = qkv[..., :H], qkv[..., H:2 * H], qkv[..., 2 * H:]
        gqkv = torch.zeros_like(qkv)
        dq, dk, dv = gqkv[..., :H], gqkv[..., H:2 * H], gqkv[..., 2 * H:]
    do = torch.randn(B, Tq, H, generator=g).bfloat16().to(dev)
    o = torch.zeros(B, Tq, H, dtype=torch.bfloat16, device=dev)
    stats = torch.zeros(B, nh, Tq, 2, device=dev)
    dsum = torch.zeros(B, nh, Tq, device=dev)
    scale = 0.125
    d = _lib.AttnDesc()
    d.q, d.k, d.v, d.o = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr()
    d.q_ld, d.q_bs = q.stride(1), q.stride(0)
    d.kv_ld, d.kv_bs = k.stride(1), k.stride(0)
    d.o_ld, d.o_bs = H, Tq * H
    d.stats = stats.data_ptr()
    d.batch, d.heads, d.tq, d.tk, d.head_dim = B, nh, Tq, Tk, 64
    d.scale, d.mask_mode, d.drop, d.seed = scale, mask, 0.0, 1
    d.d_o, d.dq, d.dk, d.dv = do.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    d.dq_ld, d.dq_bs, d.dkv_ld, d.dkv_bs = dq.stride(1), dq.stride(0), dk.stride(1), dk.stride(0)
    d.dsum = dsum.data_ptr()
    ctx.check(ctx.lib.ts_attn_fwd(ctx.h, C.byref(d), stream_ptr()))
    ctx.check(ctx.lib.ts_attn_bwd(ctx.h, C.byref(d), stream_ptr()))
    torch.cuda.synchronize()
    ctx.watchdog()
    # fp32 reference
    qf = q.float().reshape(B, Tq, nh, 64).transpose(1, 2).detach().requires_grad_(True)
    kf = k.float().reshape(B, Tk, nh, 64).transpose(1, 2).detach().requires_grad_(True)
    vf = v.float().reshape(B, Tk, nh, 64).transpose(1, 2).detach().requires_grad_(True)
    s = (qf @ kf.transpose(-1, -2)) * scale
    if mask == 1:
        i = torch.arange(Tq, device=dev)[:, None]
        j = torch.arange(Tk, device=dev)[None, :]
        s = s + torch.where(j <= i, torch.tensor(-1e9, device=dev), torch.tensor(0.0, device=dev))
    ref = torch.softmax(s, -1) @ vf
    ref_o = ref.transpose(1, 2).reshape(B, Tq, H)
    ref_o.backward(do.float())

    def rel(a, b):
        return float((a.float() - b).abs().max() / b.abs().max())

    errs = {"o": rel(o, ref_o),
            "dq": rel(dq, qf.grad.transpose(1, 2).reshape(B, Tq, H)),
            "dk": rel(dk, kf.grad.transpose(1, 2).reshape(B, Tk, H)),
            "dv": rel(dv, vf.grad.transpose(1, 2).reshape(B, Tk, H))}
    assert all(e < TOL for e in errs.values()), errs
    return o, stats


def test_self_attention_ragged_tiles():
    _run(2, 3, 200, 200, 0, False)


def test_self_attention_many_tiles():
    _run(1, 2, 750, 750, 0, False)


def test_cross_attention_short_queries():
    _run(2, 4, 100, 300, 0, True)


def test_decoder_anticausal_mask_fully_masked_row_is_uniform():
    """W:416-418 + W:150-154: query i sees only keys j > i; the last query has every key at -1e9 and, through fp32
    absorption, attends uniformly. Checked against the literal fp32 formula."""
    _run(2, 2, 100, 100, 1, False)


def test_unsupported_head_dim_is_an_error_not_a_fallback():
    from tethys_speech_b200 import _lib
    from tethys_speech_b200.runtime import stream_ptr

    ctx = _lib.context(0)
    d = _lib.AttnDesc()
    d.head_dim = 32
    rc = ctx.lib.ts_attn_fwd(ctx.h, C.byref(d), stream_ptr())
    assert rc == -6  # TS_EUNSUPPORTED
