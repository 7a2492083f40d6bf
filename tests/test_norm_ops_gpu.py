"""K11 / K5 single-operator entry points (ts_layernorm_fwd/bwd, ts_groupnorm_gelu_fwd) against plain PyTorch fp32."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ctx():
    from tethys_speech_b200 import _lib
    from tethys_speech_b200.runtime import stream_ptr
    return _lib, _lib.context(0), stream_ptr


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


@pytest.mark.parametrize("rows,cols,dtype,tol", [(300, 768, torch.float32, 1e-5), (257, 256, torch.float32, 1e-5),
                                                 (6000, 768, torch.bfloat16, 2e-2)])
def test_layernorm_fwd_bwd(rows, cols, dtype, tol):
    _lib, ctx, sp = _ctx()
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(rows + cols)
    x = (torch.randn(rows, cols, generator=g) * 2 + 0.5).to(dtype).to(dev)
    dy = torch.randn(rows, cols, generator=g).to(dtype).to(dev)
    dres = torch.randn(rows, cols, generator=g).to(dtype).to(dev)
    gamma = (1 + 0.1 * torch.randn(cols, generator=g)).to(dev)
    beta = (0.1 * torch.randn(cols, generator=g)).to(dev)
    y = torch.empty_like(x); dx = torch.empty_like(x)
    mean = torch.empty(rows, device=dev); rstd = torch.empty(rows, device=dev)
    dgamma = torch.zeros(cols, device=dev); dbeta = torch.zeros(cols, device=dev)
    dt = _lib.TS_F32 if dtype == torch.float32 else _lib.TS_BF16
    ctx.check(ctx.lib.ts_layernorm_fwd(ctx.h, dt, _p(x), _p(gamma), _p(beta), _p(y), _p(mean), _p(rstd), rows, cols, 1e-5, sp()))
    ctx.check(ctx.lib.ts_layernorm_bwd(ctx.h, dt, _p(dy), _p(x), _p(gamma), _p(mean), _p(rstd), _p(dres), _p(dx), _p(dgamma),
                                       _p(dbeta), rows, cols, sp()))
    torch.cuda.synchronize()
    xr = x.float().detach().requires_grad_(True)
    gr = gamma.clone().requires_grad_(True); br = beta.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (cols,), gr, br, 1e-5)
    yr.backward(dy.float())

    def rel(a, b):
        return float((a.float() - b).norm() / b.norm())

    assert rel(y, yr.detach()) < tol
    assert rel(dx, xr.grad + dres.float()) < tol
    assert rel(dgamma, gr.grad) < tol and rel(dbeta, br.grad) < tol


@pytest.mark.parametrize("B,T,Cc,dtype,tol", [(2, 333, 512, torch.float32, 1e-5), (3, 100, 128, torch.float32, 1e-5),
                                               (2, 4000, 512, torch.bfloat16, 2e-2)])
def test_groupnorm_gelu_fwd(B, T, Cc, dtype, tol):
    _lib, ctx, sp = _ctx()
    dev = torch.device("cuda", 0)
    G = 16
    g = torch.Generator().manual_seed(T)
    x = (torch.randn(B, T, Cc, generator=g) * 1.5 + 0.3).to(dtype).to(dev)
    gamma = (1 + 0.1 * torch.randn(Cc, generator=g)).to(dev)
    beta = (0.1 * torch.randn(Cc, generator=g)).to(dev)
    y = torch.empty_like(x)
    mean = torch.empty(B, G, device=dev); rstd = torch.empty(B, G, device=dev)
    accum = torch.zeros(2 * B * G, dtype=torch.float64, device=dev)
    dt = _lib.TS_F32 if dtype == torch.float32 else _lib.TS_BF16
    ctx.check(ctx.lib.ts_groupnorm_gelu_fwd(ctx.h, dt, _p(x), _p(gamma), _p(beta), _p(y), _p(mean), _p(rstd), _p(accum), B, T, Cc, G,
                                            1e-5, sp()))
    torch.cuda.synchronize()
    # V:140-196: reshape [B,T,G,C/G], moments over (T, C/G), normalise, per-channel affine, then exact GELU (V:132-136)
    xf = x.float().reshape(B, T, G, Cc // G)
    mu = xf.mean(dim=(1, 3), keepdim=True)
    var = xf.var(dim=(1, 3), unbiased=False, keepdim=True)
    ref = ((xf - mu) / torch.sqrt(var + 1e-5)).reshape(B, T, Cc) * gamma + beta
    ref = torch.nn.functional.gelu(ref)
    assert float((y.float() - ref).norm() / ref.norm()) < tol
    assert torch.allclose(mean, mu.reshape(B, G), atol=1e-4, rtol=1e-4)
