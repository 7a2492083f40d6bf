"""bf16 gradient buckets for the cross-replica SUM (ts_grad_pack_bf16 / ts_grad_unpack_bf16; SURVEY §8e "bf16 buckets in perf
mode"): pack = round-to-nearest bf16 of grad * scale, unpack = exact widening, on arbitrary 64-element aligned sub-ranges of
the arena (the all-reduce buckets) — bit-exact against torch's own conversion."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_pack_unpack_ranges_bit_exact():
    from tethys_speech_b200 import wav2vec2 as W

    model = W.Wav2Vec2ForPreTraining(W.Wav2Vec2Config("tiny"), precision="bf16", device=0)
    prog = model._prog
    assert prog.ar_bf16()
    g = torch.Generator(device="cuda").manual_seed(0)
    prog.grads.copy_(torch.randn(prog.n, generator=g, device="cuda") * torch.logspace(-12, 2, prog.n, device="cuda"))
    ref = prog.grads.clone()
    scale = torch.tensor([0.37], device="cuda")
    ends = [0] + prog.stage_ends
    g16 = prog.grads_lp()
    g16.fill_(7.0)
    for a0, a1 in zip(ends[:-1], ends[1:]):                      # bucket by bucket, as the overlapped all-reduce does
        view = prog.pack_grads(a0, a1, scale=scale)
        assert view.data_ptr() == g16[a0:a1].data_ptr() and view.numel() == a1 - a0
    torch.cuda.synchronize()
    assert torch.equal(g16, (ref * scale).bfloat16())
    prog.grads.zero_()
    prog.unpack_grads(ends[1], ends[3])                          # a sub-range only
    torch.cuda.synchronize()
    want = torch.zeros_like(ref)
    want[ends[1]:ends[3]] = g16[ends[1]:ends[3]].float()
    assert torch.equal(prog.grads, want)
    prog.unpack_grads()
    prog.pack_grads()                                            # no scale, whole arena; bf16 -> fp32 -> bf16 is the identity
    torch.cuda.synchronize()
    assert torch.equal(prog.grads, (ref * scale).bfloat16().float())
    assert torch.equal(g16, (ref * scale).bfloat16())
    prog.ctx.watchdog()


def test_fp32_mode_never_uses_bf16_buckets(monkeypatch):
    from tethys_speech_b200 import wav2vec2 as W

    model = W.Wav2Vec2ForPreTraining(W.Wav2Vec2Config("tiny"), precision="fp32", device=0)
    assert not model._prog.ar_bf16()
    m16 = W.Wav2Vec2ForPreTraining(W.Wav2Vec2Config("tiny"), precision="bf16", device=0)
    monkeypatch.setenv("TETHYS_AR_DTYPE", "fp32")
    assert not m16._prog.ar_bf16()


@pytest.mark.parametrize("clipnorm", [None, 1.0])
def test_adam_reads_the_bf16_bucket(clipnorm):
    """ts_optim_step_lp (clipnorm + Adam straight from the all-reduced bf16 bucket) against ts_grad_unpack_bf16 + ts_optim_step on
    parameters, Adam state and the refreshed bf16 compute copy; the fp32 gradient arena is not read. Without clipnorm the two are
    bit-identical (every gradient is widened exactly); with it the per-variable norms are float atomics over several blocks, whose
    summation order differs from run to run in the last bits for either entry point."""
    from tethys_speech_b200 import wav2vec2 as W
    from tethys_speech_b200.runtime import Adam

    res = []
    for use_lp in (False, True):
        model = W.Wav2Vec2ForPreTraining(W.Wav2Vec2Config("tiny"), precision="bf16", device=0, seed=0)
        prog = model._prog
        init = prog.params.clone()
        opt = Adam(learning_rate=3e-5, epsilon=1e-8, clipnorm=clipnorm)
        g = torch.Generator(device="cuda").manual_seed(1)
        for step in range(2):
            prog.grads.copy_(torch.randn(prog.n, generator=g, device="cuda") * torch.logspace(-6, 1, prog.n, device="cuda"))
            prog.pack_grads()
            if use_lp:
                prog.grads.fill_(123.0)
                opt.update(model, grads_lp=prog.grads_lp())
            else:
                prog.unpack_grads()
                opt.update(model)
        torch.cuda.synchronize()
        st = opt._bind(model)
        res.append([prog.params - init, prog.params_lp.float(), st["m"].clone(), st["v"].clone()])
        prog.ctx.watchdog()
    assert float(res[0][0].abs().max()) > 0
    for a, b in zip(*res):
        if clipnorm is None:
            assert torch.equal(a, b)
        else:
            assert float((a - b).norm() / a.norm()) < 1e-6
