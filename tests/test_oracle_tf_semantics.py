"""Hand-computed micro-cases for every TF/Keras 2.10 semantic the oracle restates, because the
reference ships no tests to pin them (parity unpinned). CPU only."""
import math

import numpy as np
import torch

from oracle import tf_ops as T
from oracle import wav2vec2_oracle as WO
from oracle import whisper_oracle as HO


def test_a1_same_padding_extra_goes_right():
    # length 5, k=3, s=2 -> T_out=3, pad_total=(3-1)*2+3-5=2 -> left 1 / right 1
    assert T.same_pad(5, 3, 2) == (3, 1, 1)
    # length 6, k=3, s=2 -> T_out=3, pad_total=1 -> left 0 / right 1 (torch's symmetric padding would use 1/1)
    assert T.same_pad(6, 3, 2) == (3, 0, 1)
    # Whisper conv2 on 3000 frames, conv0 of wav2vec2 on 32000 samples (k=10, s=5): left 2 / right 3
    assert T.same_pad(3000, 3, 2) == (1500, 0, 1)
    assert T.same_pad(32000, 10, 5) == (6400, 2, 3)
    assert T.same_pad(750, 128, 1) == (750, 63, 64)


def test_a1_conv1d_same_values():
    x = torch.tensor([[1., 2., 3., 4., 5., 6.]]).unsqueeze(-1)          # [1,6,1]
    k = torch.tensor([1., 10., 100.]).view(3, 1, 1)
    y = T.conv1d_same(x, k, stride=2)[0, :, 0]
    # windows start at 0,2,4 (left pad 0); last window sees one zero on the right; cross-correlation (no flip)
    assert y.tolist() == [1 + 20 + 300, 3 + 40 + 500, 5 + 60 + 0]
    y1 = T.conv1d_same(x, k, stride=1)[0, :, 0]
    assert y1.tolist() == [0 + 10 + 200, 1 + 20 + 300, 2 + 30 + 400, 3 + 40 + 500, 4 + 50 + 600, 5 + 60 + 0]


def test_a2_grouped_conv_channel_blocks():
    # groups=2, 2 in-ch and 1 out-ch per group... Keras layout [k, Cin/groups, Cout]
    x = torch.arange(8, dtype=torch.float64).view(1, 2, 4)              # T=2, C=4
    kern = torch.zeros(1, 2, 2, dtype=torch.float64)
    kern[0, :, 0] = torch.tensor([1., 1.])       # out ch 0 (group 0) sums in-ch 0,1
    kern[0, :, 1] = torch.tensor([1., -1.])      # out ch 1 (group 1) = in-ch 2 - in-ch 3
    y = T.conv1d_same(x, kern, groups=2)
    assert y[0, :, 0].tolist() == [0 + 1, 4 + 5]
    assert y[0, :, 1].tolist() == [2 - 3, 6 - 7]


def test_a4_layernorm_biased_variance():
    x = torch.tensor([[1., 2., 3., 4.]], dtype=torch.float64)
    y = T.layer_norm(x, torch.ones(4, dtype=torch.float64), torch.zeros(4, dtype=torch.float64), eps=0.0)
    var = 1.25  # biased
    assert torch.allclose(y, (x - 2.5) / math.sqrt(var))


def test_a5_gelu_is_exact_erf():
    x = torch.tensor([-1.0, 0.0, 0.5, 2.0], dtype=torch.float64)
    ref = torch.tensor([0.5 * v * (1 + math.erf(v / math.sqrt(2))) for v in x.tolist()], dtype=torch.float64)
    assert torch.allclose(T.gelu(x), ref, atol=1e-15)
    tanh_approx = 0.5 * x * (1 + torch.tanh(math.sqrt(2 / math.pi) * (x + 0.044715 * x ** 3)))
    assert not torch.allclose(T.gelu(x), tanh_approx, atol=1e-5)


def test_a15_group_norm_stats_over_time_and_group_channels():
    # C=4, 2 groups: group g = channels [2g, 2g+2); statistics over (T, 2) per batch item
    x = torch.tensor([[[1., 3., 10., 10.], [5., 7., 10., 14.]]], dtype=torch.float64)     # [1,2,4]
    y = T.group_norm(x, torch.ones(4, dtype=torch.float64), torch.zeros(4, dtype=torch.float64), groups=2, eps=0.0)
    g0 = torch.tensor([1., 3., 5., 7.], dtype=torch.float64)
    mu, var = g0.mean(), g0.var(unbiased=False)
    assert torch.allclose(y[0, :, 0], (torch.tensor([1., 5.], dtype=torch.float64) - mu) / var.sqrt())
    assert torch.allclose(y[0, :, 1], (torch.tensor([3., 7.], dtype=torch.float64) - mu) / var.sqrt())
    g1 = torch.tensor([10., 10., 10., 14.], dtype=torch.float64)
    assert torch.allclose(y[0, 1, 3], (14. - g1.mean()) / g1.var(unbiased=False).sqrt())


def test_a8_anticausal_mask_and_fp32_absorption():
    S = 4
    mask = 1.0 - torch.tril(torch.ones(S, S))            # 1 - band_part(ones,-1,0): 1 strictly above the diagonal
    assert mask.tolist() == [[0, 1, 1, 1], [0, 0, 1, 1], [0, 0, 0, 1], [0, 0, 0, 0]]
    add = (1.0 - mask) * -1e9                            # W:153: -1e9 on j <= i
    assert add[2].tolist() == [-1e9, -1e9, -1e9, 0.0]
    # fp32: |score| < 32 is absorbed exactly (ulp(1e9) = 64) -> last row is uniform
    s = torch.tensor([3.25, -7.5, 0.125, 31.0], dtype=torch.float32)
    assert torch.equal(s + torch.tensor(-1e9, dtype=torch.float32), torch.full((4,), -1e9, dtype=torch.float32))
    assert not torch.equal(torch.tensor(33.0, dtype=torch.float32) + torch.tensor(-1e9, dtype=torch.float32), torch.tensor(-1e9, dtype=torch.float32))


def test_c2_double_label_shift():
    cfg = HO.WhisperConfig("tiny")
    cfg.d_model, cfg.d_ff, cfg.encoder_layers, cfg.decoder_layers = 16, 32, 1, 1
    cfg.encoder_attention_heads = cfg.decoder_attention_heads = 2
    cfg.vocab_size, cfg.n_mels, cfg.n_ctx, cfg.decoder_start_token_id = 11, 8, 16, 10
    w = HO.init_weights(cfg, dtype=torch.float64)
    feats = torch.randn(1, 8, 20, dtype=torch.float64)
    labels = torch.tensor([[1, 5, 6, 2, 0, 0]], dtype=torch.int32)
    out = HO.forward(cfg, w, feats, labels)
    # loss = mean over S-1 positions of CE(logits[:, s], labels[:, s+1]) — pads (0) included
    lp = torch.log_softmax(out["logits"][0, :-1], -1)
    manual = -torch.stack([lp[s, int(labels[0, s + 1])] for s in range(5)]).mean()
    assert torch.allclose(out["loss"], manual)


def test_a10_argmin_first_minimum_and_topk_tie_break():
    d = torch.tensor([[3., 1., 1., 2.]])
    assert int(torch.argmin(d, -1)) == 1
    # negative sampler: positions of the K smallest random ints; equal values -> lower index first
    r = torch.tensor([[5, 0, 3, 0, 9, 3]])
    neg = WO.negative_indices_from_random(r, num_negatives=4)
    assert neg.tolist() == [[1, 3, 2, 5]]
    # T-1 < K: take T-1 and tile up to K (V:911-931)
    neg2 = WO.negative_indices_from_random(torch.tensor([[2, 1, 0]]), num_negatives=5)
    assert neg2.tolist() == [[2, 1, 2, 1, 2]]


def test_legacy_sampler_formula():
    perm = torch.tensor([2, 0, 3, 1])
    neg = WO.legacy_negative_indices(4, perm, 3)          # neg[t,k] = perm[(k-(t+1)) mod T]
    assert neg[0].tolist() == [perm[3].item(), perm[0].item(), perm[1].item()]
    assert neg[2].tolist() == [perm[1].item(), perm[2].item(), perm[3].item()]


def test_a11_clipping_rules():
    g = [torch.tensor([3.0, 4.0]), torch.tensor([12.0])]             # global norm 13
    c, n = T.clip_by_global_norm(g, 1.0)
    assert abs(float(n) - 13.0) < 1e-6
    assert torch.allclose(c[0], torch.tensor([3.0, 4.0]) / 13.0) and torch.allclose(c[1], torch.tensor([12.0 / 13.0]))
    small = [torch.tensor([0.3, 0.4])]
    assert torch.equal(T.clip_by_global_norm(small, 1.0)[0][0], small[0])     # norm 0.5 < 1: untouched
    per = T.clip_by_norm_each(g, 1.0)
    assert torch.allclose(per[0], torch.tensor([0.6, 0.8])) and torch.allclose(per[1], torch.tensor([1.0]))


def test_a12_keras_legacy_adam_eps_outside_bias_correction():
    p, g = torch.tensor([1.0], dtype=torch.float64), torch.tensor([0.5], dtype=torch.float64)
    m, v = torch.zeros(1, dtype=torch.float64), torch.zeros(1, dtype=torch.float64)
    T.keras_adam_step([p], [g], [m], [v], t=1, lr=0.1, beta1=0.9, beta2=0.999, eps=1e-7)
    m1, v1 = 0.05, 0.00025
    lr_t = 0.1 * math.sqrt(1 - 0.999) / (1 - 0.9)
    expected = 1.0 - lr_t * m1 / (math.sqrt(v1) + 1e-7)
    assert abs(float(p) - expected) < 1e-15
    torch_style = 1.0 - 0.1 * (m1 / 0.1) / (math.sqrt(v1 / 0.001) + 1e-7)     # eps inside: differs at the 1e-7 level
    assert abs(expected - torch_style) > 1e-9


def test_c5_positional_table_interleaved_fp32():
    pe = T.sinusoid_pe(10, 8, torch.float64)
    assert float(pe[0, 0]) == 0.0 and float(pe[0, 1]) == 1.0
    assert abs(float(pe[3, 2]) - float(np.float32(math.sin(3 * math.exp(2 * -(math.log(10000.0) / 8)))))) < 1e-12
    assert abs(float(pe[3, 3]) - float(np.float32(math.cos(3 * math.exp(2 * -(math.log(10000.0) / 8)))))) < 1e-12


def test_vq_no_gradient_into_projection_and_diversity_has_no_gradient():
    cfg = WO.Wav2Vec2Config("tiny")
    w = WO.randomize_weights(WO.init_weights(cfg, dtype=torch.float64))
    wave = torch.randn(1, 1600, dtype=torch.float64)
    T_ = WO.num_frames(cfg, 1600)
    neg = WO.negative_indices_from_random(torch.randint(0, T_, (1, T_)), cfg.num_negatives)
    out, g = WO.loss_and_grads(cfg, w, wave, neg)
    assert float(g["quantizer.projection.kernel"].abs().max()) == 0.0        # V:1237-1240 None -> zeros
    assert float(g["quantizer.projection.bias"].abs().max()) == 0.0
    assert float(g["quantizer.codevectors"].abs().max()) > 0.0
    assert out["code_indices"].dtype == torch.int64


def test_param_counts_match_survey():
    assert sum(v.numel() for v in WO.init_weights(WO.Wav2Vec2Config("tiny")).values()) == 3_499_840
    assert sum(v.numel() for v in WO.init_weights(WO.Wav2Vec2Config("small")).values()) == 20_466_816
    assert sum(v.numel() for v in WO.init_weights(WO.Wav2Vec2Config("base")).values()) == 92_297_728
    assert sum(v.numel() for v in HO.init_weights(HO.WhisperConfig("tiny")).values()) == 56_933_376


def test_ctc_loss_restatement_against_brute_force_enumeration():
    """WS:897-929: tf.nn.ctc_loss(blank_index=0) = -log of the total probability of every frame-level path that collapses (merge
    repeats, drop blanks) to the label sequence. The oracle's restatement is checked against literally that sum, over all V^T
    paths, for sequences with a repeated label (needs a blank between), a padded label row, and an impossible one (inf)."""
    import itertools

    g = torch.Generator().manual_seed(5)
    Tn, V = 5, 4
    logits = torch.randn(3, Tn, V, generator=g, dtype=torch.float64)
    labels = torch.tensor([[1, 1, 0], [2, 3, 1], [3, 0, 0]])
    probs = torch.softmax(logits, dim=-1)

    def collapse(path):
        out, prev = [], None
        for k in path:
            if k != prev and k != 0:
                out.append(k)
            prev = k
        return out

    want = []
    for b in range(3):
        tgt = [int(x) for x in labels[b] if x > 0]
        total = 0.0
        for path in itertools.product(range(V), repeat=Tn):
            if collapse(path) == tgt:
                pr = 1.0
                for t, k in enumerate(path):
                    pr *= float(probs[b, t, k])
                total += pr
        want.append(-np.log(total))
    loss, per = WO.ctc_loss(logits, labels, reduction="sum")
    assert np.allclose(per.numpy(), want, rtol=1e-12)
    assert abs(float(loss) - sum(want)) < 1e-10
    assert abs(float(WO.ctc_loss(logits, labels, reduction="mean")[0]) - np.mean(want)) < 1e-10
    # no alignment: 3 distinct labels + a repeat need >= 5 frames; with 3 frames the loss is +inf, 0 under zero_infinity (WS:920-921)
    lab = torch.tensor([[1, 1, 2]])
    _, per = WO.ctc_loss(logits[:1, :3], lab)
    assert torch.isinf(per).all()
    assert float(WO.ctc_loss(logits[:1, :3], lab, zero_infinity=True)[0]) == 0.0
