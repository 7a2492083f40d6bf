"""Span masking utilities (V:1073-1120): the integer dilation is bit-exact against the oracle restatement."""
import numpy as np
import pytest

from oracle import masking_oracle as M


def test_expand_spans_hand_case():
    start = np.zeros((1, 12), dtype=bool)
    start[0, [2, 9]] = True
    got = M.expand_spans(start, 3)
    want = np.zeros((1, 12), dtype=bool)
    want[0, [2, 3, 4, 9, 10, 11]] = True            # spans are clipped at the end, never wrap
    assert np.array_equal(got, want)
    assert np.array_equal(M.expand_spans(start, 1), start)
    assert M.expand_spans(start, 50).sum() == 10      # everything from the first start on


def test_apply_masks_shapes_and_values():
    x = np.arange(2 * 5 * 8, dtype=np.float32).reshape(2, 5, 8) + 1
    st = np.zeros((2, 5), dtype=bool); st[1, 3] = True
    y, m = M.apply_time_mask(x, st, 2)
    assert m.shape == (2, 5, 1) and np.array_equal(y[1, 3:5], np.zeros((2, 8))) and np.array_equal(y[0], x[0])
    sf = np.zeros((2, 8), dtype=bool); sf[0, 6] = True
    y, m = M.apply_feature_mask(x, sf, 10)
    assert m.shape == (2, 1, 8) and np.all(y[0, :, 6:] == 0) and np.array_equal(y[0, :, :6], x[0, :, :6])


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,H,L,dtype", [(3, 250, 512, 10, "float32"), (2, 77, 768, 10, "bfloat16"), (1, 5, 8, 7, "float32")])
def test_span_masks_on_gpu_are_bit_exact(B, T, H, L, dtype):
    import torch
    from tethys_speech_b200 import wav2vec2 as W

    rng = np.random.default_rng(B * T)
    x = torch.from_numpy(rng.standard_normal((B, T, H)).astype(np.float32)).to(getattr(torch, dtype)).cuda()
    st = rng.random((B, T)) < 0.05
    sf = rng.random((B, H)) < 0.05
    y, m = W.apply_time_mask(x, mask_length=L, start_mask=torch.from_numpy(st))
    ry, rm = M.apply_time_mask(x.float().cpu().numpy(), st, L)
    assert np.array_equal(m.cpu().numpy(), rm) and np.array_equal(y.float().cpu().numpy(), ry)
    y, m = W.apply_feature_mask(x, mask_length=L, start_mask=torch.from_numpy(sf))
    ry, rm = M.apply_feature_mask(x.float().cpu().numpy(), sf, L)
    assert np.array_equal(m.cpu().numpy(), rm) and np.array_equal(y.float().cpu().numpy(), ry)
    # drawn on the device when no starts are given: mask rate is about 1 - (1 - p)^L
    _, m = W.apply_time_mask(torch.zeros(8, 2000, 8, device="cuda"), mask_prob=0.05, mask_length=10)
    assert abs(float(m.mean()) - (1 - 0.95 ** 10)) < 0.03
