"""K1 log-mel front end (W:739-766). CPU: the oracle restatement against closed-form known answers (TensorFlow is not
importable; SURVEY App. A-9 lists the semantics being pinned). GPU: the fused kernel against the oracle."""
import numpy as np
import pytest

from oracle import logmel_oracle as L


def test_frame_count_matches_stft_without_end_padding():
    assert L.num_frames(399) == 0 and L.num_frames(400) == 1 and L.num_frames(559) == 1 and L.num_frames(560) == 2
    assert L.num_frames(480000) == 2998        # 30 s -> 2998 frames


def test_mel_matrix_tf_properties():
    w = L.linear_to_mel_weight_matrix()
    assert w.shape == (201, 80)
    assert np.all(w[0] == 0.0)                 # DC bin zeroed
    assert w.min() >= 0.0 and w.max() <= 1.0   # un-normalised triangles (peak <= 1)
    # triangles overlap so that interior bins sum to 1 across neighbouring filters (mel-domain linear interpolation)
    assert np.allclose(w[5:190].sum(1), 1.0, atol=1e-12)
    # HTK formula
    assert abs(L.hertz_to_mel(1000.0) - 1127.0 * np.log(1.0 + 1000.0 / 700.0)) < 1e-12


def test_periodic_hann():
    h = L.hann_periodic(400)
    assert h[0] == 0.0 and abs(h[200] - 1.0) < 1e-15 and abs(h[1] - h[399]) < 1e-15   # periodic: h[n] = h[400 - n]


def test_pure_tone_known_answer():
    """A bin-centred cosine of amplitude A: the Hann-windowed power at its bin is (A * 400 / 4)^2 = (100 A)^2."""
    n = np.arange(16000)
    k = 50                                     # 2 kHz
    x = 0.5 * np.cos(2 * np.pi * k * n / 400.0)
    frames = x[np.arange(400)[None, :] + 160 * np.arange(L.num_frames(16000))[:, None]] * L.hann_periodic()
    p = np.abs(np.fft.rfft(frames, axis=-1)) ** 2
    assert np.allclose(p[:, k], (100 * 0.5) ** 2, rtol=1e-9)
    y = L.extract_fbank_features(x)
    w = L.linear_to_mel_weight_matrix()
    expect = np.log((p @ w) + 1e-6)
    assert np.allclose(y, expect, atol=1e-12)
    # silence -> log(1e-6) everywhere
    assert np.allclose(L.extract_fbank_features(np.zeros(1000)), np.log(1e-6))


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,mel_major", [(1, 400, False), (2, 16000, False), (3, 32000 + 77, True), (2, 480000, True)])
def test_logmel_kernel_matches_oracle(B, N, mel_major):
    import torch
    from tethys_speech_b200 import frontend

    rng = np.random.default_rng(B * 1000 + N)
    x = rng.standard_normal((B, N)).astype(np.float32)
    x[0, : N // 2] *= 0.01                      # wide dynamic range inside one batch
    y = frontend.extract_fbank_features(torch.from_numpy(x), mel_major=mel_major).cpu().double().numpy()
    ref = L.extract_fbank_features(x.astype(np.float64))
    if mel_major:
        ref = np.swapaxes(ref, -1, -2)
    assert y.shape == ref.shape
    # fp32 FFT vs fp64: compare mel POWER relatively where it is above the 1e-6 floor, and the logs absolutely
    assert np.max(np.abs(y - ref)) < 2e-4, np.max(np.abs(y - ref))
    big = ref > np.log(1e-3)
    assert np.max(np.abs(np.expm1(y[big] - ref[big]))) < 1e-4


@pytest.mark.gpu
def test_logmel_short_signal_and_bf16_and_unsupported():
    import torch
    from tethys_speech_b200 import _lib, frontend

    assert frontend.extract_fbank_features(torch.zeros(1, 399)).shape == (1, 0, 80)
    x = torch.randn(2, 8000)
    a = frontend.extract_fbank_features(x).float()
    b = frontend.extract_fbank_features(x, dtype=torch.bfloat16).float()
    assert float((a - b).abs().max()) < 0.05
    with pytest.raises(_lib.TethysError):
        frontend.extract_fbank_features(x, n_mels=64)
