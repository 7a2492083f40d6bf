"""TEST INFRASTRUCTURE ONLY — NumPy fp64 restatement of the reference's log-mel front end,
`extract_fbank_features` (/root/reference/speech_jobs/whisper_dist.py:739-766):

    stfts = tf.signal.stft(waveform, frame_length=400, frame_step=160, fft_length=400)      # W:744-749
    power = tf.math.square(tf.abs(stfts))                                                   # W:752
    mel_w = tf.signal.linear_to_mel_weight_matrix(80, 201, 16000, 0, 8000)                  # W:755-758
    mel   = tf.tensordot(power, mel_w, 1)                                                   # W:761
    out   = tf.math.log(mel + 1e-6)                                                         # W:764   -> [frames, 80]

TensorFlow is not importable here; the call sequence is pinned against the reference function run on oracle/tf_shim.py
(tests/test_reference_pinning.py::test_logmel_front_end_matches_the_reference_function); the op semantics restated below are those of TF 2.10
(SURVEY App. A-9): `stft` with pad_end=False -> frames = 1 + (N - 400) // 160; window = PERIODIC Hann
0.5 - 0.5 cos(2 pi n / 400); rfft of length 400 -> 201 bins; `linear_to_mel_weight_matrix`: HTK mel scale
1127 ln(1 + f / 700), 82 band edges linear in mel between mel(0) and mel(8000), triangles evaluated in the MEL domain,
not normalised, DC bin zeroed.
"""
import numpy as np


def hertz_to_mel(f):
    return 1127.0 * np.log1p(np.asarray(f, dtype=np.float64) / 700.0)


def linear_to_mel_weight_matrix(num_mel_bins=80, num_spectrogram_bins=201, sample_rate=16000, lower=0.0, upper=8000.0):
    """tf.signal.linear_to_mel_weight_matrix — [num_spectrogram_bins, num_mel_bins]."""
    bands_to_zero = 1
    nyquist = sample_rate / 2.0
    linear_freqs = np.linspace(0.0, nyquist, num_spectrogram_bins)[bands_to_zero:]
    spec_mel = hertz_to_mel(linear_freqs)[:, None]
    edges = np.linspace(hertz_to_mel(lower), hertz_to_mel(upper), num_mel_bins + 2)
    lower_edge, center, upper_edge = edges[None, :-2], edges[None, 1:-1], edges[None, 2:]
    lower_slopes = (spec_mel - lower_edge) / (center - lower_edge)
    upper_slopes = (upper_edge - spec_mel) / (upper_edge - center)
    w = np.maximum(0.0, np.minimum(lower_slopes, upper_slopes))
    return np.pad(w, [[bands_to_zero, 0], [0, 0]])


def hann_periodic(n=400):
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def num_frames(n_samples, frame_length=400, frame_step=160):
    return 0 if n_samples < frame_length else 1 + (n_samples - frame_length) // frame_step


def extract_fbank_features(waveform, sample_rate=16000, n_mels=80, n_fft=400, hop_length=160):
    """waveform [..., N] -> log-mel [..., frames, n_mels] (fp64)."""
    x = np.asarray(waveform, dtype=np.float64)
    nf = num_frames(x.shape[-1], n_fft, hop_length)
    idx = np.arange(n_fft)[None, :] + hop_length * np.arange(nf)[:, None]
    frames = x[..., idx] * hann_periodic(n_fft)
    power = np.abs(np.fft.rfft(frames, n=n_fft, axis=-1)) ** 2
    mel = power @ linear_to_mel_weight_matrix(n_mels, n_fft // 2 + 1, sample_rate, 0.0, sample_rate / 2.0)
    return np.log(mel + 1e-6)
