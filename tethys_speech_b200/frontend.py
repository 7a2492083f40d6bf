"""Audio front end with the reference's surface: extract_fbank_features (W:739-766) on the fused log-mel kernel (K1)."""
import ctypes as C

import torch

from . import _lib
from .runtime import stream_ptr, to_device


def num_frames(n_samples, n_fft=400, hop_length=160):
    """tf.signal.stft(..., pad_end=False): 1 + (N - n_fft) // hop frames."""
    return 0 if n_samples < n_fft else 1 + (n_samples - n_fft) // hop_length


def extract_fbank_features(waveform, sample_rate=16000, n_mels=80, n_fft=400, hop_length=160, mel_major=False,
                           dtype=torch.float32, device=None):
    """waveform [N] or [B, N] (torch / numpy / DLPack) -> log-mel [B?, frames, n_mels] fp32, exactly the reference's
    tf.signal.stft -> |.|^2 -> linear_to_mel_weight_matrix -> log(. + 1e-6) chain. mel_major=True returns [B?, n_mels,
    frames], the layout WhisperEncoder.call expects (W:326-329). Only the reference configuration is implemented; anything
    else raises (no fallback)."""
    if (sample_rate, n_mels, n_fft, hop_length) != (16000, 80, 400, 160):
        raise _lib.TethysError(-6, "extract_fbank_features: only sample_rate=16000, n_mels=80, n_fft=400, hop_length=160")
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    x = to_device(waveform, torch.float32, dev)
    squeeze = x.dim() == 1
    if squeeze:
        x = x.unsqueeze(0)
    B, N = x.shape
    F = num_frames(N)
    shape = (B, n_mels, F) if mel_major else (B, F, n_mels)
    out = torch.empty(shape, dtype=dtype, device=dev)
    if F > 0:
        ctx = _lib.context(dev.index)
        dt = _lib.TS_F32 if dtype == torch.float32 else _lib.TS_BF16
        ctx.check(ctx.lib.ts_logmel(ctx.h, C.c_void_p(x.data_ptr()), x.stride(0), B, N, C.c_void_p(out.data_ptr()), dt,
                                    1 if mel_major else 0, stream_ptr()))
    return out[0] if squeeze else out
