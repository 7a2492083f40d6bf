"""TEST INFRASTRUCTURE ONLY — NumPy restatement of the reference's span-mask utilities
apply_time_mask / apply_feature_mask (/root/reference/speech_jobs/wav2vec2_dist.py:1073-1095, 1098-1120).

The reference draws the span STARTS with tf.random.uniform(shape) < mask_prob (V:1078, V:1103) — a TF RNG stream that
cannot be reproduced — and then dilates every start to the right by mask_length positions with a loop of shifted ORs
(V:1083-1086, V:1108-1111). The integer work pinned here is the dilation; the starts are an input.
(The reference defines these functions but never calls them — SURVEY D5. Pinned bit-exactly against the reference functions run on
oracle/tf_shim.py: tests/test_reference_pinning.py::test_span_masks_match_the_reference_functions.)"""
import numpy as np


def expand_spans(start_mask, mask_length):
    """start_mask bool [B, L] -> expanded bool [B, L]: expanded[b, t] = OR_{i < mask_length} start[b, t - i]  (V:1083-1086)."""
    start = np.asarray(start_mask, dtype=bool)
    L = start.shape[1]
    out = np.zeros_like(start)
    for i in range(mask_length):
        shifted = np.zeros_like(start)
        if i < L:
            shifted[:, i:] = start[:, :L - i]          # tf.pad(mask[:, :L-i], [[0,0],[i,0]])
        out |= shifted
    return out


def apply_time_mask(hidden_states, start_mask, mask_length=10):
    """hidden_states [B, T, H], start_mask [B, T] -> (masked [B, T, H], expanded_mask float32 [B, T, 1])  (V:1089-1095)."""
    m = expand_spans(start_mask, mask_length).astype(np.float32)[:, :, None]
    return np.asarray(hidden_states) * (1.0 - m), m


def apply_feature_mask(hidden_states, start_mask, mask_length=10):
    """hidden_states [B, T, H], start_mask [B, H] -> (masked [B, T, H], expanded_mask float32 [B, 1, H])  (V:1114-1120)."""
    m = expand_spans(start_mask, mask_length).astype(np.float32)[:, None, :]
    return np.asarray(hidden_states) * (1.0 - m), m
