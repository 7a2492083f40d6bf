#!/usr/bin/env python
"""BASELINE config 5 — front-end microbench sweep: the log-mel kernel (K1, W:739-766) and the Wav2Vec2 conv feature
encoder forward (K4-K7, V:283-298) over batch x seconds of 16 kHz audio, against the HBM roofline (and, for the conv
encoder, the tensor roofline: it sits at the ridge). One JSON line per cell -> profiles/.
  python tools/frontend_bench.py [--out profiles/r01_frontend_sweep.jsonl] [--quick]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tethys_speech_b200 import frontend
from tethys_speech_b200 import wav2vec2 as W


def timed(fn, flush, iters=5, warm=2):
    for _ in range(warm):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="profiles/r01_frontend_sweep.jsonl")
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    batches = [1, 8, 64] if args.quick else [1, 2, 4, 8, 16, 32, 64, 128, 256]
    seconds = [1, 5, 30] if args.quick else [1, 2, 5, 10, 15, 30]
    model = W.Wav2Vec2ForPreTraining(W.Wav2Vec2Config("base"), precision="bf16", device=0, seed=0)
    lines = []
    for B in batches:
        for sec in seconds:
            N = 16000 * sec
            x = torch.randn(B, N, device=dev)
            # ---- log-mel: 4 N + 4 * 80 * F bytes per sample (fp32 out) ----
            F = frontend.num_frames(N)
            t = timed(lambda: frontend.extract_fbank_features(x, mel_major=True), flush)
            nbytes = B * (4.0 * N + 4.0 * 80 * F)
            lines.append({"kernel": "logmel", "batch": B, "seconds": sec, "us": t * 1e6, "GBs": nbytes / t / 1e9,
                          "frac_hbm": nbytes / t / 1e9 / peaks["hbm_gbs"], "audio_sec_per_sec": B * sec / t})
            # ---- conv feature encoder fwd: (404.8 e + 4) N bytes, 24.53 GFLOP per 5 s (+ pos-conv 1.05) per sample, bf16 ----
            act_bytes = B * 512.0 * N * 2 * (2 * 0.396875 - 1 / 320.0) + B * 4.0 * N
            if act_bytes > 60e9:
                continue
            try:
                t = timed(lambda: model.extract_features(x), flush, iters=3, warm=1)
            except Exception as e:  # workspace beyond the card
                lines.append({"kernel": "conv_feature_encoder_fwd", "batch": B, "seconds": sec, "skipped": str(e)[:80]})
                continue
            flops = B * (24.53e9 + 1.05e9) * sec / 5.0
            lines.append({"kernel": "conv_feature_encoder_fwd", "batch": B, "seconds": sec, "us": t * 1e6,
                          "GBs": act_bytes / t / 1e9, "frac_hbm": act_bytes / t / 1e9 / peaks["hbm_gbs"],
                          "TFLOPs": flops / t / 1e12, "frac_bf16": flops / t / 1e12 / peaks["bf16_tflops"],
                          "audio_sec_per_sec": B * sec / t})
            del x
            torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        for ln in lines:
            f.write(json.dumps(ln) + "\n")
            print(json.dumps(ln))


if __name__ == "__main__":
    main()
