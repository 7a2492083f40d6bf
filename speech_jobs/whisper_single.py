#!/usr/bin/env python
"""Single-GPU Whisper train script: --batch_size 4 --num_batches 40 defaults (WS:1303-1306).
NOTE: the reference file of this name actually contains Wav2Vec2-base pre-training on 5 s audio with the
legacy sampler/step. BASELINE config #1 describes Whisper, so the default here is whisper_dist.py's model and step without
a strategy; `--legacy_wav2vec2` runs what the reference file literally does (Wav2Vec2-base, 80 000 samples, legacy step)."""
import argparse
import time

import _path  # noqa: F401
from tethys_speech_b200 import train

if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="Whisper single-GPU training")
    parser.add_argument("--num_batches", type=int, default=40)
    parser.add_argument("--batch_size", type=int, default=4)
    parser.add_argument("--legacy_wav2vec2", action="store_true", help="literal behaviour of the reference file (WS:1183-1258)")
    parser.add_argument("--precision", type=str, default="bf16", choices=["bf16", "fp32"])
    parser.add_argument("--from_waveform", action="store_true", help="feed raw 30 s waveforms through the fused log-mel kernel (W:739-766) instead of ready-made mel features")
    parser.add_argument("--resume", type=str, default=None, help="extension (SURVEY f-3): checkpoint file or directory to restore model + optimizer from")
    parser.add_argument("--no_cuda_graph", action="store_true", help="extension: launch every step kernel by kernel instead of replaying the captured CUDA graph")
    args = parser.parse_args()
    strategy = train.make_strategy()
    start = time.time()
    if args.legacy_wav2vec2:
        train.train_wav2vec2(strategy, "pretraining", "base", batch_size=args.batch_size, num_batches=args.num_batches,
                             precision=args.precision, audio_length=80000, legacy=True, resume_from=args.resume, cuda_graph=not args.no_cuda_graph)
    else:
        train.train_whisper(strategy, "small", batch_size=args.batch_size, num_batches=args.num_batches, precision=args.precision,
                            from_waveform=args.from_waveform, resume_from=args.resume, cuda_graph=not args.no_cuda_graph)
    print("Training completed.")
    print("jct:", time.time() - start)
