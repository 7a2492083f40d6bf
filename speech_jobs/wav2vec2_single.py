#!/usr/bin/env python
"""Drop-in for the reference's speech_jobs/wav2vec2_single.py (VS:1281-1292): --batch_size 1, --num_batches 5,
--model_size small, --model_type pretraining, --learning_rate 3e-5, --num_epochs 1."""
import argparse
import time

import _path  # noqa: F401
from tethys_speech_b200 import train

if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="wav2vec2 single-GPU training")
    parser.add_argument("--num_batches", type=int, default=5)
    parser.add_argument("--batch_size", type=int, default=1)
    parser.add_argument("--model_size", type=str, default="small", choices=["tiny", "small", "base", "large"])
    parser.add_argument("--model_type", type=str, default="pretraining", choices=["pretraining", "asr", "classification"])
    parser.add_argument("--learning_rate", type=float, default=3e-5)
    parser.add_argument("--num_epochs", type=int, default=1)
    parser.add_argument("--audio_length", type=int, default=32000, help="extension: samples per clip (reference: 32000)")
    parser.add_argument("--precision", type=str, default="bf16", choices=["bf16", "fp32"])
    parser.add_argument("--resume", type=str, default=None, help="extension (SURVEY f-3): checkpoint file or directory to restore model + optimizer from")
    parser.add_argument("--no_cuda_graph", action="store_true", help="extension: launch every step kernel by kernel instead of replaying the captured CUDA graph")
    args = parser.parse_args()
    strategy = train.make_strategy()
    start = time.time()
    train.train_wav2vec2(strategy, args.model_type, args.model_size, num_epochs=args.num_epochs, learning_rate=args.learning_rate,
                         batch_size=args.batch_size, num_batches=args.num_batches, precision=args.precision, audio_length=args.audio_length,
                         resume_from=args.resume, cuda_graph=not args.no_cuda_graph)
    print("Training completed.")
    print("jct:", time.time() - start)
