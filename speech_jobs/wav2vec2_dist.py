#!/usr/bin/env python
"""Drop-in for the reference's speech_jobs/wav2vec2_dist.py (V:1443-1487): --num_batches 5, --batch_size 1 per replica,
--model_size {tiny,small,base} (default small); 2 s synthetic audio (V:1129). One process per GPU under torchrun."""
import argparse
import time

import _path  # noqa: F401
from tethys_speech_b200 import train

if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="wav2vec2 Distributed Speech Recognition")
    parser.add_argument("--num_batches", type=int, default=5, help="num_batches per replica, default is set 5")
    parser.add_argument("--batch_size", type=int, default=1, help="batch size per replica, default is set 1")
    parser.add_argument("--model_size", type=str, default="small", choices=["tiny", "small", "base", "large"])
    parser.add_argument("--audio_length", type=int, default=32000, help="extension: samples per clip (reference: 32000)")
    parser.add_argument("--precision", type=str, default="bf16", choices=["bf16", "fp32"])
    parser.add_argument("--resume", type=str, default=None, help="extension (SURVEY f-3): checkpoint file or directory to restore model + optimizer from")
    parser.add_argument("--no_cuda_graph", action="store_true", help="extension: launch every step kernel by kernel instead of replaying the captured CUDA graph")
    args = parser.parse_args()
    task_type, task_index = train.task_from_tf_config()
    strategy = train.make_strategy()
    print(f"선택된 모델 크기: {args.model_size}")
    print(f"batch size per replica: {args.batch_size}, global batch size: {args.batch_size * strategy.num_replicas_in_sync}")
    print(f"num_batches: {args.num_batches}")
    start = time.time()
    train.train_wav2vec2(strategy, "pretraining", args.model_size, batch_size=args.batch_size, num_batches=args.num_batches,
                         precision=args.precision, audio_length=args.audio_length, resume_from=args.resume, cuda_graph=not args.no_cuda_graph)
    jct = time.time() - start
    print("Training completed.")
    if strategy.rank == 0:
        train.write_jct(jct, task_type, task_index)
