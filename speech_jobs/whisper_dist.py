#!/usr/bin/env python
"""Drop-in for the reference's speech_jobs/whisper_dist.py (W:1029-1058): same flags and defaults
(--num_batches 40, --batch_size 1 per replica), model_type "small" hard-coded as in W:1005, same log lines and jct file.
Launch one process per GPU:  torchrun --nproc-per-node N speech_jobs/whisper_dist.py --batch_size 4 --num_batches 30"""
import argparse
import time

import _path  # noqa: F401
from tethys_speech_b200 import train


def main(strategy, args, task_type, task_index):
    print("Whisper 분산 학습 시작...")
    start_time = time.time()
    train.train_whisper(strategy, "small" if args.model_type is None else args.model_type, batch_size=args.batch_size,
                        num_batches=args.num_batches, precision=args.precision, from_waveform=args.from_waveform,
                        resume_from=args.resume, cuda_graph=not args.no_cuda_graph)
    jct = time.time() - start_time
    print("Training completed.")
    if strategy.rank == 0:
        train.write_jct(jct, task_type, task_index)


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="Whisper Distributed Speech Recognition")
    parser.add_argument("--num_batches", type=int, default=40, help="num_batches per replica, default is set 40")
    parser.add_argument("--batch_size", type=int, default=1, help="batch size per replica, default is set 1")
    parser.add_argument("--model_type", type=str, default=None, help="extension: tiny|base|small|medium|large (reference: small)")
    parser.add_argument("--precision", type=str, default="bf16", choices=["bf16", "fp32"], help="extension: compute precision")
    parser.add_argument("--from_waveform", action="store_true",
                        help="extension (SURVEY f-1): feed raw 30 s waveforms through the fused log-mel kernel (W:739-766)")
    parser.add_argument("--resume", type=str, default=None, help="extension (SURVEY f-3): checkpoint file or directory to restore model + optimizer from")
    parser.add_argument("--no_cuda_graph", action="store_true", help="extension: launch every step kernel by kernel instead of replaying the captured CUDA graph")
    args = parser.parse_args()
    task_type, task_index = train.task_from_tf_config()
    strategy = train.make_strategy()
    print(f"batch size per replica: {args.batch_size}, global batch size: {args.batch_size * strategy.num_replicas_in_sync}")
    print(f"num_batches: {args.num_batches}")
    main(strategy, args, task_type, task_index)
