#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the tethys-speech hot path (data-parallel train step).

  python bench.py --gpus N --steps K --warmup W            # our arm (torchrun launches N>1)
  python bench.py --impl reference --gpus N --steps K ...  # the reference arm: CPU restatement on host cores

Workload at N=1 = BASELINE.json configs[1]: Wav2Vec2-base pre-training step (wav2vec2_single.py, VS:1119-1176) on
synthetic 16 kHz 15 s waveforms, bf16 compute with fp32 master weights/Adam, dropout ON (training=True as in the
reference).  For N>1 the same per-GPU work runs under wav2vec2_dist.py's step (V:1186-1260: loss/N, local
global-norm clip, NCCL all-reduce SUM, per-variable clipnorm, Adam) — weak scaling.

One "step" = one full train step (forward, loss, backward, clip, Adam) on one batch of synthetic audio.
`value`  : samples/s with the batches already resident in HBM.
`e2e`    : samples/s through the public host API with pinned-host -> device copies of every batch and a
           device -> host read of the loss inside the timed region.
`roofline`: the dominant kernel (tcgen05 GEMM, FFN shape of this workload) timed live with CUDA events.
`cpu_baseline`: the oracle (PyTorch-CPU fp32 restatement of the identical step; TensorFlow is not installable) timed
           on this box's host cores on a bounded sample — a reported baseline, not the target.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (model_size, samples, seconds, train GFLOP/sample)
    "w2v_base_15s": ("base", 240000, 15.0, 679.4),
    "w2v_base_5s": ("base", 80000, 5.0, 212.7),
    "w2v_base_2s": ("base", 32000, 2.0, 83.4),
    "w2v_tiny_2s": ("tiny", 32000, 2.0, 25.6),
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "bf16_tflops": d.get("bf16_tflops", 1590.0),
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", 1400.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's train step on host cores
# ------------------------------------------------------------------------------------------------------------
def cpu_step_throughput(workload, steps, warmup, max_seconds=150.0):
    import torch
    from oracle import wav2vec2_oracle as O

    size, n_samples, secs, _ = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.Wav2Vec2Config(size)
    w = O.init_weights(cfg, seed=0, dtype=torch.float32)
    m = {k: torch.zeros_like(v) for k, v in w.items()}
    v = {k: torch.zeros_like(v_) for k, v_ in w.items()}
    g = torch.Generator().manual_seed(1234)
    B = 1
    T = O.num_frames(cfg, n_samples)
    times = []
    t_all = time.perf_counter()
    for it in range(warmup + steps):
        wave = torch.randn(B, n_samples, generator=g)
        neg = O.negative_indices_from_random(torch.randint(0, T, (B, T), generator=g), cfg.num_negatives)
        t0 = time.perf_counter()
        O.train_step(cfg, w, m, v, it + 1, wave, neg)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        if time.perf_counter() - t_all > max_seconds and len(times) >= 1:
            break
    mean = sum(times) / len(times)
    cpu_model = "unknown"
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                cpu_model = ln.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    return {"value": B / mean, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": f"{len(times)} timed step(s) of batch {B} x {secs:g} s audio ({workload}), PyTorch-CPU fp32 restatement "
                      f"of the reference step (TensorFlow not installable), cpu='{cpu_model}'",
            "ms_per_step": mean * 1e3, "steps_timed": len(times)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    size, n_samples, secs, gflop = WORKLOADS[args.workload]
    r = cpu_step_throughput(args.workload, max(1, min(args.steps, 3)), min(args.warmup, 1))
    line = {"impl": "reference", "metric": "train_samples_per_sec", "value": r["value"], "unit": "samples/s",
            "audio_sec_per_sec": r["value"] * secs, "n_gpus": args.gpus, "steps": r["steps_timed"], "warmup": min(args.warmup, 1),
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "model": f"wav2vec2-{size}", "audio_seconds": secs, "batch_per_step": 1,
                       "note": "reference arm = CPU restatement of the reference's TF step on host cores (rank 0 only)"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def time_dominant_gemm(ctx, M, N, K, iters=20):
    """Time the workload's dominant GEMM (FFN fc1: [M,K]x[K,N] + bias + GELU, bf16) alone with CUDA events on the
    launching stream, flushing L2 between launches."""
    import ctypes as C

    import torch
    from tethys_speech_b200 import _lib
    from tethys_speech_b200.runtime import ptr, stream_ptr

    dev = torch.device("cuda", ctx.device)
    a = torch.randn(M, K, device=dev).bfloat16()
    b = (torch.randn(K, N, device=dev) * 0.05).bfloat16()
    bias = torch.zeros(N, device=dev)
    c = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    d = _lib.GemmDesc()
    d.a, d.b, d.c = a.data_ptr(), b.data_ptr(), c.data_ptr()
    d.m, d.n, d.k = M, N, K
    d.a_major, d.b_major = 0, 1
    d.lda, d.ldb, d.ldc = K, N, N
    d.batch1 = d.batch2 = 1
    d.in_dtype = d.out_dtype = _lib.TS_BF16
    d.alpha = 1.0
    d.bias = bias.data_ptr()
    d.act = 1
    d.c_preact = pre.data_ptr()
    for _ in range(3):
        ctx.check(ctx.lib.ts_gemm(ctx.h, C.byref(d), stream_ptr()))
    total = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.check(ctx.lib.ts_gemm(ctx.h, C.byref(d), stream_ptr()))
        e1.record()
        e1.synchronize()
        total += e0.elapsed_time(e1)
    return total / iters * 1e-3


def run_ours(args):
    import torch
    from tethys_speech_b200 import wav2vec2 as W
    from tethys_speech_b200.runtime import Strategy

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (our arm) needs a B200; there is no CPU fallback"
    torch.cuda.set_device(local)
    strategy = Strategy()
    size, n_samples, secs, gflop = WORKLOADS[args.workload]
    B = args.batch
    dev = torch.device("cuda", local)

    cfg = W.Wav2Vec2Config(size)
    model = W.Wav2Vec2ForPreTraining(cfg, precision=args.precision, device=local, seed=0)
    model.broadcast_weights(strategy)
    opt = W.Adam(learning_rate=3e-5, epsilon=1e-8, clipnorm=1.0)
    ctx = model._prog.ctx

    # synthetic data: N(0,1) waveforms, rng = default_rng(1234 + rank); a pool of distinct batches
    import numpy as np

    rng = np.random.default_rng(1234 + rank)
    npool = 4
    host = [torch.from_numpy(rng.standard_normal((B, n_samples), dtype=np.float32)).pin_memory() for _ in range(npool)]
    resident = [h.to(dev) for h in host]

    def step(features):
        if world > 1:
            return W.distributed_train_step(strategy, model, (features, None), opt)
        return W.train_step(model, (features, None), opt)

    def sync_all():
        torch.cuda.synchronize()
        strategy.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------------------------
    for i in range(args.warmup):
        step(resident[i % npool])
    sync_all()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = ctx.lib.ts_launch_count(ctx.h)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = step(resident[i % npool])
    e1.record()
    sync_all()
    launches = int(ctx.lib.ts_launch_count(ctx.h) - l0)
    t_dev = e0.elapsed_time(e1) * 1e-3
    clk = clocks.stop() if rank == 0 else None
    # ---- end-to-end timing: pinned host -> device copy of every batch, loss read back every step ----------
    staging = torch.empty(B, n_samples, dtype=torch.float32, device=dev)
    for i in range(2):
        staging.copy_(host[i % npool], non_blocking=True)
        float(step(staging))
    sync_all()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    last = None
    for i in range(args.steps):
        staging.copy_(host[i % npool], non_blocking=True)
        last = float(step(staging))          # device -> host read of the step's loss
    e3.record()
    sync_all()
    t_e2e = e2.elapsed_time(e3) * 1e-3
    ctx.watchdog()
    # max over ranks
    if world > 1:
        tt = torch.tensor([t_dev, t_e2e], device=dev, dtype=torch.float64)
        strategy.dist.all_reduce(tt, op=strategy.dist.ReduceOp.MAX)
        t_dev, t_e2e = float(tt[0]), float(tt[1])
    if rank != 0:
        return 0
    peaks = measured_peaks()
    sps = B * world * args.steps / t_dev
    sps_e2e = B * world * args.steps / t_e2e
    T = model.num_frames(n_samples)
    M, Hd, F = B * T, cfg.hidden_size, cfg.intermediate_size
    t_gemm = time_dominant_gemm(ctx, M, F, Hd)
    gemm_tf = 2.0 * M * F * Hd / t_gemm / 1e12
    step_tflops = sps / world * gflop / 1e3
    line = {
        "metric": "train_samples_per_sec", "value": sps, "unit": "samples/s", "audio_sec_per_sec": sps * secs,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": args.workload, "model": f"wav2vec2-{size}", "audio_seconds": secs, "per_gpu_batch": B,
                   "global_batch": B * world, "parallelism": f"dp{world}", "dropout": "on (0.1, as the reference's training=True)",
                   "step": "VS:1119-1176 (clip_by_global_norm 1.0 + clipnorm 1.0 + Keras-legacy Adam 3e-5)" if world == 1
                   else "V:1186-1260 (loss/N, local clip, NCCL all-reduce SUM, clipnorm, Adam)",
                   "l2": "working set per step (GBs of activations) >> 126 MB L2; 4 distinct input batches cycled"},
        "clocks": clk,
        "e2e": {"value": sps_e2e, "unit": "samples/s", "h2d_bytes_per_step": B * n_samples * 4 + B * cfg.num_negatives * 4,
                "d2h_bytes_per_step": 4, "ms_per_step": t_e2e / args.steps * 1e3, "last_loss": last},
        "gpu_launches": launches,
        "launches_per_step": launches / args.steps,
        "step_tflops_per_gpu": step_tflops,
        "step_frac_of_bf16_sustained": step_tflops / peaks["bf16_tflops_sustained"],
        "roofline": {"bound": "tensor", "kernel": f"gemm_tc_kernel (FFN fc1 {M}x{F}x{Hd}, bias+GELU epilogue)", "achieved": gemm_tf,
                     "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": gemm_tf / peaks["bf16_tflops"], "traffic": None,
                     "peak_source": peaks["source"] + " (burst: kernel timed alone)"},
    }
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_step_throughput(args.workload, 1, 1, max_seconds=120.0)
        line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="w2v_base_15s", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=8, help="per-GPU batch (--batch_size of the reference CLI)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
