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
           device -> host read of every step's loss inside the timed region (two-deep input pipeline: the copy of
           batch i+1 runs on a copy stream under step i, the loss of step i is read while step i+1 runs).
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
    # name: (family, model_size, samples (w2v) / mel frames (whisper), audio seconds, train GFLOP/sample)
    "w2v_base_15s": ("w2v", "base", 240000, 15.0, 679.4),     # BASELINE.json configs[1] — the default
    "w2v_base_5s": ("w2v", "base", 80000, 5.0, 212.7),        # configs[0] as the reference file literally is
    "w2v_base_2s": ("w2v", "base", 32000, 2.0, 83.4),
    "w2v_tiny_2s": ("w2v", "tiny", 32000, 2.0, 25.6),
    "w2v_large_15s": ("w2v", "large", 240000, 15.0, 1769.8),  # configs[3]: EXTRAPOLATED preset, no reference numerics
    "whisper_small_30s": ("whisper", "small", 3000, 30.0, 449.1),   # configs[0] as BASELINE intends (1a): CLI default preset
    "whisper_base_30s": ("whisper", "base", 3000, 30.0, 325.5),     # configs[2]
}
DEFAULT_BATCH = {"w2v": 8, "whisper": 4}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "bf16_tflops": d.get("bf16_tflops", 1590.0),
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", 1400.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every 10 ms from a thread (a 20-step timed
    region lasts ~0.1-0.3 s, less than `nvidia-smi -lms` needs to start), nvidia-smi as the fallback."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines, self.sm, self.mask = [], [], 0
        self.proc = self.nv = self.h = self.t = None
        self.mx = None
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv is not None:
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
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
        if self.nv is not None:
            self._stop.set()
            self.t.join(timeout=1)
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx,
                    "reasons": sorted(name for bit, name in self.REASONS if self.mask & bit), "samples": len(sm), "source": "nvml"}
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
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's train step on host cores
# ------------------------------------------------------------------------------------------------------------
STEP_DESC = {
    ("w2v", True): "VS:1119-1176 (clip_by_global_norm 1.0 + clipnorm 1.0 + Keras-legacy Adam 3e-5)",
    ("w2v", False): "V:1186-1260 (loss/N, local clip, NCCL all-reduce SUM, clipnorm, Adam)",
    ("whisper", True): "W:823-836 single replica (forward, shifted CE, backward, Keras-legacy Adam 1e-4)",
    ("whisper", False): "W:819-848 (un-normalised NCCL all-reduce SUM in buckets overlapped with backward, Adam)",
}


def workload_config(workload, batch, world):
    """The `config` object of the JSON line: it NAMES THE WORKLOAD (a BASELINE.json config at N replicas) and nothing about how an
    arm runs it, so that both arms — ours and `--impl reference` — print the same object for the same command line. What is
    specific to a run goes elsewhere: ours in `run` (CUDA graph, collective layer), the reference arm's in `cpu_baseline.sample`."""
    family, size, n_samples, secs, _ = WORKLOADS[workload]
    B = batch or DEFAULT_BATCH[family]
    return {"workload": workload, "model": f"{'wav2vec2' if family == 'w2v' else 'whisper'}-{size}", "audio_seconds": secs,
            "per_gpu_batch": B, "global_batch": B * world, "parallelism": f"dp{world}",
            "dropout": "on (0.1, as the reference's training=True)", "step": STEP_DESC[(family, world == 1)],
            "allreduce": "none (1 replica)" if world == 1 else "SUM of the replicas' gradients inside apply_gradients (W:834 / V:1246)",
            "l2": "working set per step (GBs of activations) >> 126 MB L2; 4 distinct input batches cycled; "
                  "per-kernel timings: CUDA-graph replays of [256 MB memset (L2 flush); kernel] minus replays of the memset alone"}


def cpu_step_throughput(workload, steps, warmup, max_seconds=150.0):
    """The oracle's PyTorch-CPU fp32 restatement of the same train step on all host cores, batch 1 (bounded sample)."""
    import numpy as np
    import torch

    family, size, n_samples, secs, _ = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(1234)
    B = 1
    if family == "w2v":
        from oracle import wav2vec2_oracle as O
        cfg = O.Wav2Vec2Config(size)
        T = O.num_frames(cfg, n_samples)

        def make():
            wave = torch.randn(B, n_samples, generator=g)
            neg = O.negative_indices_from_random(torch.randint(0, T, (B, T), generator=g), cfg.num_negatives)
            return wave, neg
    else:
        from oracle import whisper_oracle as O
        cfg = O.WhisperConfig(size)
        rng = np.random.default_rng(1234)

        def make():
            return torch.randn(B, cfg.n_mels, n_samples, generator=g), O.dummy_labels(rng, B, 100)
    w = O.init_weights(cfg, seed=0, dtype=torch.float32)
    m = {k: torch.zeros_like(v) for k, v in w.items()}
    v = {k: torch.zeros_like(v_) for k, v_ in w.items()}
    times = []
    t_all = time.perf_counter()
    for it in range(warmup + steps):
        a, b = make()
        t0 = time.perf_counter()
        O.train_step(cfg, w, m, v, it + 1, a, b)
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
            "sample": f"{len(times)} timed step(s) of batch {B} x {secs:g} s audio ({workload}: one sample of the per-GPU batch, dropout "
                      f"layers off), PyTorch-CPU fp32 restatement of the reference step (TensorFlow not installable), cpu='{cpu_model}'",
            "ms_per_step": mean * 1e3, "steps_timed": len(times)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    family, size, n_samples, secs, gflop = WORKLOADS[args.workload]
    r = cpu_step_throughput(args.workload, max(1, args.steps), max(0, args.warmup), max_seconds=150.0)   # K timed steps after W warm-up, capped at 150 s
    line = {"impl": "reference", "metric": "train_samples_per_sec", "value": r["value"], "unit": "samples/s",
            "audio_sec_per_sec": r["value"] * secs, "n_gpus": args.gpus, "steps": r["steps_timed"], "warmup": max(0, args.warmup),
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, args.batch, max(1, int(os.environ.get("WORLD_SIZE", "1")))),
            "note": "reference arm = CPU restatement of the reference's TF step on host cores (rank 0 only); every timed step is a bounded "
                    "sample of the workload `config` names (cpu_baseline.sample says which)",
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def _timed_warm(fn, iters=20, warm=3):
    """Device time (s) of fn()'s kernels replayed back to back as one CUDA graph of `iters` launches, operands L2-resident where they
    fit: what a GEMM sees inside the step when its producer has just written its input. Reported next to the cold-L2 figure."""
    import torch

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def _timed(fn, flush, iters=10, warm=3):
    """Average device time (s) of fn()'s kernels on the current stream with a cold L2. Two CUDA graphs are replayed `iters` times
    between CUDA events: [256 MB memset; fn()] and [256 MB memset]; the difference of the two totals / iters is the kernels' time
    as they run inside a step graph (node-to-node gap included, host launch latency and the flush itself excluded). The earlier
    event-around-one-launch timing carried 3-5 us of launch gap per sample — half of a 10 us LayerNorm."""
    import torch

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    g_with, g_without = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_with):
        flush.zero_()
        fn()
    with torch.cuda.graph(g_without):
        flush.zero_()
    tot = []
    for g in (g_with, g_without):
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            g.replay()
        e1.record()
        e1.synchronize()
        tot.append(e0.elapsed_time(e1))
    return max(tot[0] - tot[1], 1e-6) / iters * 1e-3


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` summary, or None."""
    p = os.path.join(ROOT, "profiles", "ncu_kernel_summary.json")
    try:
        return json.load(open(p)).get(key, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def frontend_rooflines(peaks, device_index, B=8, seconds=15):
    """BASELINE config 5 inside the bench line: the Wav2Vec2 conv feature encoder FORWARD as a whole (conv0 .. conv6 with
    GroupNorm + GELU, grouped positional conv, LayerNorm: Wav2Vec2FeatureExtractor.call V:283-298 through ts_w2v_forward_features)
    at the workload's shape. It sits at the ridge, so both fractions are reported: algorithmic bytes
    (404.8 e + 4) N per sample (every activation written once and read once by its consumer, e = 2 for bf16) against the measured
    copy bandwidth, and 2 x MACs (24.53 GFLOP per 5 s of audio + 1.05 of the positional conv) against the sustained bf16 peak.
    Timed with CUDA events around eager launches, L2 flushed before each sample (the ~20 kernels of the pass take >> their launch gaps)."""
    import ctypes as C

    import torch
    from tethys_speech_b200 import wav2vec2 as W
    from tethys_speech_b200.runtime import stream_ptr

    dev = torch.device("cuda", device_index)
    N = 16000 * seconds
    model = W.Wav2Vec2ForPreTraining(W.Wav2Vec2Config("base"), precision="bf16", device=device_index, seed=0)
    p = model._prog
    x = torch.randn(B, N, device=dev)
    model.extract_features(x)                                  # plans the workspace, syncs the bf16 weights
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    xp = C.c_void_p(x.data_ptr())
    for _ in range(2):
        p.ctx.check(p.lib.ts_w2v_forward_features(p.h, xp, B, N, stream_ptr()))
    torch.cuda.synchronize()
    tot, iters = 0.0, 5
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        p.ctx.check(p.lib.ts_w2v_forward_features(p.h, xp, B, N, stream_ptr()))
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    t = tot / iters * 1e-3
    nbytes = B * ((404.8 * 2 + 4.0) * N)
    flops = B * (24.53e9 + 1.05e9) * (seconds / 5.0)
    name = f"conv feature encoder forward as a whole [{B} x {seconds} s] bf16 (V:283-298, ts_w2v_forward_features)"
    return [{"kernel": name, "bound": "hbm", "achieved": nbytes / t / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
             "frac": nbytes / t / 1e9 / peaks["hbm_gbs"], "us": t * 1e6, "traffic": None,
             "tensor_tflops": flops / t / 1e12, "frac_of_bf16_sustained": flops / t / 1e12 / peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])}]


def kernel_rooflines(ctx, peaks, B, T, H, F, nh, n_params):
    """Live per-kernel roofline fractions at this workload's shapes, each kernel timed alone through the C-ABI.
    Tensor-bound kernels: algorithmic FLOPs / time vs the measured cuBLAS bf16 burst peak. HBM-bound kernels: algorithmic
    bytes (every distinct tensor once in, once out — SURVEY §8d) / time vs the measured copy bandwidth."""
    import ctypes as C

    import torch
    from tethys_speech_b200 import _lib
    from tethys_speech_b200.runtime import stream_ptr

    dev = torch.device("cuda", ctx.device)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    M = B * T
    out = []
    bf = torch.bfloat16

    def P(t):
        return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)

    def gemm(name, m, n, k, a_major, b_major, out_f32=False, bias=False, act=0, preact=False, res=False, drop=0.0, acc=False,
             lda=None, key=None):
        a_rows, a_cols = (m, k) if a_major == 0 else (k, m)
        b_rows, b_cols = (n, k) if b_major == 0 else (k, n)
        lda_ = lda or a_cols
        a = (torch.randn((a_rows - 1) * lda_ + a_cols + 64, device=dev) * 0.5).to(bf)
        bmat = (torch.randn(b_rows, b_cols, device=dev) * 0.05).to(bf)
        c = torch.zeros(m, n, device=dev, dtype=torch.float32 if out_f32 else bf)
        d = _lib.GemmDesc()
        d.a, d.b, d.c = a.data_ptr(), bmat.data_ptr(), c.data_ptr()
        d.m, d.n, d.k, d.a_major, d.b_major = m, n, k, a_major, b_major
        d.lda, d.ldb, d.ldc = lda_, b_cols, n
        d.batch1 = d.batch2 = 1
        d.in_dtype, d.out_dtype, d.alpha = _lib.TS_BF16, (_lib.TS_F32 if out_f32 else _lib.TS_BF16), 1.0
        keep = [a, bmat, c]
        if bias:
            bv = torch.zeros(n, device=dev); d.bias = bv.data_ptr(); keep.append(bv)
        if preact:
            pv = torch.empty_like(c); d.c_preact = pv.data_ptr(); keep.append(pv)
        if res:
            rv = torch.zeros_like(c); d.residual = rv.data_ptr(); d.ldr = n; keep.append(rv)
        d.act, d.drop, d.seed, d.accumulate = act, drop, 7, 1 if acc else 0
        t = _timed(lambda: ctx.check(ctx.lib.ts_gemm(ctx.h, C.byref(d), stream_ptr())), flush)
        tw = _timed_warm(lambda: ctx.check(ctx.lib.ts_gemm(ctx.h, C.byref(d), stream_ptr())))
        tf = 2.0 * m * n * k / t / 1e12
        out.append({"kernel": name, "bound": "tensor", "achieved": tf, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": tf / peaks["bf16_tflops"], "us": t * 1e6, "us_warm_l2": tw * 1e6, "frac_warm_l2": 2.0 * m * n * k / tw / 1e12 / peaks["bf16_tflops"],
                    "traffic": ncu_traffic(key) if key else None})
        return out[-1]

    dominant = gemm(f"gemm_tc_kernel FFN fc1 {M}x{F}x{H} +bias+GELU+dropout (W:194-195 / V:391-393)", M, F, H, 0, 1, bias=True, act=1,
                    preact=True, drop=0.1, key="gemm_ffn1")
    gemm(f"gemm_tc_kernel FFN fc2 {M}x{H}x{F} +bias+dropout+residual", M, H, F, 0, 1, bias=True, res=True, drop=0.1)
    gemm(f"gemm_tc_kernel QKV {M}x{3 * H}x{H} +bias", M, 3 * H, H, 0, 1, bias=True)
    gemm(f"gemm_tc_kernel dgrad {M}x{H}x{F}", M, H, F, 0, 0)
    gemm(f"gemm_tc_kernel wgrad {H}x{F}x{M} fp32 split-K", H, F, M, 1, 1, out_f32=True, acc=True)
    gemm(f"gemm_tc_kernel conv1 window-GEMM {B * 24000}x512x1536 (k3 s2, no im2col; V:254-268)", B * 24000, 512, 1536, 0, 1, lda=1024)

    # fused attention: the workload's own self-attention shape, plus Whisper's encoder self-attention (B4 H12 T1500) and cross-attention
    # (100 queries x 1500 keys) shapes
    def attn_case(tag, Bc, nhc, Tq, Tk, drop, ref, keyf=None, keyb=None):
        Hc = nhc * 64
        cross = Tq != Tk
        if cross:
            qt = torch.randn(Bc, Tq, Hc, device=dev).to(bf); kvt = torch.randn(Bc, Tk, 2 * Hc, device=dev).to(bf)
            dqt = torch.empty_like(qt); dkvt = torch.empty_like(kvt)
        else:
            qkv_ = torch.randn(Bc, Tq, 3 * Hc, device=dev).to(bf); dqkv_ = torch.empty_like(qkv_)
        o_ = torch.empty(Bc, Tq, Hc, device=dev, dtype=bf); olo_ = torch.empty_like(o_)
        do_ = torch.randn(Bc, Tq, Hc, device=dev).to(bf)
        stats_ = torch.empty(Bc, nhc, Tq, 2, device=dev); dsum_ = torch.empty(Bc, nhc, Tq, device=dev)
        dqacc_ = torch.empty(Bc, Tq, Hc, device=dev)           # fp32 dQ accumulator: selects the fused one-kernel backward
        a = _lib.AttnDesc()
        if cross:
            a.q, a.k, a.v = qt.data_ptr(), kvt.data_ptr(), kvt.data_ptr() + 2 * Hc
            a.q_ld, a.q_bs, a.kv_ld, a.kv_bs = Hc, Tq * Hc, 2 * Hc, Tk * 2 * Hc
            a.dq, a.dk, a.dv = dqt.data_ptr(), dkvt.data_ptr(), dkvt.data_ptr() + 2 * Hc
            a.dq_ld, a.dq_bs, a.dkv_ld, a.dkv_bs = Hc, Tq * Hc, 2 * Hc, Tk * 2 * Hc
        else:
            a.q, a.k, a.v = qkv_.data_ptr(), qkv_.data_ptr() + 2 * Hc, qkv_.data_ptr() + 4 * Hc
            a.q_ld = a.kv_ld = 3 * Hc; a.q_bs = a.kv_bs = Tq * 3 * Hc
            a.dq, a.dk, a.dv = dqkv_.data_ptr(), dqkv_.data_ptr() + 2 * Hc, dqkv_.data_ptr() + 4 * Hc
            a.dq_ld = a.dkv_ld = 3 * Hc; a.dq_bs = a.dkv_bs = Tq * 3 * Hc
        a.o, a.o_lo, a.o_ld, a.o_bs = o_.data_ptr(), olo_.data_ptr(), Hc, Tq * Hc
        a.stats = stats_.data_ptr(); a.batch, a.heads, a.tq, a.tk, a.head_dim = Bc, nhc, Tq, Tk, 64
        a.scale, a.mask_mode, a.drop, a.seed = 0.125, 0, drop, 3
        a.d_o, a.dsum, a.dq_accum = do_.data_ptr(), dsum_.data_ptr(), dqacc_.data_ptr()
        fl = 4.0 * Bc * nhc * Tq * Tk * 64
        t = _timed(lambda: ctx.check(ctx.lib.ts_attn_fwd(ctx.h, C.byref(a), stream_ptr())), flush)
        out.append({"kernel": f"attn_fwd2_kernel {tag} B{Bc} H{nhc} Tq{Tq} Tk{Tk} hd64 dropout {drop} ({ref})", "bound": "tensor", "achieved": fl / t / 1e12,
                    "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": fl / t / 1e12 / peaks["bf16_tflops"], "us": t * 1e6,
                    "traffic": ncu_traffic(keyf) if keyf else None})
        t = _timed(lambda: ctx.check(ctx.lib.ts_attn_bwd(ctx.h, C.byref(a), stream_ptr())), flush)
        out.append({"kernel": f"attn_bwd_prep + attn_bwd2_kernel (fused dQ/dK/dV) + dq_store {tag} (same shape; 2x forward FLOPs counted, recompute not)",
                    "bound": "tensor", "achieved": 2 * fl / t / 1e12, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": 2 * fl / t / 1e12 / peaks["bf16_tflops"], "us": t * 1e6, "traffic": ncu_traffic(keyb) if keyb else None})

    attn_case("self", B, nh, T, T, 0.1, "V:348-362" if T != 1500 else "W:147-167", "attn_fwd", "attn_bwd")
    try:
        if T != 1500:
            attn_case("whisper encoder self", 4, 12, 1500, 1500, 0.1, "W:147-167")
        attn_case("whisper cross", 4, 12, 100, 1500, 0.1, "W:278-290")
    except Exception as ex:      # the extra shapes never take the line down
        out.append({"kernel": "attention extra shapes", "error": str(ex)[:200]})

    def hbm(name, nbytes, fn, key=None):
        t = _timed(fn, flush)
        gbs = nbytes / t / 1e9
        out.append({"kernel": name, "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                    "us": t * 1e6, "traffic": ncu_traffic(key) if key else None})

    # LayerNorm fwd / bwd on [M, H] bf16
    x = torch.randn(M, H, device=dev).to(bf); y = torch.empty_like(x); dy = torch.randn(M, H, device=dev).to(bf); dx = torch.empty_like(x)
    gam = torch.ones(H, device=dev); bet = torch.zeros(H, device=dev); mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
    dg = torch.zeros(H, device=dev); db = torch.zeros(H, device=dev)
    hbm(f"ln_fwd_kernel [{M},{H}] bf16 (V:411)", 2.0 * M * H * 2,
        lambda: ctx.check(ctx.lib.ts_layernorm_fwd(ctx.h, _lib.TS_BF16, P(x), P(gam), P(bet), P(y), P(mean), P(rstd), M, H, 1e-5, stream_ptr())))
    hbm(f"ln_bwd_ring_kernel (dx + dgamma + dbeta in one pass, cp.async ring) [{M},{H}] bf16", 3.0 * M * H * 2,
        lambda: ctx.check(ctx.lib.ts_layernorm_bwd(ctx.h, _lib.TS_BF16, P(dy), P(x), P(gam), P(mean), P(rstd), None, P(dx), P(dg), P(db), M, H,
                                                   stream_ptr())))
    # GroupNorm + GELU forward on the conv0 output [B, 48000, 512] bf16 (the largest activation of the step)
    T0 = 48000
    xg = torch.randn(B, T0, 512, device=dev).to(bf); yg = torch.empty_like(xg)
    gmean = torch.empty(B, 16, device=dev); grstd = torch.empty(B, 16, device=dev); acc = torch.zeros(2 * B * 16, dtype=torch.float64, device=dev)
    g512 = torch.ones(512, device=dev); b512 = torch.zeros(512, device=dev)
    # algorithmic bytes per SURVEY §8(d): every distinct tensor once in, once out = 2 passes of the activation.
    # (1) as the train step runs it: the moments come from the producer (conv0 store loop / conv GEMM epilogue), x is read once
    ctx.check(ctx.lib.ts_groupnorm_gelu_fwd(ctx.h, _lib.TS_BF16, P(xg), P(g512), P(b512), P(yg), P(gmean), P(grstd), P(acc), B, T0, 512, 16, 1e-5,
                                            stream_ptr()))
    hbm(f"gn_gelu_fwd_kernel [{B},{T0},512] bf16, moments from the conv epilogue (V:140-196, V:248-249; 2 passes of {B * T0 * 512 * 2 / 1e6:.0f} MB)",
        2.0 * B * T0 * 512 * 2,
        lambda: ctx.check(ctx.lib.ts_groupnorm_gelu_fwd(ctx.h, _lib.TS_BF16, P(xg), P(g512), P(b512), P(yg), P(gmean), P(grstd), None, B, T0,
                                                        512, 16, 1e-5, stream_ptr())), key="gn_gelu_fwd")
    # (2) the stand-alone layer (no producer to take the moments): x is read twice, still 2 algorithmic passes
    hbm(f"gn_stats_kernel + gn_gelu_fwd_kernel [{B},{T0},512] bf16, stand-alone GroupNormalization + GELU (x read twice, 2 algorithmic passes)",
        2.0 * B * T0 * 512 * 2,
        lambda: ctx.check(ctx.lib.ts_groupnorm_gelu_fwd(ctx.h, _lib.TS_BF16, P(xg), P(g512), P(b512), P(yg), P(gmean), P(grstd), P(acc), B, T0,
                                                        512, 16, 1e-5, stream_ptr())))
    del xg, yg
    # log-mel front end, B x 30 s
    wav = torch.randn(B, 480000, device=dev)
    nf = ctx.lib.ts_logmel_num_frames(480000)
    mel = torch.empty(B, 80, nf, device=dev)
    hbm(f"logmel_kernel [{B} x 30 s] fp32 (W:739-766)", B * (480000 * 4.0 + 80 * nf * 4.0),
        lambda: ctx.check(ctx.lib.ts_logmel(ctx.h, P(wav), 480000, B, 480000, P(mel), _lib.TS_F32, 1, stream_ptr())), key="logmel")
    return dominant, out


def measure_workload(args, strategy, workload, batch, steps, warmup, rank, world, local):
    """Builds the model of `workload`, captures its train step and times it: `steps` device-resident steps (CUDA events on the
    compute stream, max over ranks) and `steps` end-to-end steps (pinned host -> device copy of every batch + loss read back).
    Returns (result dict, handles dict) — handles keep the model / context alive for the per-kernel rooflines of the main line."""
    import numpy as np
    import torch
    from tethys_speech_b200.runtime import Adam

    family, size, n_samples, secs, gflop = WORKLOADS[workload]
    B = batch or DEFAULT_BATCH[family]
    dev = torch.device("cuda", local)
    rng = np.random.default_rng(1234 + rank)
    npool = 4
    if family == "w2v":
        from tethys_speech_b200 import wav2vec2 as W

        cfg = W.Wav2Vec2Config(size)
        model = W.Wav2Vec2ForPreTraining(cfg, precision=args.precision, device=local, seed=0)
        opt = Adam(learning_rate=3e-5, epsilon=1e-8, clipnorm=1.0)
        host = [(torch.from_numpy(rng.standard_normal((B, n_samples), dtype=np.float32)).pin_memory(), None) for _ in range(npool)]
        h2d = B * n_samples * 4
        step_desc = ("VS:1119-1176 (clip_by_global_norm 1.0 + clipnorm 1.0 + Keras-legacy Adam 3e-5)" if world == 1
                     else "V:1186-1260 (loss/N, local clip, NCCL all-reduce SUM, clipnorm, Adam)")

        def step(batch):
            if world > 1:
                return W.distributed_train_step(strategy, model, batch, opt)
            return W.train_step(model, batch, opt)
        T = model.num_frames(n_samples)
        H, F, nh = cfg.hidden_size, cfg.intermediate_size, cfg.num_attention_heads
        model_name = f"wav2vec2-{size}"
    else:
        from tethys_speech_b200 import whisper as W

        model = W.create_whisper_model(size, precision=args.precision, device=local, seed=0)
        cfg = model.config
        opt = Adam(learning_rate=1e-4)
        labels = np.zeros((B, 100), dtype=np.int32)
        for i in range(B):
            n = int(rng.integers(50, 90))
            labels[i, 0] = 1; labels[i, 1:n - 1] = rng.integers(3, 100, size=n - 2); labels[i, n - 1] = 2
        host = [(torch.from_numpy(rng.standard_normal((B, 80, n_samples), dtype=np.float32)).pin_memory(),
                 torch.from_numpy(labels).pin_memory()) for _ in range(npool)]
        h2d = B * 80 * n_samples * 4 + B * 100 * 4
        step_desc = ("W:823-836 single replica (forward, shifted CE, backward, Keras-legacy Adam 1e-4)" if world == 1
                     else "W:819-848 (un-normalised NCCL all-reduce SUM in buckets overlapped with backward, Adam)")

        def step(batch):
            if world > 1:
                return W.distributed_train_step(strategy, model, batch, opt)
            return W.train_step(model, batch, opt)
        T = n_samples // 2
        H, F, nh = cfg.d_model, cfg.d_ff, cfg.encoder_attention_heads
        model_name = f"whisper-{size}"
    model.broadcast_weights(strategy)
    ctx = model._prog.ctx
    resident = [tuple(t.to(dev) if t is not None else None for t in hb) for hb in host]
    use_graph = not args.no_graph
    graph_launches = None
    if use_graph and world == 1:
        # single GPU: the whole step is one CUDA graph (fresh dropout masks / Adam step from the library's device state)
        from tethys_speech_b200.runtime import GraphedTrainStep

        if family == "w2v":
            def sample_aux():
                return {"neg": model._sample_negative_indices(T, B)[:, 0, :].contiguous()}   # V:907-937, outside the graph
            graphed = GraphedTrainStep(lambda batch, aux: W.train_step(model, batch, opt, neg_indices=aux["neg"]), model, opt,
                                       resident[0], sample_aux())
        else:
            def sample_aux():
                return None
            graphed = GraphedTrainStep(lambda batch, aux: W.train_step(model, batch, opt), model, opt, resident[0], None)
        graph_launches = graphed.launches_per_step

        def step(batch):  # noqa: F811
            return graphed(batch, sample_aux())
    elif use_graph:
        # N > 1: CUDA graphs around the NCCL all-reduces
        if family == "w2v":
            gstep, segs = W.make_graphed_distributed_step(strategy, model, opt, resident[0][0])

            def step(batch):  # noqa: F811
                return gstep(batch[0])
        else:
            gstep, segs = W.make_graphed_distributed_step(strategy, model, opt, resident[0][0], resident[0][1])

            def step(batch):  # noqa: F811
                return gstep(batch[0], batch[1])
        graph_launches = segs.launches_per_step

    def sync_all():
        torch.cuda.synchronize()
        strategy.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------------------------
    # N > 1: the first replays of a graph with captured collectives still pay NCCL's lazy channel / NVLS set-up on 8 ranks (measured:
    # 13.8 ms for the first 13 steps against 12.6 ms steady state, profiles/r02_comm_bench_n8.log); those are taken as part of
    # building the step, before the W warm-up steps the contract asks for
    settle = 8 if world > 1 else 0
    for i in range(settle):
        step(resident[i % npool])
    for i in range(warmup):
        step(resident[i % npool])
    sync_all()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = ctx.lib.ts_launch_count(ctx.h)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = step(resident[i % npool])  # noqa: F841
    e1.record()
    sync_all()
    launches = int(ctx.lib.ts_launch_count(ctx.h) - l0)
    if use_graph:
        launches = graph_launches * steps   # replays do not pass the host-side launch counter
    t_dev = e0.elapsed_time(e1) * 1e-3
    clk = clocks.stop() if rank == 0 else None
    # ---- end-to-end timing: pinned host -> device copy of every batch, loss read back every step ----------
    # A two-deep input pipeline, as a training loop over this API is written (and as the reference's tf.data prefetch does): the
    # pinned-host -> device copy of batch i+1 runs on a copy stream underneath step i, the loss of step i is copied to pinned host
    # memory asynchronously and read (synchronised on its own event) after step i+1 has been queued. Every step's input bytes cross
    # PCIe inside the timed region and every step's loss value is read on the host.
    copy_stream = torch.cuda.Stream(device=dev)
    stagings = [tuple(torch.empty_like(t, device=dev) if t is not None else None for t in host[0]) for _ in range(2)]
    ev_ready = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    ev_loss = [torch.cuda.Event() for _ in range(2)]
    loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]

    def prefetch(i):
        k = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_done[k])          # the step that last consumed this staging buffer has taken its inputs
            for dst, src in zip(stagings[k], host[i % npool]):
                if dst is not None:
                    dst.copy_(src, non_blocking=True)
            ev_ready[k].record(copy_stream)

    def e2e_run(n):
        cur = torch.cuda.current_stream(dev)
        last = None
        prefetch(0)
        for i in range(n):
            k = i % 2
            if i + 1 < n:
                prefetch(i + 1)
            cur.wait_event(ev_ready[k])
            loss_dev = step(stagings[k])
            ev_done[k].record(cur)
            loss_host[k].copy_(loss_dev.detach().reshape(()).float() if hasattr(loss_dev, "detach") else torch.as_tensor(float(loss_dev)), non_blocking=True)
            ev_loss[k].record(cur)
            if i > 0:                                    # read the previous step's loss while this one runs
                ev_loss[1 - k].synchronize()
                last = float(loss_host[1 - k])
        ev_loss[(n - 1) % 2].synchronize()
        return float(loss_host[(n - 1) % 2])

    for ev in ev_done:
        ev.record(torch.cuda.current_stream(dev))
    e2e_run(2)
    sync_all()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    last = e2e_run(steps)
    e3.record()
    sync_all()
    t_e2e = e2.elapsed_time(e3) * 1e-3
    ctx.watchdog()
    # max over ranks
    if world > 1:
        tt = torch.tensor([t_dev, t_e2e], device=dev, dtype=torch.float64)
        strategy.dist.all_reduce(tt, op=strategy.dist.ReduceOp.MAX)
        t_dev, t_e2e = float(tt[0]), float(tt[1])
    peaks = measured_peaks()
    sps = B * world * steps / t_dev
    sps_e2e = B * world * steps / t_e2e
    step_tflops = sps / world * gflop / 1e3
    res = {
        "metric": "train_samples_per_sec", "value": sps, "unit": "samples/s", "audio_sec_per_sec": sps * secs,
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": t_dev / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": workload_config(workload, B, world),
        "run": {"cuda_graph": bool(use_graph), "settle_steps_before_warmup": settle,
                "collectives": ("none (1 replica)" if world == 1 else strategy.allreduce_description(model._prog))},
        "clocks": clk,
        "e2e": {"value": sps_e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": t_e2e / steps * 1e3, "last_loss": last},
        "gpu_launches": launches,
        "launches_per_step": launches / steps,
        "simt_downgrades": int(ctx.lib.ts_simt_downgrades(ctx.h)),
        "step_tflops_per_gpu": step_tflops,
        "step_frac_of_bf16_sustained": step_tflops / peaks["bf16_tflops_sustained"],
    }
    del resident, stagings
    handles = {"model": model, "opt": opt, "ctx": ctx, "B": B, "T": T if family == "w2v" else 1500, "H": H, "F": F, "nh": nh,
               "n_params": int(model._prog.n), "family": family}
    return res, handles


# workloads measured besides the main line so that every BASELINE.json config is driver-visible (each: its own model, a few
# steps, < 1 s of GPU time): N = 1 -> configs[0] (Whisper default preset, B = 4) and Whisper-base; N > 1 -> configs[2]
# (Whisper-base, bf16, data parallel) and configs[3] (Wav2Vec2-large, 8 per GPU = global batch 64 at N = 8)
EXTRA_WORKLOADS = {1: ("whisper_small_30s", "whisper_base_30s"), 0: ("whisper_base_30s", "w2v_large_15s", "whisper_small_30s")}


def run_ours(args):
    import gc

    import torch
    from tethys_speech_b200.runtime import Strategy

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (our arm) needs a B200; there is no CPU fallback"
    torch.cuda.set_device(local)
    strategy = Strategy()
    line, hd = measure_workload(args, strategy, args.workload, args.batch, args.steps, args.warmup, rank, world, local)
    peaks = measured_peaks()
    if rank == 0:
        torch.cuda.empty_cache()
        kernels = []
        try:
            dominant, kernels = kernel_rooflines(hd["ctx"], peaks, hd["B"], hd["T"], hd["H"], hd["F"], hd["nh"], hd["n_params"])
            line["roofline"] = {"bound": "tensor", "kernel": dominant["kernel"], "achieved": dominant["achieved"], "peak": peaks["bf16_tflops"],
                                "unit": "TFLOP/s", "frac": dominant["frac"], "traffic": dominant["traffic"],
                                "peak_source": peaks["source"] + " (burst: kernel timed alone)"}
        except Exception as ex:  # noqa: BLE001 — the step measurement above is already complete: keep the line, say what failed
            line["roofline"] = {"bound": "tensor", "achieved": None, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": None,
                                "traffic": None, "error": f"{type(ex).__name__}: {ex}"[:300]}
        try:      # BASELINE config 5 (front-end microbench) at this workload's shape; never takes the line down
            kernels += frontend_rooflines(peaks, local)
        except Exception as ex:  # noqa: BLE001
            kernels.append({"kernel": "conv feature encoder forward as a whole", "error": f"{type(ex).__name__}: {ex}"[:200]})
        line["kernel_rooflines"] = kernels
    hd.clear()
    gc.collect()
    torch.cuda.empty_cache()
    extras = []
    if not args.no_extra:
        for wl in EXTRA_WORKLOADS[1 if world == 1 else 0]:
            if wl == args.workload:
                continue
            try:
                r, h2 = measure_workload(args, strategy, wl, 0, min(args.steps, 10), max(3, min(args.warmup, 5)), rank, world, local)
                extras.append({k: r[k] for k in ("value", "unit", "audio_sec_per_sec", "ms_per_step", "n_gpus", "steps", "warmup", "dtype",
                                                 "config", "run", "clocks", "e2e", "gpu_launches", "launches_per_step", "step_tflops_per_gpu",
                                                 "step_frac_of_bf16_sustained")})
                h2.clear()
            except Exception as ex:   # an extra workload never takes the main line down
                extras.append({"config": {"workload": wl}, "error": f"{type(ex).__name__}: {ex}"[:300]})
            gc.collect()
            torch.cuda.empty_cache()
    if rank != 0:
        if world > 1:
            strategy.dist.destroy_process_group()
        return 0
    line["extra"] = {"workloads": extras}
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_step_throughput(args.workload, 10, 1, max_seconds=40.0)   # ~10 s of CPU work (10 steps of batch 1), capped at 40 s
        line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line), flush=True)
    if world > 1:
        strategy.dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="w2v_base_15s", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (--batch_size of the reference CLI); default 8 (w2v) / 4 (whisper)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--no-graph", dest="no_graph", action="store_true", help="launch the step kernel by kernel instead of replaying a CUDA graph")
    ap.add_argument("--no-extra", dest="no_extra", action="store_true", help="skip the extra BASELINE workloads reported under extra.workloads")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
