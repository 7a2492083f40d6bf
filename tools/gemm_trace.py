#!/usr/bin/env python
"""Where does a tcgen05 GEMM launch spend its time? Runs one shape through ts_gemm with the in-kernel %globaltimer stamps on
(ts_debug_gemm_trace) and prints, over all CTAs: head (entry -> first MMA can issue), main loop, tail (last MMA issued -> stores
drained), and the spread of CTA start / end times. usage: python tools/gemm_trace.py m n k [a_major b_major] [ctas] [cold]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tethys_speech_b200 import _lib  # noqa: E402
from tethys_speech_b200._lib import Context  # noqa: E402
from tethys_speech_b200.runtime import stream_ptr  # noqa: E402


def main():
    m, n, k = (int(x) for x in sys.argv[1:4])
    amaj, bmaj = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (0, 1)
    ctas = int(sys.argv[6]) if len(sys.argv) > 6 else 1
    cold = len(sys.argv) > 7 and sys.argv[7] == "cold"
    dev = torch.device("cuda", 0)
    ctx = Context(0)
    bf = torch.bfloat16
    a = (torch.randn((m, k) if amaj == 0 else (k, m), device=dev) * 0.5).to(bf)
    b = (torch.randn((n, k) if bmaj == 0 else (k, n), device=dev) * 0.05).to(bf)
    c = torch.zeros(m, n, device=dev, dtype=bf)
    d = _lib.GemmDesc()
    d.a, d.b, d.c = a.data_ptr(), b.data_ptr(), c.data_ptr()
    d.m, d.n, d.k, d.a_major, d.b_major = m, n, k, amaj, bmaj
    d.lda, d.ldb, d.ldc = a.shape[1], b.shape[1], n
    d.batch1 = d.batch2 = 1
    d.in_dtype, d.out_dtype, d.alpha = _lib.TS_BF16, _lib.TS_BF16, 1.0
    d.force_engine = 3 if ctas == 2 else 2
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    trace = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    for _ in range(3):
        ctx.check(ctx.lib.ts_gemm(ctx.h, C.byref(d), stream_ptr()))
    ctx.check(ctx.lib.ts_debug_gemm_trace(ctx.h, C.c_void_p(trace.data_ptr())))
    if cold:
        flush.fill_(1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ctx.check(ctx.lib.ts_gemm(ctx.h, C.byref(d), stream_ptr()))
    e1.record()
    torch.cuda.synchronize()
    ctx.check(ctx.lib.ts_debug_gemm_trace(ctx.h, C.c_void_p(0)))
    t = trace.cpu().view(148, 16).double()
    live = t[:, 0] > 0
    t = t[live]
    t0 = t[:, 0].min()
    us = lambda x: float(x) / 1e3
    print(f"gemm {m}x{n}x{k} maj={amaj}{bmaj} ctas={ctas} {'cold' if cold else 'warm'}: event time {e0.elapsed_time(e1) * 1e3:.1f} us, "
          f"{int(live.sum())} CTAs, kernel span (first entry -> last drain) {us(t[:, 7].max() - t0):.1f} us")
    lead = t[t[:, 4] > 0]        # CTAs that issue MMAs (all of them for 1-CTA tiles, the leaders for pairs)
    def col(name, v):
        print(f"  {name:44s} min {us(v.min()):7.2f}  median {us(v.median()):7.2f}  max {us(v.max()):7.2f} us")
    col("CTA entry, after the first CTA's entry", t[:, 0] - t0)
    col("setup (barriers, TMEM alloc, sync)", t[:, 1] - t[:, 0])
    col("setup done -> first operands landed", lead[:, 2] - lead[:, 1])
    col("first tile: operands landed -> all MMAs issued", lead[:, 3] - lead[:, 2])
    col("main loop: first operands -> last MMA issued", lead[:, 4] - lead[:, 2])
    col("last MMA issued -> last accumulator ready", t[:, 6] - lead[:, 4].median())
    col("last accumulator ready -> stores drained", t[:, 7] - t[:, 6])
    col("  ... -> first chunk in registers", t[:, 8] - t[:, 6])
    col("  ... -> first chunk staged in smem", t[:, 9] - t[:, 6])
    col("  ... -> last chunk staged in smem", t[:, 10] - t[:, 6])
    col("CTA exit, before the last CTA's", t[:, 7].max() - t[:, 7])


if __name__ == "__main__":
    main()
