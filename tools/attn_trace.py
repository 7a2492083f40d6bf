#!/usr/bin/env python
"""Per-tile timeline of the fused attention forward's two softmax groups (CTA 0, clock64 stamps written when a debug buffer is
registered with ts_debug_gemm_trace). usage: python tools/attn_trace.py [B H T drop]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tethys_speech_b200 import _lib  # noqa: E402
from tethys_speech_b200._lib import Context  # noqa: E402
from tethys_speech_b200.runtime import stream_ptr  # noqa: E402


def main():
    B, nh, T = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (8, 12, 750)
    drop = float(sys.argv[4]) if len(sys.argv) > 4 else 0.1
    H = nh * 64
    dev = torch.device("cuda", 0)
    ctx = Context(0)
    bf = torch.bfloat16
    qkv = torch.randn(B, T, 3 * H, device=dev).to(bf)
    o = torch.empty(B, T, H, device=dev, dtype=bf); olo = torch.empty_like(o)
    stats = torch.empty(B, nh, T, 2, device=dev)
    a = _lib.AttnDesc()
    a.q, a.k, a.v, a.o, a.o_lo = qkv.data_ptr(), qkv.data_ptr() + 2 * H, qkv.data_ptr() + 4 * H, o.data_ptr(), olo.data_ptr()
    a.q_ld = a.kv_ld = 3 * H; a.q_bs = a.kv_bs = T * 3 * H; a.o_ld = H; a.o_bs = T * H
    a.stats = stats.data_ptr(); a.batch, a.heads, a.tq, a.tk, a.head_dim = B, nh, T, T, 64
    a.scale, a.mask_mode, a.drop, a.seed = 0.125, 0, drop, 3
    if len(sys.argv) > 5 and sys.argv[5] == "bwd":
        return trace_bwd(ctx, a, B, T, H, dev)
    trace = torch.zeros(2 * 16 * 10 + 2 * 4 * 8 + 64, dtype=torch.int64, device=dev)
    for _ in range(3):
        ctx.check(ctx.lib.ts_attn_fwd(ctx.h, C.byref(a), stream_ptr()))
    ctx.check(ctx.lib.ts_debug_gemm_trace(ctx.h, C.c_void_p(trace.data_ptr())))
    ctx.check(ctx.lib.ts_attn_fwd(ctx.h, C.byref(a), stream_ptr()))
    torch.cuda.synchronize()
    ctx.check(ctx.lib.ts_debug_gemm_trace(ctx.h, C.c_void_p(0)))
    te = trace.cpu()[320:384].view(2, 4, 8)
    tm = trace.cpu()[384:].view(16, 4)
    t = trace.cpu()[:320].view(2, 16, 10)
    t0 = int(t[0, 0, 0])
    names = ["loop top", "S ready", "token", "S in regs", "max done", "exp+sum", "dropout", "P in TMEM", "arrived"]
    print(f"attention fwd B={B} H={nh} T={T} drop={drop}: cycles since group 0's first loop top (CTA 0, warps 2 / 6)")
    print("grp tile  " + "".join(f"{n:>11s}" for n in names))
    for j in range(16):
        for g in range(2):
            if int(t[g, j, 0]) == 0:
                continue
            print(f" {g}   {j:2d}   " + "".join(f"{int(t[g, j, k]) - t0:11d}" for k in range(9)))
    print("MMA thread, group 0: tile   P0 seen   P0.V issued   next S0 issued")
    for j in range(16):
        if int(tm[j, 0]):
            print(f"                      {j:2d} " + "".join(f"{int(tm[j, k]) - t0:11d}" for k in range(3)))
    print("epilogue: grp item   start  last PV done  O[0:32] in regs  stored  O[32:64] in regs  stored  end")
    for n in range(4):
        for g in range(2):
            if int(te[g, n, 0]) == 0:
                continue
            print(f"   {g}   {n}  " + "".join(f"{int(te[g, n, k]) - t0:11d}" for k in range(7)))


def trace_bwd(ctx, a, B, T, H, dev):
    """Timeline of the fused backward's element-wise warp 2 and MMA warp (CTA 0)."""
    bf = torch.bfloat16
    do = torch.randn(B, T, H, device=dev).to(bf)
    dqkv = torch.empty(B, T, 3 * H, device=dev, dtype=bf)
    dsum = torch.empty(B, a.heads, T, device=dev)
    dqacc = torch.empty(B, T, H, device=dev)
    a.d_o, a.dq, a.dk, a.dv = do.data_ptr(), dqkv.data_ptr(), dqkv.data_ptr() + 2 * H, dqkv.data_ptr() + 4 * H
    a.dq_ld = a.dkv_ld = 3 * H; a.dq_bs = a.dkv_bs = T * 3 * H; a.dsum = dsum.data_ptr(); a.dq_accum = dqacc.data_ptr()
    ctx.check(ctx.lib.ts_attn_fwd(ctx.h, C.byref(a), stream_ptr()))
    for _ in range(3):
        ctx.check(ctx.lib.ts_attn_bwd(ctx.h, C.byref(a), stream_ptr()))
    trace = torch.zeros(256, dtype=torch.int64, device=dev)
    ctx.check(ctx.lib.ts_debug_gemm_trace(ctx.h, C.c_void_p(trace.data_ptr())))
    ctx.check(ctx.lib.ts_attn_bwd(ctx.h, C.byref(a), stream_ptr()))
    torch.cuda.synchronize()
    ctx.check(ctx.lib.ts_debug_gemm_trace(ctx.h, C.c_void_p(0)))
    t = trace.cpu()
    ew, mm = t[:128].view(16, 8), t[128:].view(16, 8)
    t0 = int(ew[0, 0])
    print("element-wise warp 2: tile  loop top  stats synced  S/dP ready  in regs+released  computed  PZ/dS free  stored+arrived")
    for j in range(16):
        if int(ew[j, 0]):
            print(f"                     {j:3d} " + "".join(f"{int(ew[j, k]) - t0:11d}" for k in range(7)))
    print("MMA warp:            tile  loop top  S/dP released  next S/dP issued  PZ/dS ready  dV/dK/dQ issued")
    for j in range(16):
        if int(mm[j, 0]):
            print(f"                     {j:3d} " + "".join(f"{int(mm[j, k]) - t0:11d}" for k in range(5)))


def _unused():
    pass


if __name__ == "__main__":
    main()
    # appended: epilogue stamps
