#!/usr/bin/env python
"""GEMM shapes of one configs[1] step, timed per tile choice (tools/bench_gemm.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asr_b200 import ops, _lib
dev = "cuda"
R = 501 * 64
shapes = [("in-proj fwd l0", R, 4800, 1312), ("in-proj fwd", R, 4800, 800), ("dgrad l1+", R, 800, 4800), ("wgrad l0", 4800, 1312, R),
          ("wgrad", 4800, 800, R), ("dW_hh", 2400, 800, R), ("fc", R, 29, 800)]
def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, M, N, K in shapes:
    Kp = (K + 3) // 4 * 4
    A = torch.randn(M, Kp, device=dev); B = torch.randn(N, Kp, device=dev); C = torch.empty(M, N, device=dev)
    out = []
    for bn in (128, 256):
        _lib.query("asrb_debug_gemm_tile", bn, 0)
        t = timed(lambda: ops._call("asrb_gemm_tn", ops._p(A), Kp, ops._p(B), Kp, ops._p(C), N, None, M, N, K, 0))
        out.append(f"BN={bn}: {t:.3f} ms {2.0*M*N*K/t/1e9:.0f} TF/s")
    _lib.query("asrb_debug_gemm_tile", 0, 0)
    ref = A[:, :K] @ B[:, :K].t()
    ops._call("asrb_gemm_tn", ops._p(A), Kp, ops._p(B), Kp, ops._p(C), N, None, M, N, K, 0)
    err = ((C - ref).abs().max() / ref.abs().max()).item()
    print(f"{name:16s} M={M} N={N} K={K}: " + " | ".join(out) + f" | rel err {err:.1e}")
