"""Which co-running kernel slows the recurrent backward chain?  (configs[1] layer shape)

The weight-gradient work of layer l+1 runs on a second stream under layer l's recurrent backward kernel
(asr_b200/functional.py); in the step the chain then takes ~8 % longer.  This times ONE recurrent backward launch alone
and with each kind of side-stream kernel looping next to it -- bf16 weight-gradient GEMMs at several CTA caps, bf16
transposes, bias row sums -- so that the next change (throttle, reorder, or move a kernel off the side stream) is
picked from a measurement.  Usage: python tools/bench_overlap.py   (one GPU, ~20 s)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from asr_b200 import ops

T, B, H = 501, 64, 800
cell, G, R = ops.GRU, 3 * 800, 501 * 64
dev = "cuda"
torch.manual_seed(0)
gi = torch.randn(T, B, 2, G, device=dev)
b_hh = torch.randn(2, G, device=dev) * 0.03
w = [(torch.rand(G, H, device=dev) * 2 - 1) * 0.035 for _ in range(2)]
lens = torch.full((B,), T, dtype=torch.int32, device=dev)
pf, pb = ops.rnn_pack_weights(cell, w[0], w[1], B)
hseq, cseq, saved = ops.rnn_fwd(cell, gi, b_hh, pf, lens, T, B, H)
dout = torch.randn(T, B, H, device=dev)
dgi, dgiT, dghT = ops.rnn_bwd(cell, dout, pb, lens, hseq, cseq, saved, T, B, H)
x2 = torch.randn(R, H, device=dev)
xt = ops.transpose_bf16(x2)
side = torch.cuda.Stream()


def chain_ms(co_runner=None, reps=3):
    """mean duration of one recurrent backward launch; co_runner() is issued on the side stream, enough times to
    cover the whole launch, right before each launch"""
    out = []
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if co_runner is not None:
            with torch.cuda.stream(side):
                co_runner()
        e0.record()
        ops.rnn_bwd(cell, dout, pb, lens, hseq, cseq, saved, T, B, H)
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return sum(out[1:]) / reps


def gemms(cap, n):
    def run():
        old = ops.gemm_cta_limit(cap)
        try:
            for _ in range(n):
                ops.gemm_tn_bf16(dgiT[:G, :R], xt)
        finally:
            ops.gemm_cta_limit(old)
    return run


def alone_ms(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


base = chain_ms()
print(f"recurrent backward alone: {base:.3f} ms ({base * 1e3 / T:.2f} us/step)")
for cap in (0, 64, 48, 32, 16):
    t1 = alone_ms(gemms(cap, 1))
    n = max(1, int(base * 1.2 / t1) + 1)
    t = chain_ms(gemms(cap, n))
    print(f"  + wgrad GEMMs, cap {cap:3d} ({t1:.3f} ms each alone, {n} queued): chain {t:.3f} ms ({100 * (t / base - 1):+.1f} %)")
for name, fn in (("transpose_bf16 [32064x800]", lambda: ops.transpose_bf16(x2)),
                 ("row_sums bf16 [4800x32064]", lambda: ops.row_sums(dgiT, R))):
    t1 = alone_ms(fn)
    n = max(1, int(base * 1.2 / t1) + 1)
    t = chain_ms(lambda: [fn() for _ in range(n)])
    print(f"  + {name} ({t1:.3f} ms each alone, {n} queued): chain {t:.3f} ms ({100 * (t / base - 1):+.1f} %)")
