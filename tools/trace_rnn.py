"""DEBUG: per-phase timing of the persistent recurrent kernels (SM clock stamps), cfg2 shape by default.
usage: trace_rnn.py [T] [ksplit] [cell: gru|lstm] [H] [B]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asr_b200 import ops, _lib
T = int(sys.argv[1]) if len(sys.argv) > 1 else 501
ks = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cellname = sys.argv[3] if len(sys.argv) > 3 else "gru"
H = int(sys.argv[4]) if len(sys.argv) > 4 else 800
B = int(sys.argv[5]) if len(sys.argv) > 5 else 64
cell = ops.GRU if cellname == "gru" else ops.LSTM
G = (3 if cell == ops.GRU else 4) * H
dev = "cuda"
torch.manual_seed(0)
gi = torch.randn(T, B, 2, G, device=dev); b_hh = torch.randn(2, G, device=dev) * 0.03
w = [(torch.rand(G, H, device=dev) * 2 - 1) * 0.035 for _ in range(2)]
lens = torch.full((B,), T, dtype=torch.int32, device=dev)
dout = torch.randn(T, B, H, device=dev)
# slot names of the exchange-by-data kernel (rnn2.cu)
names2 = {4: "E:step top", 0: "P:canaries ok", 10: "E:go", 8: "E:deferred stores issued", 1: "P:copies issued", 2: "M:first K block",
          3: "M:last commit", 5: "E:tfull", 6: "E:ld+verdict", 11: "E:exchange done", 7: "E:operand stored"}
names1 = {4: "E:step top", 14: "chain1 step top", 12: "chain1 counter ok", 13: "chain1 commit", 0: "P:counter ok", 9: "P:fence done", 2: "M:first full", 1: "P:loads issued", 3: "M:commit",
          5: "E:tfull", 6: "E:ld done", 11: "E:exchange done", 7: "E:math+stores", 8: "E:bar done", 10: "E:red"}


def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


# outputs of every variant against the counter + TMA kernels of rnn.cu
ref = None
for dbg, label in ((8, "rnn.cu"), (0, "rnn3 forward + backward (TMEM weights, two chains)"), (1024, "rnn3 forward, rnn.cu backward")):
    _lib.query("asrb_debug_rnn_dbg", dbg)
    pf, pb = ops.rnn_pack_weights(cell, w[0], w[1], B)
    hs, cs, sv = ops.rnn_fwd(cell, gi, b_hh, pf, lens, T, B, H)
    dgi, dgiT, dghT = ops.rnn_bwd(cell, dout, pb, lens, hs, cs, sv, T, B, H)
    torch.cuda.synchronize()
    if ref is None:
        ref = (hs.clone(), dgi.float().clone(), dgiT.float().clone(), None if dghT is None else dghT.float().clone())
    else:
        msg = f"## {label} vs rnn.cu: hseq max abs diff {(hs - ref[0]).abs().max().item():.3e} (|h| max {ref[0].abs().max().item():.3f}); " \
              f"dgi {(dgi.float() - ref[1]).abs().max().item():.3e} (max {ref[1].abs().max().item():.3f}); dgiT {(dgiT.float() - ref[2]).abs().max().item():.3e}"
        if dghT is not None:
            msg += f"; dghT {(dghT.float() - ref[3]).abs().max().item():.3e}"
        print(msg, flush=True)

variants = [(0, "default: rnn3.cu forward and backward (weights in tensor memory, two chains of 32 rows), VERIFIED hand-over (TMA-store first pass with sentinel check + conditional release second pass)"),
            (8192, "rnn3, verified hand-over with a withheld tile: both passes run"),
            (4096, "rnn3, hand-over = generic stores + red.release, one pass"),
            (16, "rnn3, hand-over = TMA store + completion + relaxed increment, unverified (UNSAFE: 1e-7 race)"),
            (8, "counter + TMA (rnn.cu)")]
for dbg, label in variants:
    _lib.query("asrb_debug_rnn_dbg", dbg)
    _lib.query("asrb_debug_rnn_ksplit", ks)
    pf, pb = ops.rnn_pack_weights(cell, w[0], w[1], B)
    nj, P, _, _ = ops.rnn_plan(cell, H, B, ops.rnn_use_bf16(H))
    hseq, cseq, saved = ops.rnn_fwd(cell, gi, b_hh, pf, lens, T, B, H)
    tf = timed(lambda: ops.rnn_fwd(cell, gi, b_hh, pf, lens, T, B, H))
    tb = timed(lambda: ops.rnn_bwd(cell, dout, pb, lens, hseq, cseq, saved, T, B, H))
    print(f"## {label}: {cellname} H={H} B={B} T={T} nj={nj} P={P} ksplit={ks}: fwd {tf:.3f} ms ({tf*1e3/T:.2f} us/step)  "
          f"bwd {tb:.3f} ms ({tb*1e3/T:.2f} us/step)", flush=True)
    names = names2 if dbg & 256 else names1
    for which in ("fwd", "bwd"):
        grid = 2 * (P + 3)
        trace = torch.zeros(grid + 8, T, 16, dtype=torch.int64, device=dev)
        _lib.call("asrb_debug_rnn_trace", ctypes.c_void_p(trace.data_ptr()))
        if which == "fwd":
            ops.rnn_fwd(cell, gi, b_hh, pf, lens, T, B, H)
        else:
            ops.rnn_bwd(cell, dout, pb, lens, hseq, cseq, saved, T, B, H)
        torch.cuda.synchronize()
        _lib.call("asrb_debug_rnn_trace", None)
        tr = trace.cpu().double()
        s0, s1 = min(50, T // 4), max(T - 50, T // 2)
        tot = tr[0, T - 1, 4] - tr[0, 0, 4]
        print(f" {which}: first to last step top {tot:.0f} cycles = {tot / 1.965e6:.3f} ms @1965 MHz", flush=True)
        if which == "fwd" and not dbg & (8 | 256):
            pro = tr[:2 * P, 0, 4] - tr[:2 * P, 0, 15]
            ns = (tr[0, T - 1, 11] - tr[0, 0, 11]).item()
            cyc = (tr[0, T - 1, 4] - tr[0, 0, 4]).item()
            print(f" fwd: first to last step top: {cyc:.0f} SM cycles in {ns / 1e3:.1f} us of wall clock = {cyc / max(ns, 1) * 1e3:.0f} MHz", flush=True)
            print(f" fwd: kernel entry -> first step top, per CTA: min {pro.min().item():.0f}  median {pro.median().item():.0f}  max {pro.max().item():.0f} cycles", flush=True)
        for cta in (0, P):
            x = tr[cta, s0:s1]
            top = x[:, 4]
            step_cycles = (top[1:] - top[:-1]).mean().item()
            rel = {k: (x[:, k] - top).mean().item() for k in names}
            extra = f"; steps repeated by the CTA over {T} steps: {tr[cta, T - 1, 9].item():.0f}" if dbg & 256 else ""
            if which == "fwd" and not dbg & (8 | 256) and x[:, 14].abs().sum() > 0:
                c1 = x[:, 14]
                print(f" {which} cta {cta}: chain 1 cycles/step {(c1[1:] - c1[:-1]).mean().item():.0f}; whole run: chain 0 {(tr[cta, T - 1, 4] - tr[cta, 0, 4]).item():.0f} cycles, "
                      f"chain 1 {(tr[cta, T - 1, 14] - tr[cta, 0, 14]).item():.0f}; chain 1 ends {(tr[cta, T - 1, 14] - tr[cta, T - 1, 4]).item():.0f} cycles after chain 0", flush=True)
            print(f" {which} cta {cta}: cycles/step {step_cycles:.0f}; " +
                  "; ".join(f"{names[k]}={rel[k]:.0f}" for k in names if abs(rel[k]) < 1e6) + extra, flush=True)
        if not dbg & 256:
            # per CTA: cycles from its own release (slot 10) to its control thread seeing the next step's counter (slot 0 of s+1).
            # The CTA that publishes last waits only for the counter's round trip; the others also wait for it (skew).
            ncta = 2 * P
            wt = (tr[:ncta, s0 + 1:s1, 0] - tr[:ncta, s0:s1 - 1, 10]).mean(dim=1)
            own = (tr[:ncta, s0:s1 - 1, 10] - tr[:ncta, s0:s1 - 1, 0]).mean(dim=1)      # counter seen -> own release
            print(f" {which}: release -> next counter seen, per CTA: min {wt.min().item():.0f}  median {wt.median().item():.0f}  max {wt.max().item():.0f};"
                  f"  counter seen -> own release: min {own.min().item():.0f}  median {own.median().item():.0f}  max {own.max().item():.0f}", flush=True)
_lib.query("asrb_debug_rnn_dbg", 0)
