"""DEBUG: per-phase timing of the persistent recurrent kernel (SM clock stamps), cfg2 shape."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asr_b200 import ops, _lib
T, B, H = int(sys.argv[1]) if len(sys.argv) > 1 else 501, 64, 800
cell = ops.GRU; G = 3 * H
dev = "cuda"
torch.manual_seed(0)
gi = torch.randn(T, B, 2, G, device=dev); b_hh = torch.randn(2, G, device=dev) * 0.03
w = [(torch.rand(G, H, device=dev) * 2 - 1) * 0.035 for _ in range(2)]
lens = torch.full((B,), T, dtype=torch.int32, device=dev)
pf, pb = ops.rnn_pack_weights(cell, w[0], w[1], B)
nj, P, _, _ = ops.rnn_plan(cell, H, B, ops.rnn_use_bf16(H))
grid = 2 * P
names = ["P:counter ok", "P:loads issued", "M:first full", "M:commit", "E:step top", "E:tfull", "E:ld done", "E:math+stores", "E:bar done", "P:fence done", "E:red", "E:exchange done", "E:staged+bar", "E:bulk issued"]
def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for flags in (0, 2, 4):
    _lib.query("asrb_debug_rnn_ksplit", flags)
    pf, pb = ops.rnn_pack_weights(cell, w[0], w[1], B)
    hseq, cseq, saved = ops.rnn_fwd(cell, gi, b_hh, pf, lens, T, B, H)
    dout = torch.randn(T, B, H, device=dev)
    tf = timed(lambda: ops.rnn_fwd(cell, gi, b_hh, pf, lens, T, B, H))
    tb = timed(lambda: ops.rnn_bwd(cell, dout, pb, lens, hseq, cseq, saved, T, B, H))
    print(f"ksplit={flags} (no trace): fwd {tf:.3f} ms ({tf*1e3/T:.2f} us/step)  bwd {tb:.3f} ms ({tb*1e3/T:.2f} us/step)")
_lib.query("asrb_debug_rnn_ksplit", int(sys.argv[2]) if len(sys.argv) > 2 else 2)
pf, pb = ops.rnn_pack_weights(cell, w[0], w[1], B)
for which in ("fwd", "bwd"):
    for _ in range(2):
        hseq, cseq, saved = ops.rnn_fwd(cell, gi, b_hh, pf, lens, T, B, H)
    trace = torch.zeros(grid + 8, T, 16, dtype=torch.int64, device=dev)
    _lib.call("asrb_debug_rnn_trace", ctypes.c_void_p(trace.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if which == "fwd":
        ops.rnn_fwd(cell, gi, b_hh, pf, lens, T, B, H)
    else:
        dout = torch.randn(T, B, H, device=dev)
        ops.rnn_bwd(cell, dout, pb, lens, hseq, cseq, saved, T, B, H)
    e1.record(); torch.cuda.synchronize()
    _lib.call("asrb_debug_rnn_trace", None)
    tr = trace.cpu().double()
    ms = e0.elapsed_time(e1)
    print(f"== {which}: {ms:.3f} ms total, {ms*1e3/T:.2f} us/step, grid {grid}")
    s0, s1 = 50, T - 50
    for cta in (0, 1, 2, 3, P // 2):
        x = tr[cta, s0:s1]
        top = x[:, 4]
        step_cycles = (top[1:] - top[:-1]).mean().item()
        rel = [(x[:, k] - top).mean().item() for k in range(14)]
        print(f" cta {cta}: cycles/step {step_cycles:.0f}; " + "; ".join(f"{n}={r:.0f}" for n, r in zip(names, rel)))
