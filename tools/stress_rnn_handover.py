"""Race detector for the recurrent kernels' step hand-over: every buffer the kernels allocate is pre-poisoned with bf16 / fp32
NaN bit patterns (the caching allocator hands the same blocks back to torch.empty), so a consumer that reads an operand tile
before its producer's data is in L2 turns the outputs NaN.  Usage: stress_rnn_handover.py <dbg bits> <launch pairs> [cell]"""
import sys
import torch
sys.path.insert(0, ".")
from asr_b200 import _lib, ops

dbg = int(sys.argv[1]) if len(sys.argv) > 1 else 0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
cellname = sys.argv[3] if len(sys.argv) > 3 else "gru"
dev = torch.device("cuda")
T, B, H = 501, 64, 800
cell = ops.GRU if cellname == "gru" else ops.LSTM
G = (3 if cellname == "gru" else 4) * H
g = torch.Generator().manual_seed(5)
k = H ** -0.5
gi = torch.randn(T, B, 2, G, generator=g).to(dev)
w = ((torch.rand(2, G, H, generator=g) * 2 - 1) * k).to(dev)
b_hh = ((torch.rand(2, G, generator=g) * 2 - 1) * k).to(dev)
dout = torch.randn(T, B, H, generator=g).to(dev)
lens = torch.full((B,), T, dtype=torch.int32, device=dev)
ops.RNN_BF16 = True
_lib.query("asrb_debug_rnn_dbg", dbg)
pf, pb = ops.rnn_pack_weights(cell, w[0].contiguous(), w[1].contiguous(), B)


def poison():
    # ~3 GB of 0xFFFFFFFF (bf16 NaN pairs / fp32 NaN), freed again: the next torch.empty calls get these blocks
    junk = [torch.full((64 * 1024 * 1024,), -1, dtype=torch.int32, device=dev) for _ in range(12)]
    del junk


ref = None
bad = 0
for r in range(reps):
    poison()
    hs, cs, sv = ops.rnn_fwd(cell, gi, b_hh, pf, lens, T, B, H)
    dgi, dgiT, dghT = ops.rnn_bwd(cell, dout, pb, lens, hs, cs, sv, T, B, H)
    torch.cuda.synchronize()
    ok_f = bool(torch.isfinite(hs[:, 1:T + 1]).all())
    ok_b = bool(torch.isfinite(dgi.float()).all())
    cur = (hs[:, 1:T + 1].clone(), dgi.clone())
    same = ref is None or (torch.equal(cur[0], ref[0]) and torch.equal(cur[1], ref[1]))
    if ref is None and ok_f and ok_b:
        ref = cur
    if not (ok_f and ok_b and same):
        bad += 1
        print(f"launch pair {r}: forward finite {ok_f}, backward finite {ok_b}, identical to the first run {same}", flush=True)
    del hs, cs, sv, dgi, dgiT, dghT, cur
print(f"dbg={dbg} {cellname}: {bad} bad of {reps} launch pairs ({reps * (T - 1) * 2} hand-overs per direction and chain)", flush=True)
