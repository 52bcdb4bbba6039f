"""Repeat the configs[1] fwd+bwd on ragged batches (full batch and its halves) and count non-finite gradients, under the
two hand-over protocols of rnn3.cu (argument: dbg bits, repetitions)."""
import sys
import torch
sys.path.insert(0, ".")
import bench
from asr_b200 import _lib
from tests.test_gpu_fullsize import _ragged_batch, _step, _halves

dbg = int(sys.argv[1]) if len(sys.argv) > 1 else 0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
_lib.query("asrb_debug_rnn_dbg", dbg)
cfg = dict(bench.CFG)
model = bench.build_model(cfg, torch.device("cuda")).train()
for m in model.modules():
    if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
        m.eval()
def poison():
    # fill ~3 GB with 0xFFFFFFFF (bf16 / fp32 NaN) and free it: what torch.empty returns next is NaN, so a read of a buffer
    # before its producer wrote it shows up as a non-finite gradient
    junk = [torch.full((64 * 1024 * 1024,), -1, dtype=torch.int32, device="cuda") for _ in range(12)]
    # ... and the small-block pool (allocations below 1 MB live in their own 2 MB segments)
    for n in (128, 1024, 8 * 1024, 64 * 1024, 200 * 1024):
        junk += [torch.full((n,), -1, dtype=torch.int32, device="cuda") for _ in range(400)]
    del junk


bad = 0
redos0 = _lib.query("asrb_debug_rnn_redos")
for r in range(reps):
    if r % 8 == 0:
        poison()
    batch, lens = _ragged_batch(cfg["B"], cfg["T"], 60, cfg["C"], seed=77 + r)
    for hb in [batch] + _halves(batch, lens, cfg["T"]):
        poison()
        loss, grads = _step(model, hb)
        nf = [k for k, g in grads.items() if not torch.isfinite(g).all()]
        if nf or loss != loss:
            bad += 1
            print(f"rep {r} B={hb[0].shape[0]}: loss {loss}, non-finite grads in {len(nf)} tensors, first {nf[:3]}", flush=True)
print(f"dbg={dbg}: {bad} bad of {3 * reps} steps; second passes run by the verified hand-over: {_lib.query('asrb_debug_rnn_redos') - redos0}", flush=True)
