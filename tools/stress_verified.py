"""Verified hand-over of rnn3.cu under load: the configs[1] fwd+bwd on full ragged batches (batch 64 = two chains per SM, the
only case in which a late operand tile was ever seen), back to back for a wall-clock budget.  Counts the launches whose second
(release) pass had to run (asrb_debug_rnn_redos) and the steps whose loss differs from the one-pass release protocol's or
whose gradients are not finite.  The operand slabs are sentinel-filled (NaN) per launch, so a stale sector the check missed
would surface as a non-finite result.  Arguments: seconds of stepping (default 30), dbg bits (default 0 = verified)."""
import sys
import time

import torch

sys.path.insert(0, ".")
import bench
from asr_b200 import _lib
from tests.test_gpu_fullsize import _ragged_batch, _step

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
dbg = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cfg = dict(bench.CFG)
dev = torch.device("cuda")
model = bench.build_model(cfg, dev).train()
for m in model.modules():
    if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
        m.eval()
batches = []
for seed in (77, 78):
    b, _ = _ragged_batch(cfg["B"], cfg["T"], 60, cfg["C"], seed=seed)
    batches.append((b[0].to(dev),) + tuple(b[1:]))
_lib.query("asrb_debug_rnn_dbg", 4096)                 # one-pass release: the losses to reproduce
want = [_step(model, b)[0] for b in batches]
_lib.query("asrb_debug_rnn_dbg", dbg)
redos0 = _lib.query("asrb_debug_rnn_redos")
steps = bad = 0
t0 = last = time.time()
while time.time() - t0 < budget:
    for b, w in zip(batches, want):
        loss, grads = _step(model, b)
        finite = all(bool(torch.isfinite(s)) for s in torch.stack([g.sum() for g in grads.values()]).cpu())
        steps += 1
        if abs(loss - w) > 1e-6 * abs(w) or not finite:
            bad += 1
            print(f"step {steps}: loss {loss!r} (release: {w!r}), finite gradients: {finite}", flush=True)
    if time.time() - last > 5:
        last = time.time()
        print(f"... {steps} steps, {bad} bad, second passes {_lib.query('asrb_debug_rnn_redos') - redos0}", flush=True)
_lib.query("asrb_debug_rnn_dbg", 0)
print(f"dbg={dbg}: {steps} full-batch steps ({steps * 10} recurrent launches) in {time.time() - t0:.1f} s: {bad} bad; "
      f"second passes run by the verified hand-over: {_lib.query('asrb_debug_rnn_redos') - redos0}", flush=True)
