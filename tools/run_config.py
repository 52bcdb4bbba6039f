#!/usr/bin/env python
"""One training step (fwd + CTC + bwd) of an arbitrary BASELINE.json-style configuration, timed with CUDA events.
   python tools/run_config.py --cell lstm --hidden 1024 --layers 7 --batch 128 --seconds 15 --classes 90 --target-len 150"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--cell", default="gru"); ap.add_argument("--hidden", type=int, default=800)
ap.add_argument("--layers", type=int, default=5); ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--seconds", type=int, default=10); ap.add_argument("--classes", type=int, default=29)
ap.add_argument("--target-len", type=int, default=100); ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--ragged", action="store_true", help="utterance lengths uniform in [seconds/4, seconds], sorted descending")
a = ap.parse_args()
from asr_b200.trainers import CTCLoss, fit
from oracle.make_golden import synth_batch
cfg = dict(rnn_type=a.cell, hidden=a.hidden, layers=a.layers, C=a.classes, B=a.batch, seconds=a.seconds, T=100 * a.seconds + 1,
           U=a.target_len, seed=1237)
dev = torch.device("cuda", 0)
model = bench.build_model(cfg, dev)
lens = None
if a.ragged:
    g = torch.Generator().manual_seed(5)
    lens = sorted((torch.randint(cfg["T"] // 4, cfg["T"] + 1, (a.batch,), generator=g)).tolist(), reverse=True)
    lens[0] = cfg["T"]
host = synth_batch(cfg["seed"], a.batch, cfg["T"], min(a.target_len, cfg["T"] // 8), a.classes, lens)
x = host[0].to(dev)
crit = CTCLoss(reduction="sum")
def step():
    for p in model.parameters(): p.grad = None
    _, loss, lv = fit(model, crit, (x, host[1], host[2], host[3]), dev)
    loss.backward()
    return lv
for _ in range(2): lv = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps): lv = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
audio = float(sum(lens)) / 100.0 if lens else a.batch * a.seconds
gn = sum(float(p.grad.norm()) ** 2 for p in model.parameters()) ** 0.5
print(json.dumps({"config": vars(a), "ms_per_step": ms, "utt_sec_per_s": audio / (ms * 1e-3), "loss": lv, "grad_norm": gn,
                  "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}))
