#!/usr/bin/env python
"""One training step of the bench workload between cudaProfilerStart/Stop, for `ncu --profile-from-start off`.
Usage: python tools/profile_step.py [--batch 64] [--frames 1001]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=bench.CFG["B"])
    ap.add_argument("--frames", type=int, default=bench.CFG["T"])
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    from asr_b200.trainers import CTCLoss, fit
    from asr_b200 import _lib
    if os.environ.get("ASRB_RNN_KSPLIT", "1") == "0":   # ncu cannot launch the cooperative + cluster backward kernel
        _lib.query("asrb_debug_rnn_ksplit", 0)

    cfg = dict(bench.CFG, B=args.batch, T=args.frames, U=min(bench.CFG["U"], args.frames // 10))
    dev = torch.device("cuda", 0)
    model = bench.build_model(cfg, dev)
    crit = CTCLoss()
    host = bench.make_batch(cfg["B"], cfg)
    x = host[0].to(dev)

    def step():
        for p in model.parameters():
            p.grad = None
        _, loss, _ = fit(model, crit, (x, host[1], host[2], host[3]), dev)
        loss.backward()

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
