"""Settles every greedy-decode index that differs from the reference's fp32 argmax (decoders/greedy_decoder.py:61)
against the fp64 run of the same unmodified reference modules that oracle/make_golden.py stores in each golden record
(`eval64_top2`, `eval64_margin`, `eval64_logp_digest`).

A frame (n, t) where the device picked class a and the reference's fp32 run picked r != a is accepted only if
  (1) a is the fp64 winner                      -- the fp32 reference is the one that left the exact result, or
  (2) a is the fp64 runner-up and the fp64 log-probability margin between winner and runner-up is at most `tol`,
where tol = 4 x the largest error of the device's own log-probabilities against fp64 over the 64 sampled entries of the
record (what the arithmetic demonstrably cannot resolve), capped at `max_tol`.  Anything else raises."""
import torch

from oracle.make_golden import sample_idx


def settle_argmax_flips(probs, sizes, g, max_tol=5e-3):
    """probs [N,T,C] float cpu (eval-mode output), sizes list[int], g golden record -> (flips, by_fp64, ties, checked, tol)"""
    logp = probs.clamp_min(1e-45).log()
    flat = logp.flatten()
    d = g["eval64_logp_digest"]
    tol = 4.0 * (flat[sample_idx(flat.numel())] - d["samples"]).abs().max().item()
    assert tol <= max_tol, f"device log-probabilities are off by {tol / 4:.2e} at the sampled entries"
    idx = torch.max(probs, 2)[1]
    ref32 = g["eval_argmax"].long()
    win, second = g["eval64_top2"][..., 0].long(), g["eval64_top2"][..., 1].long()
    margin = g["eval64_margin"]
    flips = by_fp64 = ties = checked = 0
    for n, tn in enumerate(sizes):
        checked += tn
        bad = (idx[n, :tn] != ref32[n, :tn]).nonzero().flatten().tolist()
        for t in bad:
            flips += 1
            a = int(idx[n, t])
            if a == int(win[n, t]):
                by_fp64 += 1
            elif a == int(second[n, t]) and float(margin[n, t]) <= tol:
                ties += 1
            else:
                raise AssertionError(f"frame ({n},{t}): device picked {a}, reference fp32 {int(ref32[n, t])}, fp64 winner "
                                     f"{int(win[n, t])} / runner-up {int(second[n, t])} with margin {float(margin[n, t]):.3e} > tol {tol:.3e}")
    return flips, by_fp64, ties, checked, tol
