"""Fused AdamW on our kernel: drop-in for the optimizer the reference builds at
asr_deepspeech/trainers/__main__.py:41-47 (`AdamW(lr=1.5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5)`) and
steps at trainers/deepspeech_trainer.py:86-95.

It is a `torch.optim.Optimizer` (so `StepLR`, `GradScaler.step(optimizer)` and `state_dict()` keep working), with
torch.optim.AdamW's state layout (`step`, `exp_avg`, `exp_avg_sq` per parameter).  `step()` issues ONE kernel per
parameter tensor -- or ONE kernel for the whole model when the parameters and gradients were flattened with
`FlatGradBucket(..., flatten_params=True)`: 28 bytes per parameter, a single HBM pass.  In the flat form the per-parameter
`exp_avg` / `exp_avg_sq` of `self.state` are VIEWS into two flat moment buffers, so `state_dict()` saves them and
`load_state_dict()` (which replaces them by copies) re-flattens: a resumed run continues with its moments and step count.
"""
from __future__ import annotations

import torch

from . import ops


def _bump_versions(params):
    """The kernel writes the parameters behind autograd's back: bump their version counters like an in-place op would, so
    that anything keyed on `_version` (functional._cached: packed / concatenated weights) sees the update."""
    for p in params:
        torch.autograd.graph.increment_version(p)


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, bucket=None):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.bucket = bucket if bucket is not None and getattr(bucket, "flat_params", None) is not None else None
        self._flat_state = None

    def _flatten_state(self):
        """flat form: the moments of all parameters in two flat buffers, `self.state[p]` holding views into them; state
        that is already there (first step after `load_state_dict`) is copied in"""
        if self._flat_state is not None:
            return
        flat = self.bucket.flat_params
        st = dict(step=0, exp_avg=torch.zeros_like(flat), exp_avg_sq=torch.zeros_like(flat))
        off = 0
        for p in self.bucket.params:
            n = p.numel()
            views = {k: st[k][off:off + n].view_as(p) for k in ("exp_avg", "exp_avg_sq")}
            old = self.state.get(p)
            if old:
                for k, v in views.items():
                    v.copy_(old[k])
                st["step"] = max(st["step"], int(old["step"]))
            self.state[p] = dict(step=st["step"], **views)
            off += n
        for p in self.bucket.params:
            self.state[p]["step"] = st["step"]
        self._flat_state = st

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._flat_state = None          # the loaded moments are copies: re-flatten them at the next step

    @torch.no_grad()
    def step(self, closure=None, inv_scale=None):
        """inv_scale: optional 1-element CUDA tensor multiplied into every gradient (GradScaler's unscale, fused)."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self.bucket is not None and len(self.param_groups) == 1:
            g = self.param_groups[0]
            self._flatten_state()
            for p, v in zip(self.bucket.params, self.bucket.views):
                if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                    raise RuntimeError("FusedAdamW(bucket=...): a parameter's .grad is not the bucket's view any more (was "
                                       "zero_grad(set_to_none=True) called?  use bucket.zero() to re-arm the gradients)")
            st = self._flat_state
            st["step"] += 1
            for p in self.bucket.params:
                self.state[p]["step"] = st["step"]
            ops.adamw_step(self.bucket.flat_params, self.bucket.flat, st["exp_avg"], st["exp_avg_sq"], g["lr"], g["betas"][0],
                           g["betas"][1], g["eps"], g["weight_decay"], st["step"], inv_scale)
            _bump_versions(self.bucket.params)
            return loss
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is None:
                    continue
                ops.require_cuda(p, "FusedAdamW")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                if not p.is_contiguous() or not p.grad.is_contiguous():
                    raise ValueError("FusedAdamW needs contiguous parameters and gradients")
                ops.adamw_step(p, p.grad, st["exp_avg"], st["exp_avg_sq"], g["lr"], g["betas"][0], g["betas"][1], g["eps"],
                               g["weight_decay"], st["step"], inv_scale)
                _bump_versions((p,))
        return loss
