"""Data-parallel plumbing for the training step: one process per GPU, the batch dimension
sharded, ONE all-reduce over a single flat gradient buffer (NCCL over NVLink on GPUs; gloo in the CPU tests).

The reference has no live distributed path; its dead remnants state the rule we follow: every rank takes a
disjoint set of utterance bins (data/samplers/distributed_bucketing_sampler.py:28-34) and reduced values are
SUM-then-divided by the world size (functional.py:35-42).  BatchNorm statistics stay per rank (no SyncBN in the
reference).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class FlatGradBucket:
    """All parameter gradients live in one contiguous fp32 buffer (`flat`); `param.grad` are views into it,
    so autograd accumulates straight into the bucket and the all-reduce needs no packing."""

    def __init__(self, params, flatten_params=False):
        """flatten_params: also move the parameters themselves into one contiguous buffer (`flat_params`, same
        order; `param.data` become views), so that `asr_b200.optim.FusedAdamW` updates the whole model in one launch."""
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_params = torch.empty(n, device=dev, dtype=torch.float32) if flatten_params else None
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            if flatten_params:
                dst = self.flat_params[off:off + p.numel()].view_as(p)
                dst.copy_(p.data)
                p.data = dst
            off += p.numel()

    def zero(self):
        """Re-arm before backward: zero the bucket and (re)attach the views as .grad."""
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            p.grad = v

    def all_reduce_mean(self, group=None):
        """grad <- mean over ranks (each rank's loss is already averaged over its local batch,
        trainers/deepspeech_trainer.py:112, so this is the global-batch mean)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            if self.flat.is_cuda:
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)
            else:  # gloo has no AVG
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
                self.flat.div_(dist.get_world_size(group))


def frame_balanced_shards(num_frames, world_size):
    """Assign utterances to ranks so that every rank gets the same COUNT and a nearly equal SUM of frames
    (longest-first greedy with a per-rank capacity).  Whole-bin dealing, the reference's rule, leaves a 3.3x
    straggler on bucketed 5-20 s batches.  Returns a list of index lists, each
    sorted by decreasing length (the order pack_padded_sequence / _collate_fn require, functional.py:13)."""
    n = len(num_frames)
    if n % world_size != 0:
        raise ValueError("global batch must be divisible by the world size")
    cap = n // world_size
    order = sorted(range(n), key=lambda i: (-num_frames[i], i))
    loads = [0] * world_size
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = min((r for r in range(world_size) if len(shards[r]) < cap), key=lambda r: (loads[r], r))
        shards[r].append(i)
        loads[r] += num_frames[i]
    return shards
