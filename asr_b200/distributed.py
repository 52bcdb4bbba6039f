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


class OverlappedGradSync:
    """The one gradient all-reduce of data-parallel training, cut in two so that it hides under the backward pass: the
    tail of the flat bucket (recurrent layers + head: 99.7 % of the bytes) is reduced on a communication stream as soon as
    the recurrent stack's backward is done -- while the conv backward still runs -- and the few conv gradients follow at
    the end.  Same arithmetic as `FlatGradBucket.all_reduce_mean` (the reference's SUM-then-divide rule, functional.py:35-42).

        sync = OverlappedGradSync(bucket, model)       # sets model.after_rnn_backward
        ... bucket.zero(); loss.backward(); sync.finish()
    """

    def __init__(self, bucket, model, group=None):
        self.bucket, self.group = bucket, group
        conv_ids = {id(p) for p in model.conv.parameters()}
        head = 0
        for p in bucket.params:          # bucket order = model.parameters() order: the conv stack comes first
            if id(p) not in conv_ids:
                break
            head += p.numel()
        self.head = head
        self.work = None
        self.comm = torch.cuda.Stream(device=bucket.flat.device) if bucket.flat.is_cuda else None
        model.after_rnn_backward = self.start_tail

    def _active(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def start_tail(self):
        if not self._active() or self.comm is None:
            return
        from . import functional as F_

        dev = self.bucket.flat.device
        self.comm.wait_stream(torch.cuda.current_stream(dev))
        side = F_._side_streams.get(dev.index if dev.index is not None else torch.cuda.current_device())
        if side is not None:             # the weight gradients of the recurrent layers are issued on the side stream
            self.comm.wait_stream(side)
        with torch.cuda.stream(self.comm):
            self.work = dist.all_reduce(self.bucket.flat[self.head:], op=dist.ReduceOp.AVG, group=self.group, async_op=True)

    def finish(self):
        """after backward(): the rest of the bucket, and the main stream waits for the overlapped part"""
        if not self._active():
            return
        flat = self.bucket.flat
        if self.work is None:            # the hook did not fire (no gradient into the recurrent stack's input): one piece
            self.bucket.all_reduce_mean(self.group)
            return
        if self.head:
            if flat.is_cuda:
                dist.all_reduce(flat[:self.head], op=dist.ReduceOp.AVG, group=self.group)
            else:
                dist.all_reduce(flat[:self.head], op=dist.ReduceOp.SUM, group=self.group)
                flat[:self.head].div_(dist.get_world_size(self.group))
        self.work.wait()
        self.work = None


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
