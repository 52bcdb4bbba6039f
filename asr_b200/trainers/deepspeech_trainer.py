"""The criterion / fit() seam of asr_deepspeech.trainers (trainers/__main__.py:53, deepspeech_trainer.py:102-117).

* `CTCLoss`  -- drop-in for torch.nn.CTCLoss(blank=0, reduction="sum", zero_infinity=False) as the reference
  constructs it: called as criterion(log_probs[T,N,C], targets int32 cpu [sum U], input_lengths int32 cpu [N],
  target_lengths int32 cpu [N]) and returns a 0-dim tensor with grad_fn.  Backward hands back what torch hands
  back for log_probs (the gradient w.r.t. the logits), so the reference's `.log_softmax(2)` upstream of it keeps
  working unchanged.
* `fit`      -- the six arithmetic lines of DeepSpeechTrainer.fit on our kernels (log_softmax included), same
  return triple (valid_loss, loss, loss_value).
* `DeepSpeechStep` -- one whole training step (fit + backward [+ all-reduce] + optimizer), what bench.py times.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import functional as F_
from .. import ops


class CTCLoss(nn.Module):
    def __init__(self, blank: int = 0, reduction: str = "sum", zero_infinity: bool = False):
        super().__init__()
        if reduction != "sum":
            raise ValueError("asr_b200.CTCLoss implements reduction='sum' (the reference's setting, trainers/__main__.py:53)")
        self.blank = blank
        self.reduction = reduction
        self.zero_infinity = zero_infinity

    def forward(self, log_probs, targets, input_lengths, target_lengths):
        ops.require_cuda(log_probs, "CTCLoss")
        if targets.dim() != 1:
            raise ValueError("targets must be the 1-D concatenation of all label sequences (functional.py:30-31)")
        dev = log_probs.device
        tl_host = torch.as_tensor(target_lengths).cpu()
        max_u = int(tl_host.max()) if tl_host.numel() else 0
        t_dev = ops.lengths_to_device(targets, dev)
        il_dev = ops.lengths_to_device(input_lengths, dev)
        tl_dev = ops.lengths_to_device(tl_host, dev)
        return F_.CtcLossSum.apply(log_probs, t_dev, il_dev, tl_dev, max_u, self.blank, self.zero_infinity)


def fit(model, criterion, data, device):
    """DeepSpeechTrainer.fit (deepspeech_trainer.py:102-117)."""
    inputs, targets, input_percentages, target_sizes = data
    input_sizes = input_percentages.mul(int(inputs.size(3))).int()  # (the reference uses the in-place mul_)
    inputs = inputs.to(device, non_blocking=True)
    out, output_sizes = model.forward(inputs, input_sizes)
    out = out.transpose(0, 1)                       # T x N x C
    log_probs = F_.LogSoftmaxLastDim.apply(out.float())
    loss = criterion(log_probs, targets, output_sizes, target_sizes).to(device)
    loss = loss / inputs.size(0)                    # average the loss by minibatch
    loss_value = loss.item()
    valid_loss, _ = F_.check_loss(loss, loss_value)
    return valid_loss, loss, loss_value


class DeepSpeechStep:
    """fit -> zero_grad -> backward -> [gradient all-reduce over the data-parallel group] -> optimizer.step,
    i.e. one iteration of DeepSpeechTrainer.train (deepspeech_trainer.py:77-97).  mixed_precision=True reproduces the
    reference's CUDA AMP branch (:80-91: fit() under fp16 autocast, GradScaler around backward/step); our operators
    keep computing in fp32/TF32 under autocast (asr_b200/functional.py `_amp_fwd`), so only the loss scaling is live."""

    def __init__(self, model, criterion=None, optimizer=None, device="cuda", grad_sync=None, mixed_precision=False,
                 bucket=None):
        """bucket: an asr_b200.distributed.FlatGradBucket over the model's parameters.  With a bucket the gradients must
        stay the bucket's views: the step re-arms it with `bucket.zero()` instead of `zero_grad(set_to_none=True)` (which
        would detach the views and leave the all-reduce and the flat FusedAdamW reading a stale buffer), and the
        all-reduce defaults to `bucket.all_reduce_mean`."""
        self.model, self.device = model, torch.device(device)
        self.criterion = criterion if criterion is not None else CTCLoss(reduction="sum")
        self.optimizer = optimizer
        self.bucket = bucket
        if bucket is not None and grad_sync is None:
            grad_sync = lambda _model: bucket.all_reduce_mean()   # noqa: E731
        self.grad_sync = grad_sync
        self.use_amp = bool(mixed_precision) and self.device.type == "cuda"
        self.scaler = torch.amp.GradScaler("cuda", enabled=True) if self.use_amp else None

    def __call__(self, data):
        if self.use_amp:
            with torch.amp.autocast("cuda", enabled=True):
                valid, loss, loss_value = fit(self.model, self.criterion, data, self.device)
        else:
            valid, loss, loss_value = fit(self.model, self.criterion, data, self.device)
        if valid:
            if self.bucket is not None:
                self.bucket.zero()
            elif self.optimizer is not None:
                self.optimizer.zero_grad(set_to_none=True)
            else:
                for prm in self.model.parameters():
                    prm.grad = None
            if self.use_amp:
                self.scaler.scale(loss).backward()
            else:
                loss.backward()
            if self.grad_sync is not None:
                self.grad_sync(self.model)
            if self.optimizer is not None:
                if self.use_amp:
                    self.scaler.step(self.optimizer)
                    self.scaler.update()
                else:
                    self.optimizer.step()
        else:
            print("Loss non valid, skipped")
        return valid, loss_value
