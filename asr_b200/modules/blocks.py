"""Drop-in counterparts of asr_deepspeech.modules.blocks (same class names, constructor signatures, parameter
names and state_dict keys -- SURVEY.md section 8b), with every arithmetic step running on our sm_100a kernels.

The torch.nn modules held inside (nn.Conv2d, nn.BatchNorm2d, nn.GRU, ...) are used ONLY as parameter
containers so that reference checkpoints load with strict=True; their forward() is never called.
CPU tensors raise: there is no fallback path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import functional as F_
from .. import ops


def _lengths_dev(lengths, device):
    if isinstance(lengths, torch.Tensor) and lengths.is_cuda and lengths.dtype == torch.int32:
        return lengths
    return ops.lengths_to_device(lengths, device)


def _need_cuda(x, who):
    ops.require_cuda(x, who)


def _bn_tick(bn: nn.modules.batchnorm._BatchNorm):
    if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked += 1


def _bn_momentum(bn):
    if bn.momentum is None:
        raise ValueError("cumulative-average BatchNorm (momentum=None) is not supported")
    return float(bn.momentum)


def batch_norm_rows(bn: nn.BatchNorm1d, x):
    """BatchNorm1d over the rows of x[..., H] with the module's parameters / running stats."""
    training = bn.training or not bn.track_running_stats
    y = F_.BatchNormRows.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, training,
                               _bn_momentum(bn), float(bn.eps))
    _bn_tick(bn)
    return y


class SequenceWise(nn.Module):
    """asr_deepspeech/modules/blocks.py:6-27.  Collapses T*N*H to (T*N)*H and applies `module`;
    supported modules: BatchNorm1d, Linear, Sequential of those (everything the reference wraps)."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def _apply_one(self, m, x):
        if isinstance(m, nn.BatchNorm1d):
            return batch_norm_rows(m, x)
        if isinstance(m, nn.Linear):
            return F_.LinearRows.apply(x, m.weight, m.bias)
        if isinstance(m, nn.Sequential):
            for sub in m:
                x = self._apply_one(sub, x)
            return x
        raise TypeError(f"SequenceWise: no sm_100a kernel for {type(m).__name__}")

    def forward(self, x):
        _need_cuda(x, "SequenceWise")
        t, n = x.size(0), x.size(1)
        y = self._apply_one(self.module, x.reshape(t * n, -1))
        return y.view(t, n, -1)

    def __repr__(self):
        return self.__class__.__name__ + " (\n" + self.module.__repr__() + ")"


class MaskConv(nn.Module):
    """asr_deepspeech/modules/blocks.py:30-56: every module of `seq_module` is followed by zeroing
    x[i, :, :, lengths[i]:].  Conv2d, BatchNorm2d and Hardtanh are supported (fused kernels); input BxCxDxT."""

    def __init__(self, seq_module):
        super().__init__()
        self.seq_module = seq_module

    def forward(self, x, lengths):
        _need_cuda(x, "MaskConv")
        ldev = _lengths_dev(lengths, x.device)
        mods = list(self.seq_module)
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.Conv2d):
                if m.dilation != (1, 1) or m.groups != 1 or m.padding_mode != "zeros" or isinstance(m.padding, str):
                    raise ValueError("MaskConv: only dense zero-padded Conv2d (dilation 1, groups 1) is supported")
                x = F_.Conv2dMask.apply(x, m.weight, m.bias, ldev, tuple(m.stride), tuple(m.padding))
                i += 1
            elif isinstance(m, (nn.BatchNorm2d, nn.Hardtanh)):
                bn = m if isinstance(m, nn.BatchNorm2d) else None
                act = None
                j = i + 1
                if bn is not None and j < len(mods) and isinstance(mods[j], nn.Hardtanh):
                    act = mods[j]
                    j += 1
                elif bn is None:
                    act = m
                lo, hi = (float(act.min_val), float(act.max_val)) if act is not None else (0.0, 0.0)
                if bn is not None:
                    training = bn.training or not bn.track_running_stats
                    x = F_.BnActMask.apply(x, ldev, bn.weight, bn.bias, bn.running_mean, bn.running_var, True,
                                           act is not None, lo, hi, training, _bn_momentum(bn), float(bn.eps))
                    _bn_tick(bn)
                else:
                    x = F_.BnActMask.apply(x, ldev, None, None, None, None, False, True, lo, hi, False, 0.0, 0.0)
                i = j
            else:
                raise TypeError(f"MaskConv: no sm_100a kernel for {type(m).__name__}")
        return x, lengths


class InferenceBatchSoftmax(nn.Module):
    """asr_deepspeech/modules/blocks.py:59-64: identity in training, softmax(dim=-1) in eval."""

    def forward(self, input_):
        if self.training:
            return input_
        _need_cuda(input_, "InferenceBatchSoftmax")
        return F_.softmax_last_dim(input_)


class BatchRNN(nn.Module):
    """asr_deepspeech/modules/blocks.py:67-93.  [BatchNorm1d over rows] -> bidirectional GRU/LSTM over the packed
    (length-masked) sequences -> directions summed.  `self.rnn` keeps torch's parameter names
    (weight_ih_l0, weight_hh_l0, bias_ih_l0, bias_hh_l0 [+ _reverse])."""

    def __init__(self, input_size, hidden_size, rnn_type=nn.LSTM, bidirectional=False, batch_norm=True):
        super().__init__()
        if rnn_type not in (nn.LSTM, nn.GRU):
            raise ValueError(f"BatchRNN: rnn_type must be nn.LSTM or nn.GRU, got {rnn_type!r}")
        self.input_size = input_size
        self.hidden_size = hidden_size
        self._bidirectional = bidirectional
        self.batch_norm = SequenceWise(nn.BatchNorm1d(input_size)) if batch_norm else None
        self.rnn = rnn_type(input_size=input_size, hidden_size=hidden_size, bidirectional=bidirectional, bias=True)
        self.num_directions = 2 if bidirectional else 1
        self._cell = ops.LSTM if rnn_type is nn.LSTM else ops.GRU

    def flatten_parameters(self):
        pass  # nothing to flatten: weights are re-laid-out for the persistent kernel at every call

    def forward(self, x, output_lengths):
        _need_cuda(x, "BatchRNN")
        if self.batch_norm is not None:
            x = self.batch_norm(x)
        ldev = _lengths_dev(output_lengths, x.device)
        r = self.rnn
        if self._bidirectional:
            rev = (r.weight_ih_l0_reverse, r.weight_hh_l0_reverse, r.bias_ih_l0_reverse, r.bias_hh_l0_reverse)
        else:
            # Unidirectional layer on the two-direction kernel: a reverse direction whose weights and biases are all
            # zero keeps its state at exactly 0 (GRU: n = tanh(0) = 0, h' = z h = 0; LSTM: g = 0, c' = f c = 0,
            # h' = o tanh(0) = 0), so the sum of the directions IS the forward direction.  The two directions run on
            # different SMs at the same time, so this costs SMs, not time.
            rev = tuple(torch.zeros_like(t) for t in (r.weight_ih_l0, r.weight_hh_l0, r.bias_ih_l0, r.bias_hh_l0))
        y = F_.BiRnnLayer.apply(x, ldev, self._cell, r.weight_ih_l0, r.weight_hh_l0, r.bias_ih_l0, r.bias_hh_l0, *rev)
        t_max = int(torch.as_tensor(output_lengths).max())
        return y[:t_max] if t_max < y.size(0) else y   # pad_packed_sequence trims to the longest sequence


class Lookahead(nn.Module):
    """asr_deepspeech/modules/blocks.py:96-132 (Wang et al. 2016): depthwise convolution over time looking `context`
    frames ahead, [T,N,H] -> [T,N,H].  `self.conv` keeps the reference's parameter (conv.weight [H,1,context])."""

    def __init__(self, n_features, context):
        super().__init__()
        assert context > 0
        self.context = context
        self.n_features = n_features
        self.pad = (0, self.context - 1)
        self.conv = nn.Conv1d(self.n_features, self.n_features, kernel_size=self.context, stride=1,
                              groups=self.n_features, padding=0, bias=None)

    def forward(self, x, act=None):
        """act=(lo, hi): also apply the Hardtanh that follows this layer in the model (one fused pass)"""
        _need_cuda(x, "Lookahead")
        return F_.LookaheadConv.apply(x, self.conv.weight, self.context, act)

    def __repr__(self):
        return f"{self.__class__.__name__}(n_features={self.n_features}, context={self.context})"
