"""TEST INFRASTRUCTURE -- CPU oracle #2: the same path written out as explicit arithmetic
(no nn.GRU / nn.LSTM / F.ctc_loss / F.batch_norm / librosa), any dtype incl. float64.

Nothing under ``asr_b200/`` imports this file (allowed importers: ``tests/``, ``smoke()``,
``bench.py``'s CPU-baseline legs).  It pins the *published algorithms* of the third-party
routines the reference calls, independent of torch's fused kernels:

* GRU / LSTM cell equations -- torch.nn.GRU / torch.nn.LSTM documentation (gate row order
  ``[r|z|n]`` / ``[i|f|g|o]``), called by the reference at asr_deepspeech/modules/blocks.py:76-78,87-89.
* Packed-sequence semantics (pack_padded_sequence / pad_packed_sequence, blocks.py:87,89): an
  utterance takes part only for ``t < len``; the reverse direction starts at ``t = len-1`` from h=0;
  outputs past ``len`` are zero.
* Batch norm (training: biased batch variance; running estimate uses the unbiased one,
  momentum 0.1, eps 1e-5) -- torch.nn.BatchNorm{1,2}d, deepspeech.py:62,65,104; blocks.py:75.
* CTC (Graves et al. 2006) log-space alpha/beta over the blank-extended label sequence, and
  the gradient torch returns: ``exp(lp) - exp(logsumexp_{s:l'_s=c}(alpha+beta) + nll - lp)``
  (torch.nn.CTCLoss(reduction='sum'), trainers/__main__.py:53, called at
  trainers/deepspeech_trainer.py:111).
* STFT -> |.| -> log1p -> (x-mean)/std -- asr_deepspeech/data/parsers/spectrogram_parser.py:45-60
  on librosa 0.11.0 ``stft`` defaults (center=True zero padding of n_fft//2, periodic scipy
  window, 1 + len//hop frames, rfft).  librosa is NOT installed anywhere we can run, so this
  one function is **parity unpinned** against librosa itself; it is pinned only against
  ``torch.stft`` (an independent implementation of the same definition) in
  tests/test_oracle_golden.py.

Every function is checked against oracle/torch_path.py (and through it against the golden
vectors of the unmodified reference) in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import numpy as np
import torch

from .torch_path import BN_EPS, BN_MOMENTUM, CONV_SPECS, get_seq_lens, n_rnn_layers


# ----------------------------------------------------------------------------- spectrogram
def spectrogram(y: np.ndarray, sample_rate=16000, window_size=0.02, window_stride=0.01,
                window="hamming", normalize=True) -> torch.Tensor:
    """spectrogram_parser.py:45-60.  y: float32 mono waveform -> FloatTensor [1+n_fft/2, 1+len//hop]."""
    import scipy.signal

    n_fft = int(sample_rate * window_size)           # :45
    hop = int(sample_rate * window_stride)           # :47
    win = scipy.signal.get_window(window, n_fft, fftbins=True)   # librosa.filters.get_window
    y = np.asarray(y, dtype=np.float32)
    ypad = np.pad(y, (n_fft // 2, n_fft // 2), mode="constant")  # center=True, pad_mode="constant"
    n_frames = 1 + (len(ypad) - n_fft) // hop
    idx = np.arange(n_fft)[:, None] + hop * np.arange(n_frames)[None, :]
    frames = ypad[idx]                                            # [n_fft, n_frames]
    D = np.fft.rfft(win[:, None] * frames, axis=0).astype(np.complex64)   # :49-51 (complex64 result)
    spect = np.log1p(np.abs(D))                                   # magphase -> |D| ; log1p  (:52-54)
    spect = torch.from_numpy(spect).float()                       # :55
    if normalize:                                                 # :56-60 (torch unbiased std)
        spect = (spect - spect.mean()) / spect.std()
    return spect


# ----------------------------------------------------------------------------- building blocks
def time_mask(x, lengths):
    t = torch.arange(x.size(-1))
    keep = (t[None, :] < lengths[:, None]).to(x.dtype)
    return x * keep[:, None, None, :]


def batch_norm(x, weight, bias, running_mean, running_var, training, channel_dim=1):
    """Returns (y, new_running_mean, new_running_var)."""
    dims = [d for d in range(x.dim()) if d != channel_dim]
    shape = [1] * x.dim()
    shape[channel_dim] = -1
    if training:
        n = x.numel() // x.size(channel_dim)
        mean = x.mean(dims)
        var = ((x - mean.view(shape)) ** 2).mean(dims)           # biased
        new_rm = (1 - BN_MOMENTUM) * running_mean + BN_MOMENTUM * mean.detach()
        new_rv = (1 - BN_MOMENTUM) * running_var + BN_MOMENTUM * var.detach() * n / max(n - 1, 1)
    else:
        mean, var, new_rm, new_rv = running_mean, running_var, running_mean, running_var
    y = (x - mean.view(shape)) / torch.sqrt(var.view(shape) + BN_EPS) * weight.view(shape) + bias.view(shape)
    return y, new_rm, new_rv


def conv2d(x, w, b, stride, padding):
    """Direct cross-correlation via unfold (im2col) + matmul; torch.nn.Conv2d definition."""
    B, Cin, Hh, Ww = x.shape
    Cout, _, kh, kw = w.shape
    cols = torch.nn.functional.unfold(x, (kh, kw), padding=padding, stride=stride)  # [B, Cin*kh*kw, L]
    ho = (Hh + 2 * padding[0] - kh) // stride[0] + 1
    wo = (Ww + 2 * padding[1] - kw) // stride[1] + 1
    y = w.reshape(Cout, -1) @ cols + b.view(1, -1, 1)
    return y.view(B, Cout, ho, wo)


def gru_direction(gi, w_hh, b_hh, lengths, reverse):
    """gi = x W_ih^T + b_ih, [T,N,3H]; returns outputs [T,N,H] (zeros for t >= len)."""
    T, N, G = gi.shape
    H = G // 3
    h = gi.new_zeros(N, H)
    outs = [None] * T
    for t in (range(T - 1, -1, -1) if reverse else range(T)):
        gh = h @ w_hh.t() + b_hh
        r = torch.sigmoid(gi[t, :, :H] + gh[:, :H])
        z = torch.sigmoid(gi[t, :, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[t, :, 2 * H:] + r * gh[:, 2 * H:])
        h_new = (1 - z) * n + z * h
        m = (t < lengths).to(gi.dtype)[:, None]
        h = m * h_new + (1 - m) * h
        outs[t] = m * h_new
    return torch.stack(outs)


def lstm_direction(gi, w_hh, b_hh, lengths, reverse):
    T, N, G = gi.shape
    H = G // 4
    h = gi.new_zeros(N, H)
    c = gi.new_zeros(N, H)
    outs = [None] * T
    for t in (range(T - 1, -1, -1) if reverse else range(T)):
        g = gi[t] + h @ w_hh.t() + b_hh
        i_, f_, g_, o_ = (torch.sigmoid(g[:, :H]), torch.sigmoid(g[:, H:2 * H]),
                          torch.tanh(g[:, 2 * H:3 * H]), torch.sigmoid(g[:, 3 * H:]))
        c_new = f_ * c + i_ * g_
        h_new = o_ * torch.tanh(c_new)
        m = (t < lengths).to(gi.dtype)[:, None]
        h = m * h_new + (1 - m) * h
        c = m * c_new + (1 - m) * c
        outs[t] = m * h_new
    return torch.stack(outs)


def birnn_layer(x, lengths, w, rnn_type, bidirectional=True):
    """w: dict with weight_ih_l0, weight_hh_l0, bias_ih_l0, bias_hh_l0 [+ _reverse].  Sum of directions."""
    fn = gru_direction if rnn_type == "gru" else lstm_direction
    lens = lengths.to(torch.int64)
    out = None
    for sfx, rev in (("", False), ("_reverse", True)) if bidirectional else (("", False),):
        gi = x @ w["weight_ih_l0" + sfx].t() + w["bias_ih_l0" + sfx]
        o = fn(gi, w["weight_hh_l0" + sfx], w["bias_hh_l0" + sfx], lens, rev)
        out = o if out is None else out + o
    return out


def forward(p, x, lengths, rnn_type="gru", training=True, stats_out=None, conv_fn=conv2d):
    """Explicit-arithmetic DeepSpeech.forward.  Returns (out[N,T',C], output_lengths)."""
    out_lengths = get_seq_lens(lengths)
    lens = out_lengths.to(torch.int64)

    def bn(x, prefix, channel_dim=1):
        y, rm, rv = batch_norm(x, p[prefix + ".weight"], p[prefix + ".bias"], p[prefix + ".running_mean"],
                               p[prefix + ".running_var"], training, channel_dim)
        if stats_out is not None and training:
            stats_out[prefix + ".running_mean"], stats_out[prefix + ".running_var"] = rm, rv
        return y

    for spec in CONV_SPECS:
        x = conv_fn(x, p[f"conv.seq_module.{spec['idx']}.weight"], p[f"conv.seq_module.{spec['idx']}.bias"],
                    spec["stride"], spec["padding"])
        x = time_mask(x, lens)
        x = time_mask(bn(x, f"conv.seq_module.{spec['bn']}"), lens)
        x = time_mask(torch.clamp(x, 0.0, 20.0), lens)
    b, c, d, t = x.shape
    x = x.reshape(b, c * d, t).permute(2, 0, 1)
    for layer in range(n_rnn_layers(p)):
        pre = f"rnns.{layer}"
        if pre + ".batch_norm.module.weight" in p:
            T, N, I = x.shape
            x = bn(x.reshape(T * N, I), pre + ".batch_norm.module").reshape(T, N, I)
        w = {k[len(pre) + 5:]: v for k, v in p.items() if k.startswith(pre + ".rnn.")}
        x = birnn_layer(x, lens, w, rnn_type)
    T, N, H = x.shape
    x = bn(x.reshape(T * N, H), "fc.0.module.0") @ p["fc.0.module.1.weight"].t()
    x = x.view(T, N, -1).transpose(0, 1)
    if not training:
        x = torch.softmax(x, dim=-1)
    return x, out_lengths


# ----------------------------------------------------------------------------- CTC
def _logsumexp2(a, b):
    m = np.maximum(a, b)
    m_safe = np.where(np.isneginf(m), 0.0, m)
    return np.where(np.isneginf(m), -np.inf, m_safe + np.log(np.exp(a - m_safe) + np.exp(b - m_safe)))


def _shift(v, k):
    """v shifted right by k (k>0) or left (k<0), filled with -inf; same length as v."""
    out = np.full_like(v, -np.inf)
    if k > 0 and k < len(v):
        out[k:] = v[:-k]
    elif k < 0 and -k < len(v):
        out[:k] = v[-k:]
    elif k == 0:
        out[:] = v
    return out


def ctc(log_probs: np.ndarray, targets: np.ndarray, input_lengths, target_lengths, blank=0):
    """log_probs [T,N,C]; targets 1-D concatenated (functional.py:30-31 layout) or [N,U] padded.
    Returns (nll[N], grad[T,N,C]) with grad exactly the tensor torch's CTC backward produces for
    grad_output=1 (zero for t >= input_length; NaN-free for feasible alignments)."""
    lp = np.asarray(log_probs, dtype=np.float64)
    T, N, C = lp.shape
    targets = np.asarray(targets)
    nll = np.zeros(N)
    grad = np.zeros_like(lp)
    off = 0
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        for n in range(N):
            Tn, U = int(input_lengths[n]), int(target_lengths[n])
            if targets.ndim == 1:
                lab = targets[off:off + U]
                off += U
            else:
                lab = targets[n, :U]
            S = 2 * U + 1
            ext = np.full(S, blank, dtype=np.int64)
            ext[1::2] = lab
            # skip transition s-2 -> s allowed iff ext[s] != blank and ext[s] != ext[s-2]
            can_skip = np.zeros(S, dtype=bool)
            can_skip[2:] = (ext[2:] != blank) & (ext[2:] != ext[:-2])
            alpha = np.full((Tn, S), -np.inf)
            beta = np.full((Tn, S), -np.inf)
            if Tn > 0:
                alpha[0, 0] = lp[0, n, blank]
                if S > 1:
                    alpha[0, 1] = lp[0, n, ext[1]]
                for t in range(1, Tn):
                    a = alpha[t - 1]
                    a1 = _shift(a, 1)
                    a2 = np.where(can_skip, _shift(a, 2), -np.inf)
                    alpha[t] = _logsumexp2(_logsumexp2(a, a1), a2) + lp[t, n, ext]
                beta[Tn - 1, S - 1] = lp[Tn - 1, n, blank]
                if S > 1:
                    beta[Tn - 1, S - 2] = lp[Tn - 1, n, ext[S - 2]]
                for t in range(Tn - 2, -1, -1):
                    b = beta[t + 1]
                    b1 = _shift(b, -1)
                    skip_from = np.zeros(S, dtype=bool)
                    skip_from[:max(S - 2, 0)] = can_skip[2:]
                    b2 = np.where(skip_from, _shift(b, -2), -np.inf)
                    beta[t] = _logsumexp2(_logsumexp2(b, b1), b2) + lp[t, n, ext]
                ll = _logsumexp2(alpha[Tn - 1, S - 1], alpha[Tn - 1, S - 2] if S > 1 else -np.inf)
            else:
                ll = 0.0 if U == 0 else -np.inf
            nll[n] = -ll
            # gradient (w.r.t. the log_softmax INPUT; torch's formula)
            ab = alpha + beta                                    # [Tn,S]
            lcab = np.full((Tn, C), -np.inf)
            for s in range(S):
                lcab[:, ext[s]] = _logsumexp2(lcab[:, ext[s]], ab[:, s])
            grad[:Tn, n, :] = np.exp(lp[:Tn, n, :]) - np.exp(lcab + nll[n] - lp[:Tn, n, :])
    return nll, grad
