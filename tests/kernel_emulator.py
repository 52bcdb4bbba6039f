"""TEST INFRASTRUCTURE (CPU only, never imported by asr_b200): a torch-CPU emulation of every wrapper in
asr_b200/ops.py, with the SAME buffer layouts and the SAME formulas the CUDA kernels implement
(asr_b200/csrc/*.cu).  It exists so that the host-side logic -- autograd plumbing, module wiring, state
layouts, the hand-derived backward formulas of the GRU/LSTM/BatchNorm/CTC kernels -- can be checked against
the oracle and the reference's golden vectors in the `-m "not gpu"` suite, where no GPU exists.

`install(monkeypatch)` swaps the functions in asr_b200.ops for the emulated ones and lifts the "CUDA tensor
required" guards; nothing outside tests/ can reach this file.  GPU parity is proven separately by the
`-m gpu` tests, which call the real library.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

GRU, LSTM = 0, 1


def gemm_tn(A, B, out=None, bias=None, accumulate=False):
    r = A @ B.t()
    if bias is not None:
        r = r + bias
    if out is None:
        return r
    if accumulate:
        out += r
    else:
        out.copy_(r)
    return out


def transpose(x, out=None):
    if out is None:
        return x.t().contiguous()
    out[:, : x.shape[0]].copy_(x.t())
    return out


def split3(x, mode):
    hi = ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    lo = x - hi
    return torch.cat([hi, lo, hi] if mode == 0 else [hi, hi, lo], dim=1)


def col_sums(a, cols=None):
    return a.sum(0)


def conv_out_size(n, k, s, p):
    return (n + 2 * p - k) // s + 1


def _tmask(x, lengths):
    if lengths is None:
        return torch.ones_like(x)
    t = torch.arange(x.shape[-1])
    return (t[None, :] < lengths[:, None].long()).to(x.dtype)[:, None, None, :].expand_as(x)


def conv2d_mask_fwd(x, w, bias, lengths, stride, padding):
    y = F.conv2d(x, w, bias, stride=stride, padding=padding)
    return y * _tmask(y, lengths)


def conv2d_mask_bwd_data(dy, w, lengths, x_shape, stride, padding):
    dym = dy * _tmask(dy, lengths)
    return torch.nn.grad.conv2d_input(x_shape, w, dym, stride=stride, padding=padding)


def conv2d_mask_bwd_weight(dy, x, lengths, w_shape, stride, padding, need_bias=True):
    dym = dy * _tmask(dy, lengths)
    dw = torch.nn.grad.conv2d_weight(x, w_shape, dym, stride=stride, padding=padding)
    return dw, (dym.sum((0, 2, 3)) if need_bias else None)


def conv32_supported(w_shape, stride, padding):
    return w_shape[0] == 32 and w_shape[1] == 32 and stride[1] == 1 and w_shape[3] <= 16


def conv1_supported(x_shape, w_shape, stride, padding):
    return w_shape[1] == 1 and w_shape[0] == 32 and w_shape[3] == 11 and tuple(stride) == (2, 2) and padding[1] == 5


def conv1_fwd(x, w, bias, lengths, stride, padding):
    return conv2d_mask_fwd(x, w, bias, lengths, stride, padding)


def conv1_bwd_weight(x, dy_masked, w_shape, padding):
    return torch.nn.grad.conv2d_weight(x, w_shape, dy_masked, stride=(2, 2), padding=padding)


def nchw_to_nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def pad_rows4(x):
    W = x.shape[-1]
    W4 = (W + 3) // 4 * 4
    return (x, W) if W4 == W else (F.pad(x, (0, W4 - W)), W4)


def mask_time(x, lengths):
    return x * _tmask(x, lengths)


def nchw_channel_sums(a, lengths):
    return (a * _tmask(a, lengths)).sum((0, 2, 3))


def conv32_pack_weights(w, fwd=True, dgrad=True):
    return (w if fwd else None), (w if dgrad else None)


CONV_ROWS = 4


def conv32_pack_rows(pack, w_shape, stride_h, mode, rows=None):
    return pack


def conv32_fwd(x_nhwc, pack_fwd, bias, lengths, w_shape, stride, padding, rows=0):
    return conv2d_mask_fwd(x_nhwc.permute(0, 3, 1, 2), pack_fwd, bias, lengths, stride, padding)


def conv32_bwd_data(dy_nhwc, pack_dgrad, x_shape, w_shape, stride, padding, rows=0):
    return torch.nn.grad.conv2d_input(x_shape, pack_dgrad, dy_nhwc.permute(0, 3, 1, 2), stride=stride, padding=padding)


def conv32_bwd_weight(x, dy_masked, w_shape, stride, padding):
    return torch.nn.grad.conv2d_weight(x, w_shape, dy_masked, stride=stride, padding=padding)


def _finalize_stats(s0, s1, count, running_mean, running_var, momentum, eps):
    mean = s0 / count
    var = (s1 / count - mean * mean).clamp_min(0)
    if running_mean is not None:
        unb = var * count / (count - 1) if count > 1 else var
        running_mean.mul_(1 - momentum).add_(momentum * mean.float())
        running_var.mul_(1 - momentum).add_(momentum * unb.float())
    return mean.float(), (1.0 / torch.sqrt(var + eps)).float()


def bn2d_stats(y, running_mean, running_var, training, momentum=0.1, eps=1e-5):
    if not training:
        return running_mean.clone(), torch.rsqrt(running_var + eps)
    yd = y.double()
    n = y.numel() // y.shape[1]
    return _finalize_stats(yd.sum((0, 2, 3)), (yd * yd).sum((0, 2, 3)), n, running_mean, running_var, momentum, eps)


def _bn_hat(y, mean, invstd, gamma, beta, has_bn):
    if not has_bn:
        return y, y
    sh = (1, -1, 1, 1)
    xh = (y - mean.view(sh)) * invstd.view(sh)
    return xh, xh * gamma.view(sh) + beta.view(sh)


def bn_act_mask_fwd(y, lengths, mean, invstd, gamma, beta, has_bn, has_act, lo, hi):
    _, v = _bn_hat(y, mean, invstd, gamma, beta, has_bn)
    if has_act:
        v = v.clamp(lo, hi)
    return v * _tmask(y, lengths)


def bn_act_mask_bwd(dz, y, lengths, mean, invstd, gamma, beta, has_bn, has_act, lo, hi, training):
    m = _tmask(y, lengths)
    xh, yh = _bn_hat(y, mean, invstd, gamma, beta, has_bn)
    g = dz * m
    if has_act:
        g = g * ((yh > lo) & (yh < hi)).to(g.dtype)
    if not has_bn:
        return g, None, None
    s0 = g.sum((0, 2, 3))
    s1 = (g * xh).sum((0, 2, 3))
    count = y.numel() // y.shape[1]
    sh = (1, -1, 1, 1)
    if training:
        g = g - s0.view(sh) / count - xh * s1.view(sh) / count
    dy = g * gamma.view(sh) * invstd.view(sh) * m
    return dy, s1, s0


def nchw_to_tnf(x):
    B, C, D, T = x.shape
    return x.reshape(B, C * D, T).permute(2, 0, 1).contiguous()


def tnf_to_nchw(x, C, D):
    T, B, _ = x.shape
    return x.permute(1, 2, 0).reshape(B, C, D, T).contiguous()


def bn_rows_fwd(x, gamma, beta, running_mean, running_var, training, momentum=0.1, eps=1e-5):
    if training:
        xd = x.double()
        mean, invstd = _finalize_stats(xd.sum(0), (xd * xd).sum(0), x.shape[0], running_mean, running_var, momentum, eps)
    else:
        mean, invstd = running_mean.clone(), torch.rsqrt(running_var + eps)
    return (x - mean) * invstd * gamma + beta, mean, invstd


def bn_rows_bwd(dy, x, mean, invstd, gamma, training):
    xh = (x - mean) * invstd
    s0, s1 = dy.sum(0), (dy * xh).sum(0)
    g = dy
    if training:
        g = dy - s0 / x.shape[0] - xh * s1 / x.shape[0]
    return g * gamma * invstd, s1, s0


# ----------------------------------------------------------------------------- recurrence (mirrors rnn.cu)
def rnn_use_bf16(H):
    return False


def rnn_plan(cell, H, B, bf16):
    return 8, (H + 7) // 8, 0, 0


def rnn_pack_weights(cell, w_hh_fwd, w_hh_rev, B, fwd=True, bwd=True):
    w = torch.stack([w_hh_fwd, w_hh_rev])
    return (w if fwd else None), (w if bwd else None)


def rnn_fwd(cell, gi, b_hh, wpack_fwd, lengths, T, B, H, want_sum=False):
    gates = 3 if cell == GRU else 4
    G = gates * H
    gi = gi.view(T, B, 2, G)
    hseq = torch.zeros(2, T + 2, B, H)
    cseq = torch.zeros(2, T + 2, B, H) if cell == LSTM else None
    saved = torch.zeros(2, T, B, 4, H)
    lens = lengths.long()
    for d in range(2):
        W = wpack_fwd[d]
        h = torch.zeros(B, H)
        c = torch.zeros(B, H)
        for s in range(T):
            t = s if d == 0 else T - 1 - s
            act = (t < lens).float()[:, None]
            acc = h @ W.t()                     # h is hseq of the previous step (zeros where inactive)
            g_in = gi[t, :, d]
            bh = b_hh[d]
            if cell == GRU:
                gr, gz, gn = (acc[:, :H] + bh[:H], acc[:, H:2 * H] + bh[H:2 * H], acc[:, 2 * H:] + bh[2 * H:])
                r = torch.sigmoid(g_in[:, :H] + gr)
                z = torch.sigmoid(g_in[:, H:2 * H] + gz)
                n = torch.tanh(g_in[:, 2 * H:] + r * gn)
                hn = (1 - z) * n + z * h
                sv = torch.stack([r, z, n, gn], 1)
                cn = c
            else:
                pre = g_in + acc + bh
                i_, f_, g_, o_ = (torch.sigmoid(pre[:, :H]), torch.sigmoid(pre[:, H:2 * H]),
                                  torch.tanh(pre[:, 2 * H:3 * H]), torch.sigmoid(pre[:, 3 * H:]))
                cn = f_ * c + i_ * g_
                hn = o_ * torch.tanh(cn)
                sv = torch.stack([i_, f_, g_, o_], 1)
            h = hn * act
            c = cn * act
            hseq[d, t + 1] = h
            if cell == LSTM:
                cseq[d, t + 1] = c
            saved[d, t] = sv * act[:, None]
    if want_sum:
        return hseq, cseq, saved, hseq[0, 1:T + 1] + hseq[1, 1:T + 1]
    return hseq, cseq, saved


def rnn_bwd(cell, dout, wpack_bwd, lengths, hseq, cseq, saved, T, B, H):
    gates = 3 if cell == GRU else 4
    G = gates * H
    dgi = torch.zeros(T, B, 2, G)
    dgh = torch.zeros(2, T, B, G)
    lens = lengths.long()
    for d in range(2):
        W = wpack_bwd[d]                         # [G, H]
        state_h = torch.zeros(B, H)
        state_c = torch.zeros(B, H)
        prev_dgh = torch.zeros(B, G)
        for s in range(T):
            t = T - 1 - s if d == 0 else s
            act = (t < lens)[:, None]
            acc = prev_dgh @ W                   # (dgates of the previously processed step) x W_hh
            carry = acc + state_h
            sv = saved[d, t]
            tprev_slot = t if d == 0 else t + 2
            dh = carry + dout[t]
            if cell == GRU:
                r, z, n, gn = sv[:, 0], sv[:, 1], sv[:, 2], sv[:, 3]
                hp = hseq[d, tprev_slot]
                dn = dh * (1 - z) * (1 - n * n)
                d2, e2 = dn, dn * r
                d1 = dh * (hp - n) * z * (1 - z)
                d0 = dn * gn * r * (1 - r)
                new_h = dh * z
                gi_g = torch.cat([d0, d1, d2], 1)
                gh_g = torch.cat([d0, d1, e2], 1)
                new_c = state_c
            else:
                i_, f_, g_, o_ = sv[:, 0], sv[:, 1], sv[:, 2], sv[:, 3]
                cp = cseq[d, tprev_slot]
                tcv = torch.tanh(cseq[d, t + 1])
                dc = state_c + dh * o_ * (1 - tcv * tcv)
                d0 = dc * g_ * i_ * (1 - i_)
                d1 = dc * cp * f_ * (1 - f_)
                d2 = dc * i_ * (1 - g_ * g_)
                d3 = dh * tcv * o_ * (1 - o_)
                gi_g = gh_g = torch.cat([d0, d1, d2, d3], 1)
                new_c = dc * f_
                new_h = torch.zeros_like(dh)
            state_h = torch.where(act, new_h, carry)
            state_c = torch.where(act, new_c, state_c)
            gi_g = torch.where(act, gi_g, torch.zeros_like(gi_g))
            gh_g = torch.where(act, gh_g, torch.zeros_like(gh_g))
            dgi[t, :, d] = gi_g
            dgh[d, t] = gh_g
            prev_dgh = gh_g
    R = T * B
    R4 = (R + 3) // 4 * 4
    dgiT = torch.zeros(2 * G, R4)
    dgiT[:, :R] = dgi.view(R, 2 * G).t()
    dghT = None
    if cell == GRU:
        dghT = torch.zeros(2 * G, R4)
        for d in range(2):
            dghT[d * G:(d + 1) * G, :R] = dgh[d].view(R, G).t()
    return dgi, dgiT, dghT


def row_sums(a, cols=None):
    return a[:, :cols].sum(1) if cols is not None else a.sum(1)


def rnn_sum_dirs(hseq, T, B, H):
    return hseq[0, 1:T + 1] + hseq[1, 1:T + 1]


# ----------------------------------------------------------------------------- softmax / CTC / spectrogram
def log_softmax_fwd(logits2d, C, want_lp=True, want_probs=False, want_argmax=False):
    x = logits2d[:, :C]
    lp = x.log_softmax(-1)
    return (lp if want_lp else None, lp.exp() if want_probs else None, x.argmax(-1) if want_argmax else None)


def log_softmax_bwd(g, lp):
    return g - lp.exp() * g.sum(-1, keepdim=True)


def ctc_fwd(log_probs, targets, input_lengths, target_lengths, max_target_len, blank=0):
    nll = F.ctc_loss(log_probs, targets, input_lengths, target_lengths, blank=blank, reduction="none")
    return nll.sum().view(1), nll, None


def ctc_bwd(log_probs, targets, input_lengths, target_lengths, alpha, nll, grad_scale, max_target_len, blank=0):
    with torch.enable_grad():
        lp = log_probs.detach().clone().requires_grad_(True)
        F.ctc_loss(lp, targets, input_lengths, target_lengths, blank=blank, reduction="sum").backward()
    return lp.grad * (grad_scale if grad_scale is not None else 1.0)


def greedy_collapse(idx, sizes, blank=0):
    N, T = idx.shape
    labels = torch.zeros(N, T, dtype=torch.int32)
    offsets = torch.zeros(N, T, dtype=torch.int32)
    counts = torch.zeros(N, dtype=torch.int32)
    for n in range(N):
        L = T if sizes is None else max(0, min(int(sizes[n]), T))
        k = 0
        for t in range(L):
            c = int(idx[n, t])
            if c != blank and (t == 0 or c != int(idx[n, t - 1])):
                labels[n, k], offsets[n, k] = c, t
                k += 1
        counts[n] = k
    return labels, offsets, counts


def adamw_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step, inv_scale=None):
    g = grad * (inv_scale if inv_scale is not None else 1.0)
    param.mul_(1 - lr * weight_decay)
    exp_avg.lerp_(g, 1 - beta1)
    exp_avg_sq.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    denom = (exp_avg_sq.sqrt() / (1 - beta2 ** step) ** 0.5).add_(eps)
    param.addcdiv_(exp_avg, denom, value=-lr / (1 - beta1 ** step))


def lookahead_fwd(x, w, context, act=None):
    T = x.shape[0]
    y = torch.zeros_like(x)
    for k in range(min(context, T)):
        y[:T - k] += w[:, k] * x[k:]
    return y.clamp(act[0], act[1]) if act is not None else y


def lookahead_bwd(dy, x, y, w, context, act=None, need_dx=True, need_dw=True):
    T = x.shape[0]
    g = dy if act is None else dy * ((y > act[0]) & (y < act[1])).to(dy.dtype)
    dx = torch.zeros_like(x)
    dw = torch.zeros_like(w)
    for k in range(min(context, T)):
        dx[k:] += w[:, k] * g[:T - k]
        dw[:, k] = (g[:T - k] * x[k:]).sum((0, 1))
    return (dx if need_dx else None), (dw if need_dw else None)


def dft_basis(n_fft, device):
    Fb = n_fft // 2 + 1
    n = torch.arange(n_fft, dtype=torch.float64)
    f = torch.arange(Fb, dtype=torch.float64)
    ang = 2 * torch.pi * ((f[:, None] * n[None, :]) % n_fft) / n_fft
    basis = torch.cat([torch.cos(ang), torch.sin(ang)]).float()
    return split3(basis, 1)


def spectrogram(wav, n_samples, window, basis, n_fft, hop, normalize=True):
    B, S = wav.shape
    Fb, Tmax = n_fft // 2 + 1, 1 + S // hop
    pad = F.pad(wav, (n_fft // 2, n_fft // 2))
    idx = torch.arange(n_fft)[None, :] + hop * torch.arange(Tmax)[:, None]
    spec = torch.zeros(B, 1, Fb, Tmax)
    for b in range(B):
        ns = int(n_samples[b])
        nf = 1 + ns // hop
        y = pad[b].clone()
        y[n_fft // 2 + ns:] = 0
        frames = y[idx] * window[None, :]
        reim = split3(frames.contiguous(), 0) @ basis.t()
        mag = torch.log1p(torch.sqrt(reim[:, :Fb] ** 2 + reim[:, Fb:] ** 2)).t()
        mag[:, nf:] = 0
        if normalize:
            v = mag[:, :nf]
            mag[:, :nf] = (v - v.mean()) / v.std()
        spec[b, 0] = mag
    return spec


def lengths_to_device(lengths, device):
    return torch.as_tensor(lengths, dtype=torch.int32).contiguous()


def set_debug_flags(flags):
    pass


_NAMES = [n for n, v in list(globals().items()) if callable(v) and not n.startswith("_") and n not in ("install",)]


def install(monkeypatch):
    """Route asr_b200.ops to this emulator for the duration of a test."""
    import asr_b200.ops as ops

    for name in _NAMES:
        if hasattr(ops, name):
            monkeypatch.setattr(ops, name, globals()[name])
    monkeypatch.setattr(ops, "require_cuda", lambda t, who: None)
    return ops
