"""Tensor-level wrappers over the C ABI (one function per entry point group).

PyTorch is used here only as the owner of device memory and of the CUDA stream: every function
checks that its tensors live on a CUDA device (there is no CPU implementation -- a CPU tensor raises),
allocates outputs / workspaces through the caching allocator and enqueues our kernels on the current
stream.  `LAUNCHES` counts kernel-launching calls (bench.py reports it as `gpu_launches`).
"""
from __future__ import annotations

import ctypes
import os

import torch

from . import _lib

GRU, LSTM = 0, 1
LAUNCHES = 0
BN_MOMENTUM = 0.1
BN_EPS = 1e-5


_LAST_DEV = None   # device of the tensors of the call being assembled (set by _p): the kernels go to ITS current stream


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream(_LAST_DEV).cuda_stream)


def _p(t):
    global _LAST_DEV
    if t is None:
        return None
    if t.is_cuda:
        _LAST_DEV = t.device
    return ctypes.c_void_p(t.data_ptr())


def _chk(*tensors, dtype=torch.float32):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("asr_b200 kernels need CUDA tensors (no CPU path exists)")
        if t.dtype != dtype:
            raise ValueError(f"expected {dtype}, got {t.dtype}")
        if not t.is_contiguous():
            raise ValueError("expected a contiguous tensor")


def require_cuda(t, who):
    """The one guard every public entry point goes through: no CPU implementation exists."""
    if not t.is_cuda:
        raise RuntimeError(f"asr_b200.{who}: CUDA tensor required (this package has no CPU or PyTorch fallback)")


PROFILE = None   # set to a dict to collect (start, end) CUDA-event pairs per entry point (bench.py)

# kernels launched per call of an entry point (memsets not counted); used for the `gpu_launches` claim
KERNELS_PER_CALL = {
    "asrb_conv2d_mask_bwd_weight": 3, "asrb_bn2d_stats": 2, "asrb_bn_act_mask_bwd": 3, "asrb_bn_rows_fwd": 3,
    "asrb_bn_rows_bwd": 3, "asrb_col_sums": 2, "asrb_ctc_fwd": 2, "asrb_ctc_bwd": 2, "asrb_spectrogram": 4, "asrb_rnn_pack_weights": 1, "asrb_rnn_fwd": 3, "asrb_rnn_bwd": 3, "asrb_rnn_fwd_sum": 4,   # (sentinel fill of the operand slabs + two passes of the recurrence, the second leaving at once: csrc/rnn3.cu)
    "asrb_conv32_bwd_weight": 2, "asrb_conv1_fwd": 3, "asrb_conv1_bwd_weight": 2,
}


def _call(name, *args):
    global LAUNCHES
    LAUNCHES += KERNELS_PER_CALL.get(name, 1)
    if _LAST_DEV is not None and _LAST_DEV.index != torch.cuda.current_device():
        # tensors on another device than the thread's current one: the library launches on cudaGetDevice() and builds
        # its TMA descriptors in that context, so make the tensors' device current for the call
        with torch.cuda.device(_LAST_DEV):
            _call_here(name, *args)
        return
    _call_here(name, *args)


def _call_here(name, *args):
    if PROFILE is None:
        _lib.call(name, *args, _stream())
        return
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record(torch.cuda.current_stream(_LAST_DEV))
    _lib.call(name, *args, _stream())
    end.record(torch.cuda.current_stream(_LAST_DEV))
    PROFILE.setdefault(name, []).append((start, end))


DEBUG_FLAGS = 0   # bit 0: CUDA-core GEMM, bit 1: CUDA-core recurrent product, bit 2: CUDA-core conv2 (tests only)


def set_debug_flags(flags: int):
    global DEBUG_FLAGS
    DEBUG_FLAGS = flags
    _lib.call("asrb_set_debug_flags", flags)


def lengths_to_device(lengths, device):
    """int32 lengths (CPU, as the reference keeps them) -> device int32 copy."""
    return torch.as_tensor(lengths, dtype=torch.int32).to(device, non_blocking=True).contiguous()


# ----------------------------------------------------------------------------- GEMM family
def _ld(t):
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError("expected a 2-D tensor with unit inner stride")
    return t.stride(0)


def gemm_tn(A, B, out=None, bias=None, accumulate=False):
    """out[M,N] (+)= A[M,K] @ B[N,K]^T (+ bias).  A, B, out may be row-strided views (unit inner stride)."""
    for t in (A, B, out, bias):
        if t is not None and (not t.is_cuda or t.dtype != torch.float32):
            raise RuntimeError("gemm_tn needs fp32 CUDA tensors")
    M, K = A.shape
    N, K2 = B.shape
    if K != K2:
        raise ValueError(f"inner dimensions differ: {K} vs {K2}")
    if out is None:
        out = torch.empty(M, N, device=A.device, dtype=torch.float32)
    _call("asrb_gemm_tn", _p(A), _ld(A), _p(B), _ld(B), _p(out), _ld(out), _p(bias), M, N, K, 1 if accumulate else 0)
    return out


def gemm_cta_limit(n):
    """cap the persistent CTAs of the GEMM launches that follow (0 = no cap); returns the previous cap"""
    return _lib.query("asrb_gemm_cta_limit", int(n))


def gemm_tn_bf16(A, B, out=None, bias=None, accumulate=False):
    """out[M,N] fp32 (+)= A[M,K] @ B[N,K]^T with bf16 operands (row strides multiples of 8 elements)."""
    for t in (A, B):
        if not t.is_cuda or t.dtype != torch.bfloat16:
            raise RuntimeError("gemm_tn_bf16 needs bf16 CUDA operands")
    M, K = A.shape
    N, K2 = B.shape
    if K != K2:
        raise ValueError(f"inner dimensions differ: {K} vs {K2}")
    if out is None:
        out = torch.empty(M, N, device=A.device, dtype=torch.float32)
    _call("asrb_gemm_tn_bf16", _p(A), _ld(A), _p(B), _ld(B), _p(out), _ld(out), _p(bias), M, N, K, 1 if accumulate else 0)
    return out


def transpose_bf16(x, out=None):
    """x [R, C] fp32 -> x^T in bf16 as a [C, R] view of a [C, R8] buffer (row stride multiple of 8 elements), or into
    `out` ([C, R] bf16 view whose row stride is a multiple of 8 elements)."""
    R, C = x.shape
    if out is not None:
        if out.dtype != torch.bfloat16 or tuple(out.shape) != (C, R) or out.stride(1) != 1 or out.stride(0) % 8:
            raise ValueError("transpose_bf16: out must be a [C, R] bf16 view with unit inner stride and a row stride % 8 == 0")
        _call("asrb_transpose_bf16", _p(x), R, C, _ld(x), _p(out), out.stride(0))
        return out
    R8 = (R + 7) // 8 * 8
    buf = torch.empty(C, R8, device=x.device, dtype=torch.bfloat16)
    _call("asrb_transpose_bf16", _p(x), R, C, _ld(x), _p(buf), R8)
    return buf[:, :R]


def transpose(x, out=None):
    """out[C,R] = x[R,C]^T (x may be a row-strided view)."""
    R, C = x.shape
    if out is None:
        out = torch.empty(C, R, device=x.device, dtype=torch.float32)
    _call("asrb_transpose", _p(x), R, C, _ld(x), _p(out), _ld(out))
    return out


def split3(x, mode):
    R, C = x.shape
    out = torch.empty(R, 3 * C, device=x.device, dtype=torch.float32)
    _call("asrb_split3", _p(x), R, C, _ld(x), _p(out), 3 * C, mode)
    return out


def gemm_tn_3x(A, B, out=None, bias=None):
    """~fp32-accurate product on TF32 tensor cores (3xTF32 operand expansion, K -> 3K)."""
    return gemm_tn(split3(A, 0), split3(B, 1), out=out, bias=bias)


def col_sums(a, cols=None):
    R = a.shape[0]
    cols = a.shape[1] if cols is None else cols
    out = torch.empty(cols, device=a.device, dtype=torch.float32)
    ws = torch.empty(_lib.query("asrb_rows_workspace_bytes", cols) // 4, device=a.device, dtype=torch.float32)
    _call("asrb_col_sums", _p(a), _ld(a), _p(out), _p(ws), ws.numel() * 4, R, cols)
    return out


# ----------------------------------------------------------------------------- conv / BN2d / layout
def conv_out_size(n, k, s, p):
    return (n + 2 * p - k) // s + 1


def conv2d_mask_fwd(x, w, bias, lengths, stride, padding):
    _chk(x, w, bias)
    _chk(lengths, dtype=torch.int32)
    B, Cin, Hin, Win = x.shape
    Cout, _, KH, KW = w.shape
    Hout, Wout = conv_out_size(Hin, KH, stride[0], padding[0]), conv_out_size(Win, KW, stride[1], padding[1])
    y = torch.empty(B, Cout, Hout, Wout, device=x.device, dtype=torch.float32)
    _call("asrb_conv2d_mask_fwd", _p(x), _p(w), _p(bias), _p(lengths), _p(y), B, Cin, Hin, Win, Cout, Hout, Wout,
          KH, KW, stride[0], stride[1], padding[0], padding[1])
    return y


def conv2d_mask_bwd_data(dy, w, lengths, x_shape, stride, padding):
    _chk(dy, w)
    B, Cin, Hin, Win = x_shape
    Cout, _, KH, KW = w.shape
    _, _, Hout, Wout = dy.shape
    dx = torch.empty(x_shape, device=dy.device, dtype=torch.float32)
    _call("asrb_conv2d_mask_bwd_data", _p(dy), _p(w), _p(lengths), _p(dx), B, Cin, Hin, Win, Cout, Hout, Wout, KH, KW,
          stride[0], stride[1], padding[0], padding[1])
    return dx


def conv2d_mask_bwd_weight(dy, x, lengths, w_shape, stride, padding, need_bias=True):
    _chk(dy, x)
    B, Cin, Hin, Win = x.shape
    Cout, _, KH, KW = w_shape
    _, _, Hout, Wout = dy.shape
    dw = torch.empty(w_shape, device=dy.device, dtype=torch.float32)
    db = torch.empty(Cout, device=dy.device, dtype=torch.float32) if need_bias else None
    nb = _lib.query("asrb_nchw_reduce_workspace_bytes", B, Cout, Hout, Wout)
    ws = torch.empty(nb // 8, device=dy.device, dtype=torch.float64)
    _call("asrb_conv2d_mask_bwd_weight", _p(dy), _p(x), _p(lengths), _p(dw), _p(db), _p(ws), nb, B, Cin, Hin, Win, Cout,
          Hout, Wout, KH, KW, stride[0], stride[1], padding[0], padding[1])
    return dw, db


def conv32_supported(w_shape, stride, padding):
    """True when the tcgen05 implicit-GEMM conv applies (32->32 channels, time stride 1)."""
    Cout, Cin, KH, KW = w_shape
    if DEBUG_FLAGS & 4:
        return False
    return bool(_lib.query("asrb_conv32_supported", Cin, Cout, KH, KW, stride[0], stride[1], padding[0], padding[1]))


def nchw_to_nhwc(x):
    _chk(x)
    B, C, H, W = x.shape
    out = torch.empty(B, H, W, C, device=x.device, dtype=torch.float32)
    _call("asrb_transpose_batched", _p(x), C, H * W, H * W, C * H * W, _p(out), C, H * W * C, B)
    return out


def pad_rows4(x):
    """NCHW tensor whose last dim is not a multiple of 4 -> (buffer [..., W4], W4) with zero-filled padding
    (TMA needs 16-byte row strides); returns (x, W) unchanged when already aligned."""
    W = x.shape[-1]
    if W % 4 == 0:
        return x, W
    W4 = (W + 3) // 4 * 4
    out = torch.empty(*x.shape[:-1], W4, device=x.device, dtype=torch.float32)
    rows = x.numel() // W
    _call("asrb_copy_rows_padded", _p(x), W, _p(out), W4, rows, W)
    return out, W4


def mask_time(x, lengths):
    """x[b, :, :, t >= lengths[b]] = 0 (a copy)."""
    return bn_act_mask_fwd(x, lengths, None, None, None, None, False, False, 0.0, 0.0)


def nchw_channel_sums(a, lengths):
    _chk(a)
    B, C, H, W = a.shape
    out = torch.empty(C, device=a.device, dtype=torch.float32)
    nb = _lib.query("asrb_nchw_reduce_workspace_bytes", B, C, H, W)
    ws = torch.empty(nb // 8, device=a.device, dtype=torch.float64)
    _call("asrb_nchw_channel_sums", _p(a), _p(lengths), _p(out), _p(ws), nb, B, C, H, W)
    return out


def conv32_pack_weights(w, fwd=True, dgrad=True):
    _chk(w)
    _, _, KH, KW = w.shape
    pf = torch.empty(KH * KW * 1024, device=w.device, dtype=torch.float32) if fwd else None
    pd = torch.empty(KH * KW * 1024, device=w.device, dtype=torch.float32) if dgrad else None
    _call("asrb_conv32_pack_weights", _p(w), _p(pf), _p(pd), KH, KW)
    return pf, pd


# output rows per work item of the 32->32 conv's forward / data-gradient kernels (1 = the one-row kernel, 2, 4)
CONV_ROWS = int(os.environ.get("ASRB_CONV_ROWS", "4"))

# Step hand-over of the tensor-memory recurrent kernels (asrb_debug_rnn_dbg bits, csrc/rnn.cu): set once at import when the
# environment asks for something else than the library's default (A/B timing of whole steps through bench.py).
if os.environ.get("ASRB_RNN_DBG"):
    _lib.query("asrb_debug_rnn_dbg", int(os.environ["ASRB_RNN_DBG"]))


def conv32_pack_rows(pack, w_shape, stride_h, mode, rows=None):
    """plain pack of conv32_pack_weights (mode 0: forward, 1: data gradient) -> tap matrices stacked for `rows` output rows"""
    rows = CONV_ROWS if rows is None else rows
    _chk(pack)
    _, _, KH, KW = w_shape
    out = torch.empty((KH + stride_h * (rows - 1)) * KW * rows * 1024, device=pack.device, dtype=torch.float32)
    _call("asrb_conv32_pack_rows", _p(pack), _p(out), KH, KW, stride_h, rows, mode)
    return out


def conv32_fwd(x_nhwc, pack_fwd, bias, lengths, w_shape, stride, padding, rows=0):
    """rows = 0: pack_fwd is the plain pack; rows = 2 / 4: pack_fwd comes from conv32_pack_rows(..., mode 0, rows)"""
    _chk(x_nhwc, pack_fwd, bias)
    B, Hin, Win, _ = x_nhwc.shape
    _, _, KH, KW = w_shape
    Hout, Wout = conv_out_size(Hin, KH, stride[0], padding[0]), conv_out_size(Win, KW, 1, padding[1])
    y = torch.empty(B, 32, Hout, Wout, device=x_nhwc.device, dtype=torch.float32)
    if rows > 1:
        _call("asrb_conv32_fwd_rows", _p(x_nhwc), _p(pack_fwd), _p(bias), _p(lengths), _p(y), B, Hin, Win, Hout, Wout, KH, KW,
              stride[0], padding[0], padding[1], rows)
    else:
        _call("asrb_conv32_fwd", _p(x_nhwc), _p(pack_fwd), _p(bias), _p(lengths), _p(y), B, Hin, Win, Hout, Wout, KH, KW,
              stride[0], padding[0], padding[1])
    return y


def conv32_bwd_data(dy_nhwc, pack_dgrad, x_shape, w_shape, stride, padding, rows=0):
    _chk(dy_nhwc, pack_dgrad)
    B, _, Hin, Win = x_shape
    _, Hout, Wout, _ = dy_nhwc.shape
    _, _, KH, KW = w_shape
    dx = torch.empty(x_shape, device=dy_nhwc.device, dtype=torch.float32)
    if rows > 1:
        _call("asrb_conv32_bwd_data_rows", _p(dy_nhwc), _p(pack_dgrad), _p(dx), B, Hin, Win, Hout, Wout, KH, KW, stride[0],
              padding[0], padding[1], rows)
    else:
        _call("asrb_conv32_bwd_data", _p(dy_nhwc), _p(pack_dgrad), _p(dx), B, Hin, Win, Hout, Wout, KH, KW, stride[0],
              padding[0], padding[1])
    return dx


def conv32_bwd_weight(x, dy_masked, w_shape, stride, padding):
    """x, dy NCHW (dy already masked) -> dw"""
    _chk(x, dy_masked)
    B, _, Hin, Win = x.shape
    _, _, Hout, Wout = dy_masked.shape
    _, _, KH, KW = w_shape
    if _lib.query("asrb_debug_conv_wgrad_bf16", -1):
        dyp, lddy = dy_masked, Wout                # converted to bf16 (and padded) inside the call
    else:
        dyp, lddy = pad_rows4(dy_masked)
    dw = torch.empty(w_shape, device=x.device, dtype=torch.float32)
    nb = _lib.query("asrb_conv32_bwd_weight_workspace_bytes", B, Hin, Win, Hout, Wout)
    ws = torch.empty(nb // 4, device=x.device, dtype=torch.float32)
    _call("asrb_conv32_bwd_weight", _p(x), _p(dyp), lddy, _p(dw), _p(ws), nb, B, Hin, Win, Hout, Wout, KH, KW,
          stride[0], padding[0], padding[1])
    return dw


def conv1_supported(x_shape, w_shape, stride, padding):
    """True when the tcgen05 polyphase conv applies (1->32 channels, (KH,11) kernel, stride (2,2), padding (even,5))."""
    Cout, Cin, KH, KW = w_shape
    if DEBUG_FLAGS & 4:
        return False
    return bool(_lib.query("asrb_conv1_supported", Cin, Cout, x_shape[2], KH, KW, stride[0], stride[1], padding[0], padding[1]))


def conv1_fwd(x, w, bias, lengths, stride, padding):
    _chk(x, w, bias)
    B, _, F, T = x.shape
    Cout, _, KH, KW = w.shape
    Hout, Wout = conv_out_size(F, KH, 2, padding[0]), conv_out_size(T, KW, 2, padding[1])
    y = torch.empty(B, Cout, Hout, Wout, device=x.device, dtype=torch.float32)
    nb = _lib.query("asrb_conv1_workspace_bytes", B, F, T, 0)
    ws = torch.empty(nb // 4, device=x.device, dtype=torch.float32)
    _call("asrb_conv1_fwd", _p(x), _p(w), _p(bias), _p(lengths), _p(y), _p(ws), nb, B, F, T, Hout, Wout, KH, padding[0])
    return y


def conv1_bwd_weight(x, dy_masked, w_shape, padding):
    _chk(x, dy_masked)
    B, _, F, T = x.shape
    _, _, Hout, Wout = dy_masked.shape
    _, _, KH, KW = w_shape
    dyp, lddy = pad_rows4(dy_masked)
    dw = torch.empty(w_shape, device=x.device, dtype=torch.float32)
    nb = _lib.query("asrb_conv1_workspace_bytes", B, F, T, 1)
    ws = torch.empty(nb // 4, device=x.device, dtype=torch.float32)
    _call("asrb_conv1_bwd_weight", _p(x), _p(dyp), lddy, _p(dw), _p(ws), nb, B, F, T, Hout, Wout, KH, padding[0])
    return dw


def bn2d_stats(y, running_mean, running_var, training, momentum=BN_MOMENTUM, eps=BN_EPS):
    """-> (mean[C], invstd[C]); training: batch statistics (+ running-stat update), eval: from running stats."""
    _chk(y)
    B, C, H, W = y.shape
    mean = torch.empty(C, device=y.device, dtype=torch.float32)
    invstd = torch.empty(C, device=y.device, dtype=torch.float32)
    if training:
        nb = _lib.query("asrb_nchw_reduce_workspace_bytes", B, C, H, W)
        ws = torch.empty(nb // 8, device=y.device, dtype=torch.float64)
        _call("asrb_bn2d_stats", _p(y), _p(mean), _p(invstd), _p(running_mean), _p(running_var), momentum, eps, _p(ws),
              nb, B, C, H, W)
    else:
        _call("asrb_bn_eval_stats", _p(running_mean), _p(running_var), eps, C, _p(mean), _p(invstd))
    return mean, invstd


def bn_act_mask_fwd(y, lengths, mean, invstd, gamma, beta, has_bn, has_act, lo, hi):
    _chk(y, mean, invstd, gamma, beta)
    B, C, H, W = y.shape
    z = torch.empty_like(y)
    _call("asrb_bn_act_mask_fwd", _p(y), _p(lengths), _p(mean), _p(invstd), _p(gamma), _p(beta), int(has_bn),
          int(has_act), float(lo), float(hi), _p(z), B, C, H, W)
    return z


def bn_act_mask_bwd(dz, y, lengths, mean, invstd, gamma, beta, has_bn, has_act, lo, hi, training):
    _chk(dz, y)
    B, C, H, W = y.shape
    dy = torch.empty_like(y)
    dgamma = dbeta = ws = None
    nb = 0
    if has_bn:
        dgamma = torch.empty(C, device=y.device, dtype=torch.float32)
        dbeta = torch.empty(C, device=y.device, dtype=torch.float32)
        nb = _lib.query("asrb_nchw_reduce_workspace_bytes", B, C, H, W)
        ws = torch.empty(nb // 8, device=y.device, dtype=torch.float64)
    _call("asrb_bn_act_mask_bwd", _p(dz), _p(y), _p(lengths), _p(mean), _p(invstd), _p(gamma), _p(beta), int(has_bn),
          int(has_act), float(lo), float(hi), int(training), _p(dy), _p(dgamma), _p(dbeta), _p(ws), nb, B, C, H, W)
    return dy, dgamma, dbeta


def nchw_to_tnf(x):
    """[B,C,D,T] -> [T,B,C*D] (deepspeech.py:135-137)."""
    _chk(x)
    B, C, D, T = x.shape
    out = torch.empty(T, B, C * D, device=x.device, dtype=torch.float32)
    _call("asrb_transpose_batched", _p(x), C * D, T, T, C * D * T, _p(out), B * C * D, C * D, B)
    return out


def tnf_to_nchw(x, C, D):
    _chk(x)
    T, B, F = x.shape
    out = torch.empty(B, C, D, T, device=x.device, dtype=torch.float32)
    _call("asrb_transpose_batched", _p(x), T, F, B * F, F, _p(out), T, F * T, B)
    return out


# ----------------------------------------------------------------------------- BatchNorm1d over rows
def _rows_ws(cols, device):
    return torch.empty(_lib.query("asrb_rows_workspace_bytes", cols) // 4, device=device, dtype=torch.float32)


def bn_rows_fwd(x, gamma, beta, running_mean, running_var, training, momentum=BN_MOMENTUM, eps=BN_EPS):
    """x [R, cols] -> (y, mean, invstd)"""
    _chk(x, gamma, beta, running_mean, running_var)
    R, cols = x.shape
    mean = torch.empty(cols, device=x.device, dtype=torch.float32)
    invstd = torch.empty(cols, device=x.device, dtype=torch.float32)
    y = torch.empty_like(x)
    ws = _rows_ws(cols, x.device)
    _call("asrb_bn_rows_fwd", _p(x), _p(gamma), _p(beta), _p(running_mean), _p(running_var), int(training), momentum,
          eps, _p(mean), _p(invstd), _p(y), _p(ws), ws.numel() * 4, R, cols)
    return y, mean, invstd


def bn_rows_bwd(dy, x, mean, invstd, gamma, training):
    _chk(dy, x, mean, invstd, gamma)
    R, cols = x.shape
    dx = torch.empty_like(x)
    dgamma = torch.empty(cols, device=x.device, dtype=torch.float32)
    dbeta = torch.empty(cols, device=x.device, dtype=torch.float32)
    ws = _rows_ws(cols, x.device)
    _call("asrb_bn_rows_bwd", _p(dy), _p(x), _p(mean), _p(invstd), _p(gamma), int(training), _p(dx), _p(dgamma),
          _p(dbeta), _p(ws), ws.numel() * 4, R, cols)
    return dx, dgamma, dbeta


# ----------------------------------------------------------------------------- recurrent layers
RNN_BF16 = True   # operand precision of the recurrent product: bf16 (default) or tf32; see include/asr_b200.h


RNN_BF16_MIN_HIDDEN = 256   # below this the tf32 state fits in flight anyway: keep the extra mantissa bits


def rnn_use_bf16(H):
    """bf16 recurrent operands need 16-byte rows of bf16 (H % 8 == 0); the CUDA-core debug path reads fp32."""
    return bool(RNN_BF16 and H % 8 == 0 and H >= RNN_BF16_MIN_HIDDEN and not (DEBUG_FLAGS & 2))


def rnn_plan(cell, H, B, bf16):
    nj, P = ctypes.c_int(), ctypes.c_int()
    wf, wb = ctypes.c_size_t(), ctypes.c_size_t()
    _lib.call("asrb_rnn_plan", cell, H, B, int(bf16), ctypes.byref(nj), ctypes.byref(P), ctypes.byref(wf), ctypes.byref(wb))
    return nj.value, P.value, wf.value, wb.value


def rnn_pack_weights(cell, w_hh_fwd, w_hh_rev, B, fwd=True, bwd=True):
    _chk(w_hh_fwd, w_hh_rev)
    H = w_hh_fwd.shape[1]
    bf16 = rnn_use_bf16(H)
    _, _, wf, wb = rnn_plan(cell, H, B, bf16)
    pf = torch.empty(wf, device=w_hh_fwd.device, dtype=torch.uint8) if fwd else None
    pb = torch.empty(wb, device=w_hh_fwd.device, dtype=torch.uint8) if bwd else None
    _call("asrb_rnn_pack_weights", cell, H, B, int(bf16), _p(w_hh_fwd), _p(w_hh_rev), _p(pf), _p(pb))
    return pf, pb


def rnn_fwd(cell, gi, b_hh, wpack_fwd, lengths, T, B, H, want_sum=False):
    """gi [T,B,2,G], b_hh [2,G] -> hseq [2,T+2,B,H], cseq (LSTM) or None, saved [2,T,B,4,H]
    (+ out [T,B,H] = the sum of the two directions, blocks.py:92, when want_sum)"""
    _chk(gi, b_hh)
    _chk(lengths, dtype=torch.int32)
    dev = gi.device
    bf16 = rnn_use_bf16(H)
    hseq = torch.empty(2, T + 2, B, H, device=dev, dtype=torch.float32)
    hbf = torch.empty(2, T + 2, B, (H + 63) // 64 * 64, device=dev, dtype=torch.bfloat16) if bf16 else None
    cseq = torch.empty(2, T + 2, B, H, device=dev, dtype=torch.float32) if cell == LSTM else None
    saved = torch.empty(_lib.query("asrb_rnn_saved_floats", cell, H, B, int(bf16), T), device=dev, dtype=torch.float32)
    counters = torch.empty(128, device=dev, dtype=torch.int32)
    if want_sum:
        out = torch.empty(T, B, H, device=dev, dtype=torch.float32)
        _call("asrb_rnn_fwd_sum", cell, int(bf16), _p(gi), _p(b_hh), _p(wpack_fwd), _p(lengths), _p(hseq), _p(hbf), _p(cseq),
              _p(saved), _p(out), _p(counters), T, B, H)
        return hseq, cseq, saved, out
    _call("asrb_rnn_fwd", cell, int(bf16), _p(gi), _p(b_hh), _p(wpack_fwd), _p(lengths), _p(hseq), _p(hbf), _p(cseq),
          _p(saved), _p(counters), T, B, H)
    return hseq, cseq, saved


def rnn_bwd(cell, dout, wpack_bwd, lengths, hseq, cseq, saved, T, B, H):
    """dout [T,B,H] -> (dgi [T,B,2,G], dgiT [2G, R4], dghT [2G, R4] | None): the gate gradients row-major (operand of
    the input-gradient GEMM) and transposed (operands of the weight-gradient GEMMs; R4 = T*B rounded up to 4).
    dghT (hidden-side gradients) exists for GRU only; for LSTM it equals dgiT."""
    _chk(dout, hseq, cseq, saved)
    G = (3 if cell == GRU else 4) * H
    dev = dout.device
    bf16 = rnn_use_bf16(H)
    gdt = torch.bfloat16 if bf16 else torch.float32   # gate gradients feed the backward GEMMs in the mode's operand type
    R4 = (T * B + 7) // 8 * 8 if bf16 else (T * B + 3) // 4 * 4
    dgi = torch.empty(T, B, 2, G, device=dev, dtype=gdt)
    dgh = None if bf16 else torch.empty(2, T, B, G, device=dev, dtype=torch.float32)
    dghbf = torch.empty(2, T, B, (G + 63) // 64 * 64, device=dev, dtype=torch.bfloat16) if bf16 else None
    dgiT = torch.empty(2 * G, R4, device=dev, dtype=gdt)
    dghT = torch.empty(2 * G, R4, device=dev, dtype=gdt) if cell == GRU else None
    counters = torch.empty(128, device=dev, dtype=torch.int32)
    _call("asrb_rnn_bwd", cell, int(bf16), _p(dout), _p(wpack_bwd), _p(lengths), _p(hseq), _p(cseq), _p(saved), _p(dgi),
          _p(dgh), _p(dghbf), _p(dgiT), _p(dghT), R4, _p(counters), T, B, H)
    return dgi, dgiT, dghT


def row_sums(a, cols=None):
    """out[r] = sum_c a[r, :cols] for a row-strided 2-D view."""
    rows = a.shape[0]
    cols = a.shape[1] if cols is None else cols
    out = torch.empty(rows, device=a.device, dtype=torch.float32)
    _call("asrb_row_sums_bf16" if a.dtype == torch.bfloat16 else "asrb_row_sums", _p(a), _ld(a), _p(out), rows, cols)
    return out


def rnn_sum_dirs(hseq, T, B, H):
    out = torch.empty(T, B, H, device=hseq.device, dtype=torch.float32)
    _call("asrb_rnn_sum_dirs", _p(hseq), _p(out), T, B, H)
    return out


# ----------------------------------------------------------------------------- softmax / CTC
def log_softmax_fwd(logits2d, C, want_lp=True, want_probs=False, want_argmax=False):
    """logits2d [R, ld>=C] -> (log_probs [R,C] | None, probs | None, argmax int64 [R] | None)"""
    R = logits2d.shape[0]
    dev = logits2d.device
    lp = torch.empty(R, C, device=dev, dtype=torch.float32) if want_lp else None
    pr = torch.empty(R, C, device=dev, dtype=torch.float32) if want_probs else None
    am = torch.empty(R, device=dev, dtype=torch.int64) if want_argmax else None
    _call("asrb_log_softmax_fwd", _p(logits2d), _ld(logits2d), _p(lp), _p(pr), _p(am), R, C)
    return lp, pr, am


def log_softmax_bwd(g, lp):
    _chk(g, lp)
    R, C = lp.shape
    d = torch.empty(R, C, device=lp.device, dtype=torch.float32)
    _call("asrb_log_softmax_bwd", _p(g), _p(lp), _p(d), C, R, C)
    return d


def ctc_fwd(log_probs, targets, input_lengths, target_lengths, max_target_len, blank=0):
    """log_probs [T,N,C]; int32 device tensors for the rest -> (loss[1], nll[N], alpha workspace)"""
    _chk(log_probs)
    _chk(targets, input_lengths, target_lengths, dtype=torch.int32)
    T, N, C = log_probs.shape
    nb = _lib.query("asrb_ctc_workspace_bytes", T, N, max_target_len)
    alpha = torch.empty(nb // 4, device=log_probs.device, dtype=torch.float32)
    nll = torch.empty(N, device=log_probs.device, dtype=torch.float32)
    loss = torch.empty(1, device=log_probs.device, dtype=torch.float32)
    _call("asrb_ctc_fwd", _p(log_probs), _p(targets), _p(input_lengths), _p(target_lengths), _p(alpha), nb, _p(nll),
          _p(loss), T, N, C, max_target_len, blank)
    return loss, nll, alpha


def ctc_bwd(log_probs, targets, input_lengths, target_lengths, alpha, nll, grad_scale, max_target_len, blank=0):
    T, N, C = log_probs.shape
    grad = torch.empty_like(log_probs)
    _call("asrb_ctc_bwd", _p(log_probs), _p(targets), _p(input_lengths), _p(target_lengths), _p(alpha), _p(nll),
          _p(grad_scale), _p(grad), T, N, C, max_target_len, blank)
    return grad


def greedy_collapse(idx, sizes, blank=0):
    """idx int64 [N,T] (frame-wise argmax), sizes int32 [N] on the device or None -> (labels [N,T], offsets [N,T], counts [N])"""
    _chk(idx, dtype=torch.int64)
    N, T = idx.shape
    labels = torch.empty(N, T, device=idx.device, dtype=torch.int32)
    offsets = torch.empty(N, T, device=idx.device, dtype=torch.int32)
    counts = torch.empty(N, device=idx.device, dtype=torch.int32)
    _call("asrb_greedy_collapse", _p(idx), _p(sizes), N, T, int(blank), _p(labels), _p(offsets), _p(counts))
    return labels, offsets, counts


# ----------------------------------------------------------------------------- optimizer
def adamw_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step, inv_scale=None):
    """In-place torch.optim.AdamW step on contiguous fp32 CUDA tensors of equal size (any shape)."""
    for t in (param, grad, exp_avg, exp_avg_sq):
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError("adamw_step needs contiguous fp32 CUDA tensors")
    if not (param.numel() == grad.numel() == exp_avg.numel() == exp_avg_sq.numel()):
        raise ValueError("adamw_step: size mismatch")
    _call("asrb_adamw_step", _p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), param.numel(), float(lr), float(beta1),
          float(beta2), float(eps), float(weight_decay), int(step), _p(inv_scale))


# ----------------------------------------------------------------------------- lookahead convolution
def lookahead_fwd(x, w, context, act=None):
    """x [T,N,H], w [H,context] -> y[t,n,f] = sum_k w[f,k] x[t+k,n,f] (zero past T), optionally clamped to act=(lo,hi)"""
    _chk(x, w)
    T, N, H = x.shape
    y = torch.empty_like(x)
    lo, hi = act if act is not None else (0.0, 0.0)
    _call("asrb_lookahead_fwd", _p(x), _p(w), _p(y), T, N, H, context, int(act is not None), float(lo), float(hi))
    return y


def lookahead_bwd(dy, x, y, w, context, act=None, need_dx=True, need_dw=True):
    _chk(dy, x, y, w)
    T, N, H = x.shape
    dx = torch.empty_like(x) if need_dx else None
    dw = torch.empty(H, context, device=x.device, dtype=torch.float32) if need_dw else None
    lo, hi = act if act is not None else (0.0, 0.0)
    _call("asrb_lookahead_bwd", _p(dy), _p(x), _p(y), _p(w), _p(dx), _p(dw), T, N, H, context, int(act is not None),
          float(lo), float(hi))
    return dx, dw


# ----------------------------------------------------------------------------- spectrogram
def dft_basis(n_fft, device):
    basis = torch.empty(2 * (n_fft // 2 + 1), 3 * n_fft, device=device, dtype=torch.float32)
    _call("asrb_dft_basis", _p(basis), n_fft)
    return basis


def spectrogram(wav, n_samples, window, basis, n_fft, hop, normalize=True):
    """wav [B, S] zero padded, n_samples int32[B] -> [B,1,n_fft/2+1, 1+S//hop]"""
    _chk(wav, window, basis)
    _chk(n_samples, dtype=torch.int32)
    B, S = wav.shape
    F, Tmax = n_fft // 2 + 1, 1 + S // hop
    nb = _lib.query("asrb_spectrogram_workspace_bytes", B, S, n_fft, hop)
    ws = torch.empty((nb + 3) // 4, device=wav.device, dtype=torch.float32)
    spec = torch.empty(B, 1, F, Tmax, device=wav.device, dtype=torch.float32)
    _call("asrb_spectrogram", _p(wav), S, _p(n_samples), _p(window), _p(basis), _p(spec), int(normalize), _p(ws), nb, B,
          S, n_fft, hop)
    return spec
