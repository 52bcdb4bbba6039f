"""GPU parity tests, kernel by kernel, through the C ABI (asr_b200.ops -> libasr_b200.so) against the CPU oracle
(oracle/explicit.py, oracle/torch_path.py) on the same seeded inputs.

Tolerances (stated per test): integer / index outputs bit-exact; fp32 CUDA-core kernels ~1e-5 relative;
tensor-core products carry TF32 operand rounding (10-bit mantissa, fp32 accumulate): |err| <= 2e-3 * sqrt(K) * rms(a)*rms(b).
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import explicit, torch_path

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    from asr_b200 import ops as _ops

    _ops.set_debug_flags(0)
    yield _ops
    _ops.set_debug_flags(0)


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def report(name, got, ref):
    err = (got.double().cpu() - ref.double()).abs()
    print(f"[{name}] max_abs_err={err.max().item():.3e} ref_max={ref.abs().max().item():.3e} "
          f"argmax={np.unravel_index(int(err.argmax()), err.shape)}")
    return err.max().item()


# ----------------------------------------------------------------------------- GEMM
GEMM_SHAPES = [(128, 128, 32), (128, 64, 64), (256, 256, 256), (300, 200, 100), (1000, 29, 800), (77, 2400, 800),
               (2048, 4800, 1312), (36, 40, 32064)]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_tn_tcgen05(ops, M, N, K):
    A, B, bias = rnd(M, K, seed=1), rnd(N, K, seed=2), rnd(N, seed=3)
    ref = A.double() @ B.double().t() + bias.double()
    out = ops.gemm_tn(A.to(DEV), B.to(DEV), bias=bias.to(DEV))
    torch.cuda.synchronize()
    err = report(f"gemm_tc {M}x{N}x{K}", out, ref)
    assert err <= 2e-3 * math.sqrt(K)
    # accumulate flag: C += A B^T
    out2 = ops.gemm_tn(A.to(DEV), B.to(DEV), out=out.clone(), accumulate=True)
    assert report("gemm_tc accumulate", out2, 2 * ref - bias.double()) <= 4e-3 * math.sqrt(K)


def test_gemm_tn_strided_views_and_3x(ops):
    M, N, K = 500, 96, 160
    Abuf, Bbuf = rnd(M, K + 8, seed=4).to(DEV), rnd(N, K + 4, seed=5).to(DEV)
    Cbuf = torch.zeros(M, N + 12, device=DEV)
    A, B = Abuf[:, :K], Bbuf[:, :K]
    ref = A.double().cpu() @ B.double().cpu().t()
    ops.gemm_tn(A, B, out=Cbuf[:, 4:4 + N])
    assert report("gemm_tc strided", Cbuf[:, 4:4 + N], ref) <= 2e-3 * math.sqrt(K)
    assert Cbuf[:, :4].abs().max() == 0 and Cbuf[:, 4 + N:].abs().max() == 0
    out3 = ops.gemm_tn_3x(A.contiguous(), B.contiguous())
    # tensor-core fp32 accumulation keeps fewer guard bits than an FMA chain: ~1e-6 relative per output
    assert report("gemm 3xTF32", out3, ref) <= 2e-5 * math.sqrt(K)


@pytest.mark.parametrize("M,N,K", [(300, 200, 96), (1000, 2400, 800), (129, 36, 40)])
def test_gemm_tma_store_epilogue_equals_direct_stores(ops, M, N, K):
    """the two epilogues (shared-memory staging + TMA store / reduce-add, and direct row-per-lane stores) write the
    same fp32 values: bit-equal, with bias, in accumulate mode, and clipped at ragged M and N"""
    from asr_b200 import _lib

    A, B, bias = rnd(M, K, seed=31).to(DEV), rnd(N, K, seed=32).to(DEV), rnd(N, seed=33).to(DEV)
    C0 = rnd(M, N + 4, seed=34).to(DEV)
    res = {}
    try:
        for mode in (1, 0):
            _lib.query("asrb_debug_gemm_tma_store", mode)
            out = ops.gemm_tn(A, B, bias=bias)
            buf = C0.clone()
            ops.gemm_tn(A, B, out=buf[:, :N], accumulate=True)
            bf = ops.gemm_tn_bf16(A.bfloat16(), B.bfloat16(), out=None)
            torch.cuda.synchronize()
            res[mode] = (out, buf, bf)
    finally:
        _lib.query("asrb_debug_gemm_tma_store", 1)
    for a, b in zip(res[1], res[0]):
        assert torch.equal(a, b)
    assert torch.equal(res[1][1][:, N:], C0[:, N:])           # columns past N untouched


def test_gemm_simt_debug_path_and_transpose(ops):
    M, N, K = 200, 130, 70
    A, B = rnd(M, K, seed=6), rnd(N, K, seed=7)
    ops.set_debug_flags(1)
    try:
        out = ops.gemm_tn(A.to(DEV), B.to(DEV))
    finally:
        ops.set_debug_flags(0)
    assert report("gemm_simt", out, A.double() @ B.double().t()) <= 1e-4
    t = ops.transpose(A.to(DEV))
    assert torch.equal(t.cpu(), A.t().contiguous())


# ----------------------------------------------------------------------------- MaskConv pieces
CONV_CASES = [
    dict(B=2, Cin=1, H=8, W=10, Cout=2, k=(3, 3), s=(1, 1), p=(1, 1), lens=[10, 4]),
    dict(B=3, Cin=1, H=161, W=61, Cout=32, k=(41, 11), s=(2, 2), p=(20, 5), lens=[31, 24, 15]),
    dict(B=2, Cin=32, H=81, W=150, Cout=32, k=(21, 11), s=(2, 1), p=(10, 5), lens=[150, 77]),
    dict(B=2, Cin=3, H=9, W=140, Cout=5, k=(3, 5), s=(1, 2), p=(1, 2), lens=[68, 30]),
]


@pytest.mark.parametrize("c", CONV_CASES)
def test_conv2d_mask_fwd_bwd(ops, c):
    x = rnd(c["B"], c["Cin"], c["H"], c["W"], seed=10).requires_grad_(True)
    w = (rnd(c["Cout"], c["Cin"], *c["k"], seed=11) * 0.1).requires_grad_(True)
    b = rnd(c["Cout"], seed=12).requires_grad_(True)
    lens = torch.tensor(c["lens"], dtype=torch.int32)
    y_ref = explicit.time_mask(F.conv2d(x.double(), w.double(), b.double(), stride=c["s"], padding=c["p"]), lens.long())
    dy = rnd(*y_ref.shape, seed=13)
    y_ref.backward(dy.double())
    ld = lens.to(DEV)
    y = ops.conv2d_mask_fwd(x.detach().to(DEV), w.detach().to(DEV), b.detach().to(DEV), ld, c["s"], c["p"])
    scale = y_ref.abs().max().item()
    assert report("conv fwd", y, y_ref.detach()) <= 2e-5 * scale
    dx = ops.conv2d_mask_bwd_data(dy.to(DEV), w.detach().to(DEV), ld, tuple(x.shape), c["s"], c["p"])
    assert report("conv bwd_data", dx, x.grad) <= 2e-5 * x.grad.abs().max().item()
    dw, db = ops.conv2d_mask_bwd_weight(dy.to(DEV), x.detach().to(DEV), ld, tuple(w.shape), c["s"], c["p"])
    assert report("conv bwd_weight", dw, w.grad) <= 5e-5 * w.grad.abs().max().item()
    assert report("conv bwd_bias", db, b.grad) <= 5e-5 * b.grad.abs().max().item()


CONV32_CASES = [
    dict(B=2, H=81, W=150, k=(21, 11), s=(2, 1), p=(10, 5), lens=[150, 77]),     # the DeepSpeech2 conv2 geometry
    dict(B=1, H=81, W=501, k=(21, 11), s=(2, 1), p=(10, 5), lens=[501]),         # 2 tile pairs, ragged tail, W % 4 != 0
    dict(B=3, H=20, W=31, k=(5, 3), s=(1, 1), p=(2, 1), lens=[31, 20, 7]),
    dict(B=2, H=17, W=260, k=(3, 7), s=(3, 1), p=(0, 3), lens=[260, 129]),
]


@pytest.mark.parametrize("c", CONV32_CASES, ids=lambda c: f"H{c['H']}_W{c['W']}_k{c['k'][0]}x{c['k'][1]}")
def test_conv32_tensor_core_fwd_bwd(ops, c):
    """32->32 channel conv as a tcgen05 implicit GEMM (TF32 operands) vs the fp64 oracle."""
    x = rnd(c["B"], 32, c["H"], c["W"], seed=14).requires_grad_(True)
    w = (rnd(32, 32, *c["k"], seed=15) * 0.05).requires_grad_(True)
    b = rnd(32, seed=16).requires_grad_(True)
    lens = torch.tensor(c["lens"], dtype=torch.int32)
    assert ops.conv32_supported(tuple(w.shape), c["s"], c["p"])
    y_ref = explicit.time_mask(F.conv2d(x.double(), w.double(), b.double(), stride=c["s"], padding=c["p"]), lens.long())
    dy = explicit.time_mask(rnd(*y_ref.shape, seed=17), lens.long())
    y_ref.backward(dy.double())
    K = 32 * c["k"][0] * c["k"][1]
    ld = lens.to(DEV)
    xd, wd = x.detach().to(DEV), w.detach().to(DEV)
    pf, pd = ops.conv32_pack_weights(wd)
    y = ops.conv32_fwd(ops.nchw_to_nhwc(xd), pf, b.detach().to(DEV), ld, tuple(w.shape), c["s"], c["p"])
    torch.cuda.synchronize()
    assert report("conv32 fwd", y, y_ref.detach()) <= 2e-3 * math.sqrt(K) * 0.05
    dx = ops.conv32_bwd_data(ops.nchw_to_nhwc(dy.to(DEV)), pd, tuple(x.shape), tuple(w.shape), c["s"], c["p"])
    torch.cuda.synchronize()
    assert report("conv32 dgrad", dx, x.grad) <= 2e-3 * math.sqrt(K) * 0.05
    # row-grouped kernels (2 / 4 output rows per work item share every source strip): same products, other summation order
    for rows in (2, 4):
        pfr = ops.conv32_pack_rows(pf, tuple(w.shape), c["s"][0], 0, rows)
        pdr = ops.conv32_pack_rows(pd, tuple(w.shape), c["s"][0], 1, rows)
        yr = ops.conv32_fwd(ops.nchw_to_nhwc(xd), pfr, b.detach().to(DEV), ld, tuple(w.shape), c["s"], c["p"], rows=rows)
        dxr = ops.conv32_bwd_data(ops.nchw_to_nhwc(dy.to(DEV)), pdr, tuple(x.shape), tuple(w.shape), c["s"], c["p"], rows=rows)
        torch.cuda.synchronize()
        assert report(f"conv32 fwd rows={rows}", yr, y_ref.detach()) <= 2e-3 * math.sqrt(K) * 0.05
        assert report(f"conv32 dgrad rows={rows}", dxr, x.grad) <= 2e-3 * math.sqrt(K) * 0.05
        assert (yr - y).abs().max().item() <= 1e-5 * max(1.0, y.abs().max().item())
        assert (dxr - dx).abs().max().item() <= 1e-5 * max(1.0, dx.abs().max().item())
    from asr_b200 import _lib
    npix = c["B"] * y_ref.shape[2] * y_ref.shape[3]
    try:
        for bf16, cls, tol in ((1, 1, 1.5e-2), (1, 0, 1.5e-2), (0, 0, 2e-3)):
            # bf16 operand copies (default): rounding sigma 1.6e-3*sqrt(npix) per entry, worst of 2e5 entries ~5 sigma; TF32
            # operands: a quarter of that.  cls: one CTA per (kernel-row class, M tile, chunk of source rows) -- the default
            # -- or per (kernel row, chunk)
            _lib.query("asrb_debug_conv_wgrad_bf16", bf16)
            _lib.query("asrb_debug_conv_wgrad_cls", cls)
            dw = ops.conv32_bwd_weight(xd, dy.to(DEV), tuple(w.shape), c["s"], c["p"])
            torch.cuda.synchronize()
            assert report(f"conv32 wgrad bf16={bf16} cls={cls}", dw, w.grad) <= tol * math.sqrt(npix)
    finally:
        _lib.query("asrb_debug_conv_wgrad_bf16", 1)
        _lib.query("asrb_debug_conv_wgrad_cls", 1)
    assert report("channel sums", ops.nchw_channel_sums(dy.to(DEV), ld), b.grad) <= 1e-4 * max(1.0, b.grad.abs().max().item())


CONV1_CASES = [
    dict(B=3, F=161, T=61, KH=41, PH=20, lens=[31, 24, 15]),
    dict(B=1, F=161, T=1001, KH=41, PH=20, lens=[501]),
    dict(B=2, F=161, T=300, KH=41, PH=20, lens=[150, 99]),
    dict(B=2, F=40, T=77, KH=7, PH=2, lens=[39, 20]),
]


@pytest.mark.parametrize("c", CONV1_CASES, ids=lambda c: f"F{c['F']}_T{c['T']}_KH{c['KH']}")
def test_conv1_tensor_core_fwd_wgrad(ops, c):
    """1->32 channel, stride-(2,2) conv as a polyphase tcgen05 implicit GEMM (TF32 operands) vs the fp64 oracle."""
    x = rnd(c["B"], 1, c["F"], c["T"], seed=18)
    w = (rnd(32, 1, c["KH"], 11, seed=19) * 0.1).requires_grad_(True)
    b = rnd(32, seed=20).requires_grad_(True)
    lens = torch.tensor(c["lens"], dtype=torch.int32)
    stride, pad = (2, 2), (c["PH"], 5)
    assert ops.conv1_supported(tuple(x.shape), tuple(w.shape), stride, pad)
    y_ref = explicit.time_mask(F.conv2d(x.double(), w.double(), b.double(), stride=stride, padding=pad), lens.long())
    dy = explicit.time_mask(rnd(*y_ref.shape, seed=21), lens.long())
    y_ref.backward(dy.double())
    K = c["KH"] * 11
    y = ops.conv1_fwd(x.to(DEV), w.detach().to(DEV), b.detach().to(DEV), lens.to(DEV), stride, pad)
    torch.cuda.synchronize()
    assert report("conv1 fwd", y, y_ref.detach()) <= 2e-3 * math.sqrt(K) * 0.1
    dw = ops.conv1_bwd_weight(x.to(DEV), dy.to(DEV), tuple(w.shape), pad)
    torch.cuda.synchronize()
    npix = c["B"] * y_ref.shape[2] * y_ref.shape[3]
    assert report("conv1 wgrad", dw, w.grad) <= 2e-3 * math.sqrt(npix)


@pytest.mark.parametrize("training", [True, False])
def test_bn_act_mask_fwd_bwd(ops, training):
    B, C, H, W = 3, 5, 7, 40
    lens = torch.tensor([40, 25, 9])
    # y is the output of the preceding masked conv: the kernel returns the gradient w.r.t. the UNMASKED conv
    # output (mask backward included), so the oracle leaf sits before the mask too
    y0 = (rnd(B, C, H, W, seed=20) * 3 + 1).requires_grad_(True)
    y = explicit.time_mask(y0, lens)
    gamma, beta = (rnd(C, seed=21) + 2).requires_grad_(True), rnd(C, seed=22).requires_grad_(True)
    rm, rv = rnd(C, seed=23) * 0.1, rnd(C, seed=24).abs() + 0.5
    yd = y.double()
    z_ref, rm_ref, rv_ref = explicit.batch_norm(yd, gamma.double(), beta.double(), rm.double(), rv.double(), training)
    z_ref = explicit.time_mask(torch.clamp(explicit.time_mask(z_ref, lens), 0.0, 20.0), lens)
    dz = rnd(B, C, H, W, seed=25)
    z_ref.backward(dz.double())
    ld = lens.int().to(DEV)
    rm_d, rv_d = rm.clone().to(DEV), rv.clone().to(DEV)
    mean, invstd = ops.bn2d_stats(y.detach().to(DEV), rm_d, rv_d, training)
    z = ops.bn_act_mask_fwd(y.detach().to(DEV), ld, mean, invstd, gamma.detach().to(DEV), beta.detach().to(DEV), True, True, 0.0, 20.0)
    assert report("bn2d fwd", z, z_ref.detach()) <= 1e-5 * 20
    assert report("bn2d running_mean", rm_d, rm_ref) <= 1e-6 and report("bn2d running_var", rv_d, rv_ref) <= 1e-5
    dy, dgamma, dbeta = ops.bn_act_mask_bwd(dz.to(DEV), y.detach().to(DEV), ld, mean, invstd, gamma.detach().to(DEV),
                                            beta.detach().to(DEV), True, True, 0.0, 20.0, training)
    assert report("bn2d dy", dy, y0.grad) <= 2e-5 * max(1.0, y0.grad.abs().max().item())
    assert report("bn2d dgamma", dgamma, gamma.grad) <= 1e-4 * max(1.0, gamma.grad.abs().max().item())
    assert report("bn2d dbeta", dbeta, beta.grad) <= 1e-4 * max(1.0, beta.grad.abs().max().item())


def test_layout_change_roundtrip(ops):
    x = rnd(3, 4, 5, 37, seed=30)
    t = ops.nchw_to_tnf(x.to(DEV))
    ref = x.view(3, 20, 37).transpose(1, 2).transpose(0, 1).contiguous()   # deepspeech.py:135-137
    assert torch.equal(t.cpu(), ref)
    assert torch.equal(ops.tnf_to_nchw(t, 4, 5).cpu(), x)


@pytest.mark.parametrize("training", [True, False])
def test_bn_rows_fwd_bwd(ops, training):
    R, H = 777, 44
    x = (rnd(R, H, seed=31) * 2 + 0.5).requires_grad_(True)
    gamma, beta = (rnd(H, seed=32) + 1.5).requires_grad_(True), rnd(H, seed=33).requires_grad_(True)
    rm, rv = rnd(H, seed=34) * 0.1, rnd(H, seed=35).abs() + 0.5
    y_ref, rm_ref, rv_ref = explicit.batch_norm(x.double(), gamma.double(), beta.double(), rm.double(), rv.double(), training)
    dy = rnd(R, H, seed=36)
    y_ref.backward(dy.double())
    rm_d, rv_d = rm.clone().to(DEV), rv.clone().to(DEV)
    y, mean, invstd = ops.bn_rows_fwd(x.detach().to(DEV), gamma.detach().to(DEV), beta.detach().to(DEV), rm_d, rv_d, training)
    assert report("bn rows fwd", y, y_ref.detach()) <= 2e-5
    assert report("bn rows rm", rm_d, rm_ref) <= 1e-6 and report("bn rows rv", rv_d, rv_ref) <= 1e-5
    dx, dgamma, dbeta = ops.bn_rows_bwd(dy.to(DEV), x.detach().to(DEV), mean, invstd, gamma.detach().to(DEV), training)
    assert report("bn rows dx", dx, x.grad) <= 2e-5 * max(1.0, x.grad.abs().max().item())
    assert report("bn rows dgamma", dgamma, gamma.grad) <= 1e-4 * max(1.0, gamma.grad.abs().max().item())
    assert report("bn rows dbeta", dbeta, beta.grad) <= 1e-4 * max(1.0, beta.grad.abs().max().item())
    assert report("col sums", ops.col_sums(dy.to(DEV)), dy.double().sum(0)) <= 1e-4


# ----------------------------------------------------------------------------- recurrence
RNN_CASES = [
    dict(cell="gru", T=7, B=3, H=16, lens=[7, 5, 2]),
    dict(cell="lstm", T=7, B=3, H=16, lens=[7, 5, 2]),
    dict(cell="gru", T=12, B=5, H=24, lens=[12, 12, 9, 4, 1]),
    dict(cell="lstm", T=9, B=2, H=40, lens=[9, 6]),
    dict(cell="gru", T=20, B=64, H=800, lens=None),
    dict(cell="lstm", T=10, B=33, H=800, lens=None),
    dict(cell="gru", T=6, B=128, H=160, lens=None),
    # rnn3.cu corner cases: one whole chain (staged TMA outputs); a whole chain + a partial one (staged + direct stores in
    # the same launch); K tail (H not a multiple of 64)
    dict(cell="lstm", T=8, B=32, H=64, lens=None),
    dict(cell="gru", T=9, B=48, H=272, lens=None),
    dict(cell="gru", T=1, B=64, H=64, lens=None),      # a single step: no recurrent product at all
    dict(cell="lstm", T=2, B=40, H=48, lens=[2] * 20 + [1] * 20),
]


def _rnn_reference(c, gi, w_hh, b_hh, lens, dout):
    """explicit oracle: the two directions on precomputed input projections, summed; autograd for the grads."""
    fn = explicit.gru_direction if c["cell"] == "gru" else explicit.lstm_direction
    gi = gi.double().requires_grad_(True)
    w = [w_hh[d].double().requires_grad_(True) for d in range(2)]
    b = [b_hh[d].double().requires_grad_(True) for d in range(2)]
    outs = [fn(gi[:, :, d], w[d], b[d], lens.long(), bool(d)) for d in range(2)]
    out = outs[0] + outs[1]
    out.backward(dout.double())
    return out.detach(), gi.grad, [x.grad for x in w], [x.grad for x in b], [o.detach() for o in outs]


@pytest.mark.parametrize("mode", ["simt_debug", "tf32", "bf16"])
@pytest.mark.parametrize("c", RNN_CASES, ids=lambda c: f"{c['cell']}_T{c['T']}_B{c['B']}_H{c['H']}")
def test_rnn_fwd_bwd(ops, c, mode):
    """simt_debug: fp32 CUDA-core product (tight tolerance, checks the algorithm); tf32 / bf16: the tcgen05 product
    with tf32 (10-bit mantissa) or bf16 (8-bit) operands, fp32 accumulate -- error grows with the mantissa loss."""
    simt = mode == "simt_debug"
    T, B, H = c["T"], c["B"], c["H"]
    if simt and H >= 800 and T > 10:
        T = 8
    cell = ops.GRU if c["cell"] == "gru" else ops.LSTM
    G = (3 if c["cell"] == "gru" else 4) * H
    lens = torch.tensor(c["lens"] if c["lens"] else sorted([max(1, T - (i * 7) % T) for i in range(B)], reverse=True),
                        dtype=torch.int32).clamp(max=T)
    k = 1.0 / math.sqrt(H)
    gi = rnd(T, B, 2, G, seed=40)
    w_hh = (torch.rand(2, G, H, generator=torch.Generator().manual_seed(41)) * 2 - 1) * k
    b_hh = (torch.rand(2, G, generator=torch.Generator().manual_seed(42)) * 2 - 1) * k
    dout = rnd(T, B, H, seed=43)
    out_ref, dgi_ref, dw_ref, db_ref, dirs_ref = _rnn_reference(c, gi, w_hh, b_hh, lens, dout)

    ops.set_debug_flags(2 if simt else 0)
    old_bf16, ops.RNN_BF16 = ops.RNN_BF16, mode == "bf16"
    old_min, ops.RNN_BF16_MIN_HIDDEN = ops.RNN_BF16_MIN_HIDDEN, 0
    try:
        ld = lens.to(DEV)
        pf, pb = ops.rnn_pack_weights(cell, w_hh[0].contiguous().to(DEV), w_hh[1].contiguous().to(DEV), B)
        hseq, cseq, saved = ops.rnn_fwd(cell, gi.to(DEV), b_hh.to(DEV), pf, ld, T, B, H)
        out = ops.rnn_sum_dirs(hseq, T, B, H)
        h2, _, _, out2 = ops.rnn_fwd(cell, gi.to(DEV), b_hh.to(DEV), pf, ld, T, B, H, want_sum=True)   # asrb_rnn_fwd_sum
        assert torch.equal(h2, hseq) and torch.equal(out2, out)
        torch.cuda.synchronize()
        tol = {"simt_debug": 1e-5, "tf32": 3e-3, "bf16": 1.2e-2}[mode]   # the product feeds back through T steps
        gtol = {"simt_debug": 5e-5, "tf32": 1e-2, "bf16": 4e-2}[mode]
        for d in range(2):
            assert report(f"rnn fwd dir{d}", hseq[d, 1:T + 1], dirs_ref[d]) <= tol
        assert hseq[:, 0].abs().max() == 0 and hseq[:, T + 1].abs().max() == 0
        assert report("rnn fwd sum", out, out_ref) <= 2 * tol
        # zero output past each length (pad_packed_sequence semantics)
        for bi, l in enumerate(lens.tolist()):
            assert out[l:, bi].abs().max().item() == 0 if l < T else True
        dgi, dgiT, dghT_k = ops.rnn_bwd(cell, dout.to(DEV), pb, ld, hseq, cseq, saved, T, B, H)
        torch.cuda.synchronize()
        gscale = dgi_ref.abs().max().item()
        assert report("rnn bwd dgi", dgi, dgi_ref) <= gtol * gscale
        R = T * B
        # the transposed copy the kernel writes for the weight-gradient GEMMs must be the same numbers
        assert torch.equal(dgiT[:, :R].cpu(), dgi.view(R, 2 * G).t().cpu())
        # parameter gradients from the kernel outputs, assembled on the host in fp64
        for d in range(2):
            first = 0 if d == 0 else 2
            hprev = hseq[d, first:first + T].reshape(R, H).double().cpu()
            src = dghT_k if c["cell"] == "gru" else dgiT                     # hidden-side gate gradients, [G, R]
            dghT = src[d * G:(d + 1) * G, :R].double().cpu().clone()
            dgh = [None, None]
            dgh[d] = dghT.t()
            dw = dghT @ hprev
            assert report(f"rnn dW_hh dir{d}", dw, dw_ref[d]) <= gtol * max(1.0, dw_ref[d].abs().max().item())
            assert report(f"rnn db_hh dir{d}", dgh[d].sum(0), db_ref[d]) <= \
                gtol * max(1.0, db_ref[d].abs().max().item())
    finally:
        ops.set_debug_flags(0)
        ops.RNN_BF16 = old_bf16
        ops.RNN_BF16_MIN_HIDDEN = old_min


@pytest.mark.parametrize("cellname", ["gru", "lstm"])
def test_rnn3_publish_protocols_agree(ops, cellname):
    """Step hand-over of rnn3.cu (asrb_debug_rnn_dbg).  Bit 4096 = generic stores + `red.release`, formally a release/acquire
    pair.  Bit 16 = ONE TMA store of the operand tile + its completion + a RELAXED counter increment (a release is a
    MEMBAR.GPU, which waits for the other chain's TMA copies: DESIGN.md section 6) -- measured to lose about one hand-over in
    10^7 (tools/stress_fullsize.py); bit 4 = the same + an L2 read-back of the tile before the increment.  The DEFAULT is the
    VERIFIED hand-over: the bit-16 form as a first pass that looks for a sentinel in every operand tile it consumes, and a
    release second pass that recomputes the launch only when one was seen (bit 8192 withholds one tile so that it must).
    All must give bit-identical outputs over a few hundred steps at the benchmarked width."""
    from asr_b200 import _lib

    T, B, H = 300, 64, 800
    cell = ops.GRU if cellname == "gru" else ops.LSTM
    G = (3 if cellname == "gru" else 4) * H
    k = 1.0 / math.sqrt(H)
    gi = rnd(T, B, 2, G, seed=140).to(DEV)
    w_hh = ((torch.rand(2, G, H, generator=torch.Generator().manual_seed(141)) * 2 - 1) * k).to(DEV)
    b_hh = ((torch.rand(2, G, generator=torch.Generator().manual_seed(142)) * 2 - 1) * k).to(DEV)
    dout = rnd(T, B, H, seed=143).to(DEV)
    lens = torch.tensor(sorted([max(1, T - (i * 7) % T) for i in range(B)], reverse=True), dtype=torch.int32).to(DEV)
    old_bf16, ops.RNN_BF16 = ops.RNN_BF16, True
    outs = {}
    try:
        pf, pb = ops.rnn_pack_weights(cell, w_hh[0].contiguous(), w_hh[1].contiguous(), B)
        for dbg in (0, 4096, 16, 4, 8192, 0):
            _lib.query("asrb_debug_rnn_dbg", dbg)
            redos = _lib.query("asrb_debug_rnn_redos")
            hseq, cseq, saved = ops.rnn_fwd(cell, gi, b_hh, pf, lens, T, B, H)
            dgi, dgiT, dghT = ops.rnn_bwd(cell, dout, pb, lens, hseq, cseq, saved, T, B, H)
            torch.cuda.synchronize()
            redos = _lib.query("asrb_debug_rnn_redos") - redos
            # verified hand-over: a withheld tile (bit 8192) must send both launches through their second pass; otherwise a
            # second pass is a one-in-10^7-hand-overs event (1.2e5 hand-overs here)
            assert redos == (2 if dbg & 8192 else 0) or (dbg == 0 and redos <= 2), (dbg, redos)
            cur = (hseq.clone(), dgi.clone(), dgiT.clone())
            assert torch.isfinite(cur[0]).all() and torch.isfinite(cur[1].float()).all()
            if dbg in outs:
                assert all(torch.equal(a, b) for a, b in zip(outs[dbg], cur))       # run-to-run: deterministic
            outs[dbg] = cur
        for other in (0, 16, 4, 8192):
            assert all(torch.equal(a, b) for a, b in zip(outs[4096], outs[other])), other
    finally:
        _lib.query("asrb_debug_rnn_dbg", 0)
        ops.RNN_BF16 = old_bf16


# ----------------------------------------------------------------------------- softmax / argmax / CTC
@pytest.mark.parametrize("R,C", [(100, 29), (33, 90), (17, 5000), (5, 1)])
def test_log_softmax_argmax(ops, R, C):
    x = rnd(R, C, seed=50) * 3
    x[0, : min(C, 3)] = 7.0          # exact tie: first maximum must win (torch.max semantics)
    xd = x.to(DEV)
    lp, pr, am = ops.log_softmax_fwd(xd, C, want_lp=True, want_probs=True, want_argmax=True)
    assert report("log_softmax", lp, x.double().log_softmax(-1)) <= 2e-6 * max(1.0, math.log(C))
    assert report("softmax", pr, x.double().softmax(-1)) <= 1e-6
    assert torch.equal(am.cpu(), torch.max(x, 1)[1])                 # bit-exact indices
    g = rnd(R, C, seed=51)
    ref = g.double() - x.double().softmax(-1) * g.double().sum(-1, keepdim=True)
    assert report("log_softmax bwd", ops.log_softmax_bwd(g.to(DEV), lp), ref) <= 2e-6 * math.sqrt(C) + 1e-5


def _ctc_gpu(ops, logits, targets, in_len, tgt_len, scale=1.0):
    lp = logits.float().log_softmax(2).contiguous().to(DEV)
    max_u = int(tgt_len.max()) if tgt_len.numel() else 0
    t, i, u = targets.int().to(DEV), in_len.int().to(DEV), tgt_len.int().to(DEV)
    loss, nll, alpha = ops.ctc_fwd(lp, t, i, u, max_u)
    grad = ops.ctc_bwd(lp, t, i, u, alpha, nll, torch.tensor([scale], device=DEV), max_u)
    torch.cuda.synchronize()
    return loss.cpu(), nll.cpu(), grad.cpu(), lp.cpu()


def test_ctc_reference_golden_cases(ops, golden):
    """the reference's own criterion (torch.nn.CTCLoss(reduction='sum')) on fixed cases incl. repeats, empty and
    infeasible targets -- tests/golden/ctc_cases.pt"""
    for c in golden("ctc_cases"):
        loss, nll, grad, _ = _ctc_gpu(ops, c["logits"], c["targets"], c["input_lengths"], c["target_lengths"])
        ref = c["nll"].double()
        assert torch.equal(torch.isinf(nll), torch.isinf(ref))
        fin = torch.isfinite(ref)
        assert report("ctc nll", nll[fin], ref[fin]) <= 1e-5 * max(1.0, ref[fin].abs().max().item())
        if c["grad_logits"] is not None:
            assert abs(loss.item() - c["loss"].item()) <= 1e-5 * abs(c["loss"].item())
            assert report("ctc grad", grad, c["grad_log_probs"]) <= 2e-5


@pytest.mark.parametrize("T,N,C,U", [(50, 8, 29, 10), (120, 4, 90, 30), (64, 3, 5000, 20), (300, 2, 29, 100)])
def test_ctc_vs_oracle(ops, T, N, C, U):
    g = torch.Generator().manual_seed(60 + C)
    logits = torch.randn(T, N, C, generator=g) * 2
    in_len = torch.tensor(sorted([T - (3 * i) % (T // 3) for i in range(N)], reverse=True))
    tgt_len = torch.tensor([max(1, U - 2 * i) for i in range(N)])
    tgts = torch.randint(1, C, (int(tgt_len.sum()),), generator=g)
    if U >= 10:
        tgts[1] = tgts[0]                                            # a repeated label (needs the blank between)
    loss, nll, grad, lp = _ctc_gpu(ops, logits, tgts, in_len, tgt_len, scale=0.25)
    nll_ref, grad_ref = explicit.ctc(lp.double().numpy(), tgts.numpy(), in_len.numpy(), tgt_len.numpy())
    assert report("ctc nll", nll, torch.from_numpy(nll_ref)) <= 1e-5 * nll_ref.max()
    assert abs(loss.item() - nll_ref.sum()) <= 1e-5 * nll_ref.sum()
    # fp32 log-space: alpha+beta-nll carries ~ulp(|nll|) of absolute error into the exponent of the posterior
    assert report("ctc grad", grad, torch.from_numpy(grad_ref) * 0.25) <= 0.25 * (1e-5 + 8 * 1.2e-7 * nll_ref.max())
    for n in range(N):
        assert grad[int(in_len[n]):, n].abs().max().item() == 0 if in_len[n] < T else True


# ----------------------------------------------------------------------------- spectrogram
def test_spectrogram_vs_oracle(ops):
    import scipy.signal

    lens = [16000, 12345, 8000]
    rng = np.random.default_rng(0)
    t = np.linspace(0, 1, 16000, endpoint=False, dtype=np.float32)
    waves = [np.sin(2 * np.pi * 440 * t).astype(np.float32),                 # the reference's own fixture signal
             (rng.standard_normal(12345) * 0.1).astype(np.float32),
             (rng.standard_normal(8000) * 0.3).astype(np.float32)]
    wav = torch.zeros(3, 16000)
    for i, w in enumerate(waves):
        wav[i, : len(w)] = torch.from_numpy(w)
    window = torch.from_numpy(scipy.signal.get_window("hamming", 320, fftbins=True)).float().to(DEV)
    basis = ops.dft_basis(320, DEV)
    for normalize in (False, True):
        spec = ops.spectrogram(wav.to(DEV), torch.tensor(lens, dtype=torch.int32).to(DEV), window, basis, 320, 160, normalize)
        assert spec.shape == (3, 1, 161, 101)
        for i, w in enumerate(waves):
            ref = explicit.spectrogram(w, normalize=normalize)
            nf = ref.shape[1]
            assert report(f"spectrogram utt{i} norm={normalize}", spec[i, 0, :, :nf], ref) <= (2e-4 if normalize else 2e-5)
            assert spec[i, 0, :, nf:].abs().max().item() == 0 if nf < 101 else True


def test_greedy_collapse_matches_host_loop(ops):
    """asrb_greedy_collapse against the reference's per-frame Python loop (decoders/greedy_decoder.py:27-46,
    remove_repetitions=True): kept classes and their frame offsets, bit-exact, ragged sizes, long rows."""
    g = torch.Generator().manual_seed(3)
    for N, T, C in [(5, 70, 4), (3, 501, 29), (2, 1, 3), (4, 33, 2)]:
        idx = torch.randint(0, C, (N, T), generator=g, dtype=torch.int64)
        idx[0, : T // 2] = 1                                  # a long run of repeats
        sizes = torch.randint(1, T + 1, (N,), generator=g, dtype=torch.int32)
        sizes[0] = T
        labels, offsets, counts = ops.greedy_collapse(idx.to(DEV), sizes.to(DEV), 0)
        labels, offsets, counts = labels.cpu(), offsets.cpu(), counts.cpu()
        for n in range(N):
            want_l, want_o = [], []
            for t in range(int(sizes[n])):
                c = int(idx[n, t])
                if c != 0 and (t == 0 or c != int(idx[n, t - 1])):
                    want_l.append(c)
                    want_o.append(t)
            k = int(counts[n])
            assert k == len(want_l)
            assert labels[n, :k].tolist() == want_l and offsets[n, :k].tolist() == want_o


# ----------------------------------------------------------------------------- lookahead convolution (8f n4)
@pytest.mark.parametrize("T,N,H,ctx", [(9, 3, 12, 5), (70, 5, 100, 20), (33, 2, 40, 1), (4, 2, 8, 20)])
def test_lookahead_fwd_bwd_vs_conv1d(ops, T, N, H, ctx):
    """kernels vs the reference formulation (blocks.py:123-128: pad (0, ctx-1) + depthwise conv1d) in fp64, with and
    without the fused Hardtanh(0, 20)"""
    x = (rnd(T, N, H, seed=41) * 8).requires_grad_(True)
    w = (rnd(H, 1, ctx, seed=42) * 0.5).requires_grad_(True)
    gy = rnd(T, N, H, seed=43)
    for act in (None, (0.0, 20.0)):
        x.grad = w.grad = None
        h = F.pad(x.double().transpose(0, 1).transpose(1, 2), (0, ctx - 1))
        ref = F.conv1d(h, w.double(), groups=H).transpose(1, 2).transpose(0, 1)
        if act is not None:
            ref = F.hardtanh(ref, *act)
        ref.backward(gy.double())
        xd, wd = x.detach().to(DEV), w.detach().to(DEV).view(H, ctx).contiguous()
        y = ops.lookahead_fwd(xd, wd, ctx, act)
        dx, dw = ops.lookahead_bwd(gy.to(DEV), xd, y, wd, ctx, act)
        torch.cuda.synchronize()
        assert report(f"lookahead fwd act={act}", y, ref.detach()) <= 1e-4
        # a clamp decision can differ from fp64 only where the pre-activation is within rounding of a bound
        assert report(f"lookahead dx act={act}", dx, x.grad) <= (1e-4 if act is None else 1e-4)
        assert report(f"lookahead dw act={act}", dw.view(H, 1, ctx), w.grad) <= 1e-3
