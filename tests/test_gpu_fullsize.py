"""GPU checks at BASELINE.json's FULL sizes, where the CPU oracle is too slow: size-independent properties of the path
and a cross-check of the CTC kernels against the reference's own CUDA criterion (torch.nn.functional.ctc_loss on the
device -- what asr_deepspeech/trainers/deepspeech_trainer.py:111 runs when the reference trains on a GPU).

Properties used
  * batch additivity (BatchNorm layers on their running statistics, so utterances do not interact; everything else in
    training mode): the summed CTC loss and every parameter
    gradient of a 64-utterance ragged batch equal the sums over its two 32-utterance halves; this also runs the two
    recurrent code paths (64 batch rows: 4 TMEM lane quarters; 32 rows: 2) against each other;
  * padding: the CTC gradient is exactly zero past each utterance's length, output lengths follow get_seq_lens;
  * CTC gradient rows sum to zero (sum_c exp(lp) = 1 and the class posteriors sum to 1);
  * the backward K split over CTA pairs (DSMEM exchange) and the single-CTA backward give the same gradients.
"""
import pytest
import torch

import bench
from oracle.make_golden import synth_batch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ragged_batch(B, T, U, C, seed):
    g = torch.Generator().manual_seed(seed)
    lens = sorted(torch.randint(T // 3, T + 1, (B,), generator=g).tolist(), reverse=True)
    lens[0] = T
    return synth_batch(seed, B, T, U, C, lens), lens


def _step(model, batch, dev=DEV):
    from asr_b200.trainers import CTCLoss, fit

    for p in model.parameters():
        p.grad = None
    x = batch[0].to(dev)
    valid, loss, loss_value = fit(model, CTCLoss(reduction="sum"), (x, batch[1], batch[2], batch[3]), dev)
    loss.backward()
    torch.cuda.synchronize()
    return loss_value, {k: p.grad.detach().clone() for k, p in model.named_parameters()}


def _halves(batch, lens, T):
    """the two halves of a batch at the SAME padded length (so that (percentage * T).int() of
    deepspeech_trainer.py:104 gives every utterance the same frame count in both runs)"""
    x, targets, pct, tsz = batch
    out = []
    B = x.shape[0]
    for sl in (slice(0, B // 2), slice(B // 2, B)):
        n = int(tsz[sl].sum())
        off = int(tsz[:sl.start].sum())
        out.append((x[sl].contiguous(), targets[off:off + n], pct[sl].clone(), tsz[sl]))
    return out


def test_configs1_batch_additivity_and_padding():
    cfg = dict(bench.CFG)
    model = bench.build_model(cfg, torch.device(DEV)).train()
    for m in model.modules():                                     # running statistics: utterances independent
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.eval()
    batch, lens = _ragged_batch(cfg["B"], cfg["T"], 60, cfg["C"], seed=77)
    loss, grads = _step(model, batch)
    assert loss == loss and loss != float("inf") and loss > 0
    halves = _halves(batch, lens, cfg["T"])
    # the loss the trainer reports is sum / batch size (deepspeech_trainer.py:112): undo the division
    tot, gsum = 0.0, None
    for hb in halves:
        l, g = _step(model, hb)
        nb = hb[0].shape[0]
        tot += l * nb
        gsum = {k: v * nb for k, v in g.items()} if gsum is None else {k: gsum[k] + g[k] * nb for k in g}
    B = cfg["B"]
    assert abs(loss * B - tot) <= 2e-5 * abs(tot), (loss * B, tot)
    for k, g in grads.items():
        ref = gsum[k]
        if ref.abs().max().item() == 0:
            continue
        err = ((g * B - ref).norm() / ref.norm()).item()
        # bf16 operands in the backward GEMMs, different tile paths for 32 / 64 rows: 1 % of the tensor's norm
        assert err <= 1e-2, (k, err)

    with torch.no_grad():
        x = batch[0].to(DEV)
        sizes = batch[2].mul(int(x.size(3))).int()
        out, out_sizes = model.forward(x, sizes)     # (__call__ is the reference's evaluation loop)
    assert out.shape[0] == B and out.shape[1] == (cfg["T"] - 1) // 2 + 1
    assert out_sizes.tolist() == model.get_seq_lens(sizes).tolist()


def test_backward_k_split_matches_single_cta_backward():
    """hidden 800 -> 50 slices: the cluster-of-4 split pads them to 52 (two empty CTAs per direction)"""
    from asr_b200 import _lib

    cfg = dict(bench.CFG, layers=2)
    model = bench.build_model(cfg, torch.device(DEV)).train()
    batch, _ = _ragged_batch(cfg["B"], 401, 40, cfg["C"], seed=78)
    res = {}
    try:
        for ks in (4, 2, 0):                                     # clusters of four, CTA pairs, unsplit
            _lib.query("asrb_debug_rnn_ksplit", ks)
            res[ks] = _step(model, batch)
    finally:
        _lib.query("asrb_debug_rnn_ksplit", 2)
    l0, g0 = res[0]
    for ks in (4, 2):
        l1, g1 = res[ks]
        assert abs(l1 - l0) <= 1e-6 * abs(l0)                    # same forward
        for k in g1:
            ref = g0[k]
            if ref.abs().max().item() == 0:
                continue
            err = ((g1[k] - ref).norm() / ref.norm()).item()
            assert err <= 2e-3, (ks, k, err)                     # same products, different summation order (bf16 operands)


@pytest.mark.parametrize("T,N,C,U", [(2000, 256, 5000, 200), (501, 64, 29, 100)])
def test_ctc_full_size_against_the_cuda_criterion(T, N, C, U):
    from asr_b200 import ops

    g = torch.Generator(device=DEV).manual_seed(5 + C)
    lp = (torch.randn(T, N, C, device=DEV, generator=g) * 3).log_softmax(2)
    tg = torch.randint(1, C, (N * U,), device=DEV, generator=g, dtype=torch.int32)
    il = torch.randint(T // 2, T + 1, (N,), device=DEV, generator=g, dtype=torch.int32)
    il[0] = T
    tl = torch.randint(U // 2, U + 1, (N,), device=DEV, generator=g, dtype=torch.int32)
    offs = torch.cat([torch.zeros(1, dtype=torch.int64, device=DEV), tl.long().cumsum(0)])
    tgt = torch.cat([tg[i * U:i * U + int(tl[i])] for i in range(N)])      # concatenated, batch order
    loss, nll, ws = ops.ctc_fwd(lp, tgt.contiguous(), il, tl, U)
    grad = ops.ctc_bwd(lp, tgt.contiguous(), il, tl, ws, nll, torch.ones(1, device=DEV), U)
    torch.cuda.synchronize()
    # properties
    rows = grad.sum(2)                                            # [T, N]
    # alpha + beta - nll is a difference of fp32 numbers of size |nll| (~1.6e4 at T=2000) that each carry T steps of
    # rounding (ulp 1e-3): the posterior mass of a row is only good to a few % there, to 1e-3 at T=501
    assert rows.abs().max().item() <= (1e-1 if T >= 2000 else 5e-3), rows.abs().max().item()
    tt = torch.arange(T, device=DEV)[:, None]
    assert grad[(tt >= il[None, :].long())].abs().max().item() == 0
    # the reference's CUDA criterion on the same inputs
    lp_ref = lp.detach().clone().requires_grad_(True)
    ref = torch.nn.functional.ctc_loss(lp_ref, tgt.long().cpu(), il.long().cpu(), tl.long().cpu(), blank=0, reduction="none")
    ref.sum().backward()
    assert torch.equal(torch.isinf(nll), torch.isinf(ref))
    fin = ~torch.isinf(ref)
    assert ((nll[fin] - ref[fin]).abs() / ref[fin].abs().clamp(min=1.0)).max().item() <= 1e-5
    assert abs(loss.item() - ref[fin].sum().item()) <= 1e-5 * abs(ref[fin].sum().item())
    gerr = (grad - lp_ref.grad).abs().max().item()
    assert gerr <= (3e-2 if T >= 2000 else 1e-3), gerr           # two fp32 log-space DPs at |nll| ~ 1.6e4 resp. ~ 1e3
    del offs


def test_side_stream_weight_gradients_match_single_stream():
    """functional.WGRAD_OVERLAP: the weight-gradient work of the recurrent layers issued on a second stream (under the
    next layer's recurrent backward kernel) gives the gradients of the single-stream form -- same kernels, same
    inputs, so bit-equal wherever the kernels are deterministic (the conv weight gradients use fp32 atomics) -- also
    when `.grad` already exists (accumulation into a flat bucket's views)."""
    from asr_b200 import functional as F_
    from asr_b200.distributed import FlatGradBucket

    cfg = dict(bench.CFG, layers=3)
    model = bench.build_model(cfg, torch.device(DEV)).train()
    batch, _ = _ragged_batch(cfg["B"], 601, 50, cfg["C"], seed=79)
    old = F_.WGRAD_OVERLAP
    try:
        F_.WGRAD_OVERLAP = False
        l0, g0 = _step(model, batch)
        F_.WGRAD_OVERLAP = True
        for _ in range(3):
            l1, g1 = _step(model, batch)
            assert l1 == l0
            for k, ref in g0.items():
                if ".rnn." in k:
                    assert torch.equal(g1[k], ref), k
                else:
                    assert (g1[k] - ref).norm().item() <= 1e-5 * max(ref.norm().item(), 1e-30), k
        # existing .grad tensors (views of a flat buffer): accumulated in place on the side stream
        for p in model.parameters():
            p.grad = None
        bucket = FlatGradBucket(list(model.parameters()))
        bucket.zero()
        from asr_b200.trainers import CTCLoss, fit
        x = batch[0].to(DEV)
        _, loss, _ = fit(model, CTCLoss(reduction="sum"), (x, batch[1], batch[2], batch[3]), DEV)
        loss.backward()
        torch.cuda.synchronize()
        for k, p in model.named_parameters():
            if ".rnn." in k:
                assert torch.equal(p.grad, g0[k]), k
    finally:
        F_.WGRAD_OVERLAP = old
        for p in model.parameters():
            p.grad = None
