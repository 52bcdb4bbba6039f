"""CPU: host-side logic of asr_b200 (module wiring, autograd plumbing, buffer layouts, the backward formulas the
kernels implement) checked against the reference's golden vectors, with asr_b200.ops routed to the torch-CPU
kernel emulator (tests/kernel_emulator.py).  The real kernels are checked by the -m gpu tests."""
import os
from types import SimpleNamespace

import pytest
import torch

from oracle import torch_path
from oracle.make_golden import LABELS29, sample_idx, synth_batch
from tests import kernel_emulator


def audio_conf():
    return SimpleNamespace(sample_rate=16000, window_size=0.02, window_stride=0.01, window="hamming",
                           speed_volume_perturb=False, spec_augment=False, noise_dir=None, noise_prob=0.4,
                           noise_levels=(0.0, 0.5))


def build_model(tmp_path, g):
    import pandas as pd
    from asr_b200.modules import DeepSpeech

    labels = LABELS29[:g["C"]] if g["C"] <= 29 else [chr(0x3041 + i) for i in range(g["C"])]
    path = os.path.join(tmp_path, "labels.csv")
    pd.DataFrame({"label": labels}).to_csv(path, index=False)
    model = DeepSpeech(audio_conf=audio_conf(), decoder=None, label_path=path, rnn_type=f"nn.{g['rnn_type'].upper()}",
                       rnn_hidden_size=g["hidden"], rnn_hidden_layers=g["layers"], bidirectional=True)
    p = torch_path.init_params(g["rnn_type"], g["hidden"], g["layers"], g["C"])
    missing = model.load_state_dict(p, strict=True)     # reference state_dict keys load as they are
    assert not missing.missing_keys and not missing.unexpected_keys
    return model


@pytest.mark.parametrize("name", ["gru_small", "lstm_small", "lstm_c90"])
def test_fit_and_backward_match_reference_golden(golden, monkeypatch, tmp_path, name):
    kernel_emulator.install(monkeypatch)
    from asr_b200.trainers import CTCLoss, fit

    g = golden(name)
    model = build_model(tmp_path, g)
    assert model.num_classes == g["C"]
    batch = synth_batch(g["seed"], g["B"], g["T"], g["U"], g["C"], g["lengths"])
    model.train()
    valid, loss, loss_value = fit(model, CTCLoss(reduction="sum"), batch, "cpu")
    assert valid
    assert abs(loss_value - g["loss"].item()) <= 2e-5 * abs(g["loss"].item())
    loss.backward()
    floor = 1e-6 * max(d["norm"] for d in g["grads"].values())
    for k, prm in model.named_parameters():
        d = g["grads"][k]
        got = prm.grad.flatten()
        assert abs(got.double().norm().item() - d["norm"]) <= 2e-3 * d["norm"] + floor, k
        err = (got[sample_idx(got.numel())] - d["samples"]).abs().max().item()
        scale = d["norm"] / max(1.0, got.numel() ** 0.5)
        assert err <= 5e-3 * scale + 2e-3 * d["samples"].abs().max().item() + floor, (k, err)
    sd = model.state_dict()
    for k, v in g["running_stats"].items():
        assert torch.allclose(sd[k], v, rtol=1e-4, atol=1e-6), k
    assert int(sd["conv.seq_module.1.num_batches_tracked"]) == 1
    # eval-mode greedy indices
    model.eval()
    with torch.no_grad():
        probs, sizes = model.forward(batch[0], (batch[2] * g["T"]).int())
        strings, _ = model.decoder.decode(probs, sizes)
    idx = probs.argmax(-1)
    for n, tn in enumerate(sizes.tolist()):
        assert torch.equal(idx[n, :tn], g["eval_argmax"][n, :tn])
    assert [s[0] for s in strings] == g["eval_strings"]


def test_maskconv_generic_stack_and_mask(golden, monkeypatch):
    """reference tests/test_blocks_mask.py:6-14 and tests/test_spectrogram_dataset.py:328-338"""
    kernel_emulator.install(monkeypatch)
    from asr_b200.modules import MaskConv

    mc = golden("misc")["maskconv"]
    conv = torch.nn.Conv2d(1, 2, kernel_size=3, padding=1)
    with torch.no_grad():
        conv.weight.copy_(mc["weight"])
        conv.bias.copy_(mc["bias"])
    lengths = torch.tensor([10, 4])
    out, out_lengths = MaskConv(torch.nn.Sequential(conv))(mc["x"], lengths)
    assert out.shape == mc["y"].shape and torch.equal(out_lengths, lengths)
    assert torch.allclose(out, mc["y"], atol=1e-6)
    assert torch.count_nonzero(out[1, 0, :, 4:]) == 0


def test_block_shapes_like_reference_tests(monkeypatch):
    """reference tests/test_spectrogram_dataset.py:233-322 (SequenceWise, InferenceBatchSoftmax, BatchRNN)"""
    kernel_emulator.install(monkeypatch)
    from asr_b200.modules import BatchRNN, InferenceBatchSoftmax, SequenceWise

    assert SequenceWise(torch.nn.Linear(8, 4))(torch.randn(5, 2, 8)).shape == (5, 2, 4)
    assert "SequenceWise" in repr(SequenceWise(torch.nn.Linear(4, 2)))
    sm = InferenceBatchSoftmax()
    sm.eval()
    out = sm(torch.randn(3, 5))
    assert torch.allclose(out.sum(-1), torch.ones(3), atol=1e-5)
    sm.train()
    x = torch.randn(3, 5)
    assert torch.equal(sm(x), x)
    rnn = BatchRNN(input_size=8, hidden_size=16, rnn_type=torch.nn.GRU, bidirectional=True, batch_norm=False)
    rnn.eval()
    assert rnn(torch.randn(5, 2, 8), torch.tensor([5, 5])).shape == (5, 2, 16)   # directions are SUMMED
    with pytest.raises(ValueError):
        BatchRNN(8, 16, rnn_type=torch.nn.RNN)


def test_cpu_tensors_are_rejected_without_the_emulator():
    """No CPU fallback: the product path refuses CPU tensors outright."""
    from asr_b200.modules import SequenceWise
    from asr_b200.trainers import CTCLoss

    with pytest.raises(RuntimeError, match="CUDA tensor required"):
        SequenceWise(torch.nn.Linear(8, 4))(torch.randn(5, 2, 8))
    with pytest.raises(RuntimeError, match="CUDA tensor required"):
        CTCLoss()(torch.randn(5, 2, 4).log_softmax(2), torch.tensor([1, 2], dtype=torch.int32),
                  torch.tensor([5, 5], dtype=torch.int32), torch.tensor([1, 1], dtype=torch.int32))


def test_check_loss_and_seq_lens_match_reference(golden):
    from asr_b200.functional import check_loss, conv_seq_len

    m = golden("misc")
    convs = [torch.nn.Conv2d(1, 32, (41, 11), (2, 2), (20, 5)), torch.nn.Conv2d(32, 32, (21, 11), (2, 1), (10, 5))]
    assert torch.equal(conv_seq_len(m["seq_lens_in"], convs), m["seq_lens_out"])
    for v, ok, err in m["check_loss"]:
        assert check_loss(torch.tensor(v), v) == (ok, err)


def test_greedy_decoder_strings_like_reference_tests():
    """reference tests/test_greedy_decoder.py:15-90 known answers (host-side collapse)."""
    from asr_b200.decoders import GreedyDecoder

    dec = GreedyDecoder("_abc ")
    s, off = dec.process_string(torch.tensor([0, 1, 1, 0, 2, 2, 4, 3]), 8, remove_repetitions=True)
    assert s == "ab c" and off.tolist() == [1, 4, 6, 7]
    assert dec.convert_to_strings([torch.tensor([1, 0, 1])]) == [["aa"]]
    assert dec.wer("a b c", "a b c") == 0 and dec.wer("a b", "a c") == 1
    assert dec.cer("abc", "abd") == 1


def test_fused_adamw_host_logic_matches_torch(monkeypatch):
    """FusedAdamW's bookkeeping (state, param groups, flat bucket form, StepLR) with the kernel emulated on the CPU."""
    kernel_emulator.install(monkeypatch)
    from asr_b200.distributed import FlatGradBucket
    from asr_b200.optim import FusedAdamW

    hp = dict(lr=1.5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5)
    g = torch.Generator().manual_seed(1)
    ref_p = [torch.nn.Parameter(torch.randn(7, 5, generator=g)), torch.nn.Parameter(torch.randn(11, generator=g))]
    our_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    flat_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    bucket = FlatGradBucket(flat_p, flatten_params=True)
    ref, ours, flat = torch.optim.AdamW(ref_p, **hp), FusedAdamW(our_p, **hp), FusedAdamW(flat_p, bucket=bucket, **hp)
    for _ in range(4):
        bucket.zero()
        for a, b, c in zip(ref_p, our_p, flat_p):
            gr = torch.randn(a.shape, generator=g)
            a.grad, b.grad = gr.clone(), gr.clone()
            c.grad.copy_(gr)
        ref.step(); ours.step(); flat.step()
    for a, b, c in zip(ref_p, our_p, flat_p):
        assert torch.allclose(a, b, atol=1e-7) and torch.allclose(a, c, atol=1e-7)
    assert flat_p[0].data_ptr() == bucket.flat_params.data_ptr()      # parameters really live in the flat buffer


def _expected_collate(samples):
    """asr_deepspeech/functional.py:9-32 restated on (spectrogram [F,T], target) pairs"""
    batch = sorted(samples, key=lambda s: s[0].size(1), reverse=True)
    tmax = batch[0][0].size(1)
    inputs = torch.zeros(len(batch), 1, batch[0][0].size(0), tmax)
    pct = torch.zeros(len(batch), dtype=torch.float32)
    tsz = torch.zeros(len(batch), dtype=torch.int32)
    targets = []
    for x, (spec, tgt) in enumerate(batch):
        inputs[x, 0, :, :spec.size(1)] = spec
        pct[x] = spec.size(1) / float(tmax)
        tsz[x] = len(tgt)
        targets.extend(tgt)
    return inputs, torch.tensor(targets, dtype=torch.int32), pct, tsz


def test_gpu_batch_assembler_host_logic_matches_collate_fn(monkeypatch):
    """asr_b200.data.GpuBatchAssembler = SpectrogramParser.parse_audio + _collate_fn: ordering (stable on ties),
    padding, percentages, concatenated targets -- kernels emulated, spectrograms against the oracle restatement."""
    import numpy as np
    from oracle import explicit

    kernel_emulator.install(monkeypatch)
    from asr_b200.data import GpuBatchAssembler

    rng = np.random.default_rng(3)
    lens = [8000, 16000, 4321, 16050, 160, 12345]          # 16000 and 16050 tie at 101 frames: input order is kept
    batch = [((rng.standard_normal(n) * 0.2).astype(np.float32), list(rng.integers(1, 29, size=3 + i))) for i, n in enumerate(lens)]
    asm = GpuBatchAssembler(audio_conf=dict(sample_rate=16000, window_size=0.02, window_stride=0.01, window="hamming"),
                            device="cpu")
    inputs, targets, pct, tsz = asm(batch)
    exp = _expected_collate([(explicit.spectrogram(w, normalize=True), [int(t) for t in tg]) for w, tg in batch])
    assert inputs.shape == exp[0].shape == (6, 1, 161, 101)
    assert torch.equal(targets, exp[1]) and targets.dtype == torch.int32
    assert torch.equal(pct, exp[2]) and torch.equal(tsz, exp[3])
    assert (inputs - exp[0]).abs().max().item() <= 2e-4
    assert tsz[5] == 3 + 4 and inputs[5, 0, :, 2:].abs().max().item() == 0   # the 160-sample utterance: last, 2 frames, rest zero
    # second call reuses the staging buffer
    inputs2, *_ = asm(batch[:2])
    assert inputs2.shape == (2, 1, 161, 101)


def _fresh_pair(tmp_path, monkeypatch):
    """our model (reference init values of the oracle) + FusedAdamW, kernels emulated"""
    kernel_emulator.install(monkeypatch)
    from asr_b200.optim import FusedAdamW

    model = build_model(tmp_path, dict(C=26, rnn_type="gru", hidden=8, layers=2))
    opt = FusedAdamW(model.parameters(), lr=1.5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5)
    return model, opt


def test_checkpoint_layout_matches_reference_manifest(golden, tmp_path, monkeypatch):
    """a checkpoint in the reference's format (trainers/deepspeech_trainer.py:176-188) written from
    OUR model + FusedAdamW has the structure of one written from the unmodified reference (manifest generated by
    oracle/make_golden.py from /root/reference): same keys, parameter ORDER (optimizer state is indexed by position),
    shapes, dtypes and optimizer state layout -- and loads back into torch.optim.AdamW, as the reference would do."""
    from oracle.make_golden import fake_grads, reference_checkpoint, reference_optimizer

    man = golden("ref_checkpoint_manifest")
    model, opt = _fresh_pair(tmp_path, monkeypatch)
    for s in (1, 2):
        fake_grads(model, s)
        opt.step()
    ck = reference_checkpoint(model, opt)
    path = os.path.join(tmp_path, "model.pth")
    torch.save(ck, path)
    ck = torch.load(path, map_location="cpu", weights_only=False)          # deepspeech_trainer.py:157
    assert sorted(ck.keys()) == man["ckpt_keys"]
    assert [k for k, _ in model.named_parameters()] == man["param_order"]
    assert [(k, tuple(v.shape), str(v.dtype)) for k, v in ck["state_dict"].items()] == man["state_dict"]
    osd = ck["optimizer"]
    assert list(osd["param_groups"][0]["params"]) == man["opt_group_params"]
    assert set(man["opt_state_keys"]) == set(osd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}
    for k in ("lr", "betas", "eps", "weight_decay"):
        assert k in man["opt_group_keys"] and k in osd["param_groups"][0]
    # the reference side: torch.optim.AdamW takes our optimizer state and continues identically
    ref_model = build_model(tmp_path, dict(C=26, rnn_type="gru", hidden=8, layers=2))
    ref_model.load_state_dict(ck["state_dict"])
    ref_opt = reference_optimizer(ref_model)
    ref_opt.load_state_dict(osd)
    fake_grads(model, 3)
    fake_grads(ref_model, 3)
    opt.step()
    ref_opt.step()
    for (k, a), (_, b) in zip(model.named_parameters(), ref_model.named_parameters()):
        assert torch.allclose(a, b, rtol=0, atol=1e-7), k


def test_reference_checkpoint_round_trip(tmp_path, monkeypatch):
    """the real thing, where the reference tree is present (this container; not the GPU box): checkpoint written from
    the UNMODIFIED reference model + torch AdamW -> our model + FusedAdamW resume -> same next step."""
    from oracle.reference_loader import load_reference, reference_available

    if not reference_available():
        pytest.skip("reference tree not present")
    load_reference()
    from oracle.make_golden import (LABELS29 as L, build_reference_model, fake_grads, reference_checkpoint,
                                    reference_optimizer)

    ref_model = build_reference_model("gru", 8, 2, L[:26])
    ref_opt = reference_optimizer(ref_model)
    sched = torch.optim.lr_scheduler.StepLR(ref_opt, step_size=1, gamma=0.99)
    for s in (1, 2):
        fake_grads(ref_model, s)
        ref_opt.step()
    sched.step()
    path = os.path.join(tmp_path, "ref.pth")
    torch.save(reference_checkpoint(ref_model, ref_opt, scheduler=sched), path)

    ck = torch.load(path, map_location="cpu", weights_only=False)
    model, opt = _fresh_pair(tmp_path, monkeypatch)
    model.load_state_dict(ck["state_dict"])                                  # deepspeech_trainer.py:158
    opt.load_state_dict(ck["optimizer"])                                     # :160
    scheduler = ck["scheduler"]                                              # :163-167: the pickled scheduler object
    scheduler.optimizer = opt
    assert abs(opt.param_groups[0]["lr"] - 1.5e-4 * 0.99) < 1e-12
    assert ck["epoch"] == 3 and ck["metrics"] == {"wer": 0.5, "cer": 0.25}
    fake_grads(model, 3)
    fake_grads(ref_model, 3)
    opt.step()
    ref_opt.step()
    scheduler.step()
    assert abs(opt.param_groups[0]["lr"] - 1.5e-4 * 0.99 ** 2) < 1e-12
    for (k, a), (_, b) in zip(model.named_parameters(), ref_model.named_parameters()):
        assert torch.allclose(a, b, rtol=0, atol=1e-7), k
    sd = model.state_dict()
    for k, v in ref_model.state_dict().items():
        assert sd[k].shape == v.shape and sd[k].dtype == v.dtype, k


def _reference_lookahead(x, weight, context):
    """asr_deepspeech/modules/blocks.py:123-128 + the Hardtanh(0, 20) of deepspeech.py:94-101"""
    h = x.transpose(0, 1).transpose(1, 2)
    h = torch.nn.functional.pad(h, pad=(0, context - 1), value=0)
    h = torch.nn.functional.conv1d(h, weight, groups=weight.shape[0])
    h = h.transpose(1, 2).transpose(0, 1).contiguous()
    return torch.nn.functional.hardtanh(h, 0.0, 20.0)


def test_lookahead_module_matches_reference_formulation(monkeypatch):
    kernel_emulator.install(monkeypatch)
    from asr_b200.modules import Lookahead

    torch.manual_seed(3)
    la = Lookahead(12, context=5)
    assert [k for k, _ in la.state_dict().items()] == ["conv.weight"] and la.conv.weight.shape == (12, 1, 5)
    x = (torch.randn(9, 3, 12) * 8).requires_grad_(True)
    y = la(x, act=(0.0, 20.0))
    gy = torch.randn_like(y)
    y.backward(gy)
    xr = x.detach().clone().requires_grad_(True)
    wr = la.conv.weight.detach().clone().requires_grad_(True)
    yr = _reference_lookahead(xr, wr, 5)
    yr.backward(gy)
    assert torch.allclose(y, yr, atol=1e-5)
    assert torch.allclose(x.grad, xr.grad, atol=1e-5) and torch.allclose(la.conv.weight.grad, wr.grad, atol=1e-4)
    assert "Lookahead(n_features=12, context=5)" == repr(la)


@pytest.mark.parametrize("cell", ["gru", "lstm"])
def test_unidirectional_model_host_logic(tmp_path, monkeypatch, cell):
    """bidirectional=False (deepspeech.py:75-101): the unidirectional BatchRNN is the two-direction operator with an
    all-zero reverse direction, followed by Lookahead + Hardtanh -- against torch's own modules with the same weights
    (kernels emulated; the real ones are checked in tests/test_gpu_model.py::test_unidirectional_model_with_lookahead)."""
    import copy
    import pandas as pd

    kernel_emulator.install(monkeypatch)
    from asr_b200.modules import DeepSpeech

    path = os.path.join(tmp_path, "labels.csv")
    pd.DataFrame({"label": LABELS29[:26]}).to_csv(path, index=False)
    torch.manual_seed(22)
    model = DeepSpeech(audio_conf=audio_conf(), decoder=None, label_path=path, rnn_type=f"nn.{cell.upper()}",
                       rnn_hidden_size=16, rnn_hidden_layers=2, bidirectional=False, context=4)
    keys = list(model.state_dict().keys())
    assert "lookahead.0.conv.weight" in keys and not any(k.endswith("_reverse") for k in keys)
    x = torch.randn(2, 1, 161, 41)
    lens = torch.tensor([41, 25], dtype=torch.int32)
    ref_m = copy.deepcopy(model).eval()
    with torch.no_grad():
        out_len = ref_m.get_seq_lens(lens)
        h = x
        for mod in ref_m.conv.seq_module:                                   # blocks.py:42-56
            h = mod(h)
            mask = torch.arange(h.size(3))[None, :] >= out_len[:, None]
            h = h.masked_fill(mask[:, None, None, :], 0)
        h = h.view(h.size(0), h.size(1) * h.size(2), h.size(3)).transpose(1, 2).transpose(0, 1).contiguous()
        for r in ref_m.rnns:                                                # blocks.py:84-93
            if r.batch_norm is not None:
                t, n = h.size(0), h.size(1)
                h = r.batch_norm.module(h.view(t * n, -1)).view(t, n, -1)
            pk = torch.nn.utils.rnn.pack_padded_sequence(h, out_len)
            h, _ = torch.nn.utils.rnn.pad_packed_sequence(r.rnn(pk)[0])
        la = ref_m.lookahead[0]                                             # blocks.py:123-128
        ref = _reference_lookahead(h, la.conv.weight, la.context)
        t, n = ref.size(0), ref.size(1)
        ref = ref_m.fc[0].module(ref.view(t * n, -1)).view(t, n, -1).transpose(0, 1).softmax(-1)
        model.eval()
        out, got_len = model.forward(x, lens)
    assert got_len.tolist() == out_len.tolist()
    for n_, tn in enumerate(out_len.tolist()):
        assert (out[n_, :tn] - ref[n_, :tn]).abs().max().item() <= 1e-4


def test_end_of_backward_join_hook(monkeypatch):
    """asr_b200.functional queues ONE engine callback per backward pass that joins the weight-gradient side stream
    (torch.autograd's queue_callback): it must run exactly once, after the last node, and re-arm for the next pass."""
    from asr_b200 import functional as F_

    calls = []
    real_join = F_._join_side_streams

    def join():
        calls.append("join")
        real_join()

    monkeypatch.setattr(F_, "_join_side_streams", join)

    class Probe(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            return x * 2

        @staticmethod
        def backward(ctx, g):
            calls.append("node")
            F_._queue_join()
            return g * 2

    for _ in range(2):
        calls.clear()
        x = torch.ones(3, requires_grad=True)
        Probe.apply(Probe.apply(x)).sum().backward()
        assert calls == ["node", "node", "join"]
        assert F_._join_queued is False and not F_._pending


def test_ctcloss_zero_infinity_like_torch(golden, monkeypatch):
    """nn.CTCLoss(blank=0, reduction="sum", zero_infinity=True) as the reference's pipeline test builds it
    (tests/test_pipeline_e2e.py:67): an utterance that cannot be aligned adds 0 and gets a zero gradient."""
    kernel_emulator.install(monkeypatch)
    from asr_b200.trainers import CTCLoss

    c = golden("ctc_cases")[2]                      # second utterance infeasible
    assert torch.isinf(c["nll"]).tolist() == [False, True]
    for zi in (True, False):
        x = c["logits"].clone().requires_grad_(True)
        loss = CTCLoss(blank=0, reduction="sum", zero_infinity=zi)(x.log_softmax(2), c["targets"], c["input_lengths"],
                                                                   c["target_lengths"])
        xr = c["logits"].clone().requires_grad_(True)
        ref = torch.nn.CTCLoss(blank=0, reduction="sum", zero_infinity=zi)(xr.log_softmax(2), c["targets"],
                                                                            c["input_lengths"], c["target_lengths"])
        if zi:
            assert torch.isfinite(loss) and abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
            loss.backward()
            ref.backward()
            assert torch.allclose(x.grad, xr.grad, atol=1e-6)
            assert x.grad[:, 1].abs().max().item() == 0
        else:
            assert torch.isinf(loss) and torch.isinf(ref)


def test_resolve_rnn_type_like_reference_tests():
    """reference tests/test_rnn_type.py:6-19"""
    import torch.nn as nn
    from asr_b200.modules.deepspeech import resolve_rnn_type

    assert resolve_rnn_type("nn.LSTM") is nn.LSTM
    assert resolve_rnn_type("gru") is nn.GRU
    assert resolve_rnn_type("RNN") is nn.RNN
    assert resolve_rnn_type(nn.LSTM) is nn.LSTM
    with pytest.raises(ValueError):
        resolve_rnn_type("nn.Transformer")


# ----------------------------------------------------------------------------- round-1 advisor findings
@pytest.mark.parametrize("name", ["gru_small", "lstm_small"])
def test_every_parameter_gradient_has_its_own_storage(golden, monkeypatch, tmp_path, name):
    """LSTM: bias_ih.grad and bias_hh.grad are equal in value but must not alias (an in-place op over all gradients --
    GradScaler.unscale_, clip_grad_norm_ -- would otherwise be applied twice to them)."""
    kernel_emulator.install(monkeypatch)
    from asr_b200.trainers import CTCLoss, fit

    g = golden(name)
    model = build_model(tmp_path, g)
    batch = synth_batch(g["seed"], g["B"], g["T"], g["U"], g["C"], g["lengths"])
    model.train()
    _, loss, _ = fit(model, CTCLoss(reduction="sum"), batch, "cpu")
    loss.backward()
    before = {k: p.grad.clone() for k, p in model.named_parameters()}
    spans = sorted((p.grad.untyped_storage().data_ptr() + p.grad.storage_offset() * 4, p.grad.numel() * 4, k)
                   for k, p in model.named_parameters())
    for (a0, an, ka), (b0, _, kb) in zip(spans, spans[1:]):
        assert a0 + an <= b0, f"{ka} and {kb} share gradient memory"
    for p in model.parameters():      # what GradScaler.unscale_ / clip_grad_norm_ do
        p.grad.mul_(0.5)
    for k, p in model.named_parameters():
        assert torch.allclose(p.grad, 0.5 * before[k], rtol=0, atol=0), k
    # a second backward without zeroing accumulates g, not 2 g
    _, loss, _ = fit(model, CTCLoss(reduction="sum"), batch, "cpu")
    for p in model.parameters():
        p.grad = None
    loss.backward()
    again = {k: p.grad.clone() for k, p in model.named_parameters()}
    _, loss, _ = fit(model, CTCLoss(reduction="sum"), batch, "cpu")
    loss.backward()
    for k, p in model.named_parameters():
        if "running" in k:
            continue
        assert (p.grad.norm() - 2 * again[k].norm()).abs() <= 2e-2 * again[k].norm() + 1e-7, k


def test_training_step_with_flat_bucket_and_fused_adamw(golden, monkeypatch, tmp_path):
    """DeepSpeechStep + FlatGradBucket + FusedAdamW(bucket) against the same steps with torch.optim.AdamW: the step must
    re-arm the bucket (not detach its views), and a bucket whose views were detached must raise instead of stepping on a
    stale buffer."""
    kernel_emulator.install(monkeypatch)
    from asr_b200.distributed import FlatGradBucket
    from asr_b200.optim import FusedAdamW
    from asr_b200.trainers import CTCLoss, DeepSpeechStep

    g = golden("gru_small")
    batch = synth_batch(g["seed"], g["B"], g["T"], g["U"], g["C"], g["lengths"])
    hp = dict(lr=1.5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5)
    ref_model, our_model = build_model(tmp_path, g), build_model(tmp_path, g)
    ref_model.train(); our_model.train()
    ref_step = DeepSpeechStep(ref_model, CTCLoss(reduction="sum"), torch.optim.AdamW(ref_model.parameters(), **hp), "cpu")
    bucket = FlatGradBucket(our_model.parameters(), flatten_params=True)
    opt = FusedAdamW(our_model.parameters(), bucket=bucket, **hp)
    our_step = DeepSpeechStep(our_model, CTCLoss(reduction="sum"), opt, "cpu", bucket=bucket)
    for _ in range(3):
        (va, la), (vb, lb) = ref_step(batch), our_step(batch)
        assert va and vb and abs(la - lb) <= 1e-5 * abs(la)
    for (k, a), (_, b) in zip(ref_model.named_parameters(), our_model.named_parameters()):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7), k
    for p, v in zip(bucket.params, bucket.views):
        assert p.grad.data_ptr() == v.data_ptr()
    # detached views are detected
    opt.zero_grad(set_to_none=True)
    with pytest.raises(RuntimeError):
        opt.step()


def test_fused_adamw_flat_state_survives_a_checkpoint(monkeypatch):
    """flat form: state_dict() carries the moments and the step count; after load_state_dict() the next step equals the
    step of an uninterrupted run (and of torch.optim.AdamW)."""
    kernel_emulator.install(monkeypatch)
    from asr_b200.distributed import FlatGradBucket
    from asr_b200.optim import FusedAdamW

    hp = dict(lr=1.5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5)
    gen = torch.Generator().manual_seed(3)
    init = [torch.randn(6, 4, generator=gen), torch.randn(9, generator=gen)]
    grads = [[torch.randn(t.shape, generator=gen) for t in init] for _ in range(5)]

    def make():
        ps = [torch.nn.Parameter(t.clone()) for t in init]
        b = FlatGradBucket(ps, flatten_params=True)
        return ps, b, FusedAdamW(ps, bucket=b, **hp)

    def run(ps, b, opt, steps):
        for gs in steps:
            b.zero()
            for p, gr in zip(ps, gs):
                p.grad.copy_(gr)
            opt.step()

    ps_a, b_a, opt_a = make()
    run(ps_a, b_a, opt_a, grads)                        # uninterrupted
    ps_b, b_b, opt_b = make()
    run(ps_b, b_b, opt_b, grads[:3])
    sd = opt_b.state_dict()
    assert len(sd["state"]) == 2 and int(sd["state"][0]["step"]) == 3
    assert sd["state"][0]["exp_avg"].abs().sum() > 0
    ps_c, b_c, opt_c = make()
    with torch.no_grad():
        for c, b in zip(ps_c, ps_b):
            c.copy_(b)
    opt_c.load_state_dict(sd)
    run(ps_c, b_c, opt_c, grads[3:])
    ref_p = [torch.nn.Parameter(t.clone()) for t in init]
    ref = torch.optim.AdamW(ref_p, **hp)
    for gs in grads:
        for p, gr in zip(ref_p, gs):
            p.grad = gr.clone()
        ref.step()
    for a, c, r in zip(ps_a, ps_c, ref_p):
        assert torch.allclose(a, c, rtol=0, atol=1e-8)
        assert torch.allclose(a, r, atol=1e-7)


def test_weight_cache_is_keyed_on_version_and_storage(monkeypatch):
    """functional._cached: packed / concatenated weights are rebuilt after ANY in-place update of a source tensor (what an
    optimizer step does), after the storage moved (FlatGradBucket(flatten_params=True)), and kept otherwise."""
    from asr_b200 import functional as F_

    monkeypatch.setattr(F_, "WEIGHT_CACHE", True)
    w = torch.nn.Parameter(torch.randn(6, 4))
    w_r = torch.nn.Parameter(torch.randn(6, 4))
    calls = []

    def build():
        calls.append(1)
        return torch.cat([w.detach(), w_r.detach()])

    a = F_._cached((w, w_r), "cat", build)
    b = F_._cached((w, w_r), "cat", build)
    assert a is b and len(calls) == 1
    with torch.no_grad():
        w_r.add_(1.0)                                   # in-place update of the SECOND source: still a miss
    c = F_._cached((w, w_r), "cat", build)
    assert len(calls) == 2 and torch.equal(c, torch.cat([w.detach(), w_r.detach()]))
    torch.autograd.graph.increment_version(w)           # what FusedAdamW does after its kernel wrote the parameters
    F_._cached((w, w_r), "cat", build)
    assert len(calls) == 3
    w.data = w.data.clone()                             # storage moved
    F_._cached((w, w_r), "cat", build)
    assert len(calls) == 4
    F_._cached((w, w_r), "other kind", build)           # kinds do not share entries
    F_._cached((w, w_r), "cat", build)
    assert len(calls) == 5
    monkeypatch.setattr(F_, "WEIGHT_CACHE", False)
    F_._cached((w, w_r), "cat", build)
    assert len(calls) == 6


def test_fused_adamw_bumps_the_version_counters(monkeypatch):
    """the optimizer kernel writes the parameters through raw pointers: the version counters must move all the same."""
    from asr_b200 import optim

    p = torch.nn.Parameter(torch.randn(5))
    v0 = p._version
    optim._bump_versions((p,))
    assert p._version == v0 + 1


def test_sentinel_check_word_pattern_of_the_verified_hand_over():
    """csrc/rnn3.cu: rnn3_tile_has_sentinel reads ONE word per 32-byte sector of a landed operand tile
    [K block][32 rows][128 B, 128-byte swizzle].  Restated here: every sector is visited once, the 32 loads of a warp hit 32
    different banks, the word read is a VALID column whenever the width is whole sectors (H % 16 == 0) -- and can be K padding
    when it is not, which is why rnn3_launch keeps the one-pass release hand-over for those widths."""
    nkb = 13                                     # H = 800 -> 832 columns
    seen = set()
    for w0 in range(0, nkb * 128, 32):
        banks = set()
        for i in range(w0, w0 + 32):
            kb, r, sec = i >> 7, (i >> 2) & 31, i & 3
            addr = kb * 4096 + r * 128 + (((2 * sec) ^ (r & 7)) << 4) + sec * 4
            banks.add((addr >> 2) & 31)
            chunk = ((addr >> 4) & 7) ^ (r & 7)               # undo the swizzle: logical 16-byte chunk of the row
            assert chunk == 2 * sec                           # first chunk of sector `sec`
            seen.add((kb, r, sec))
        assert len(banks) == 32
    assert len(seen) == nkb * 32 * 4

    def checked_columns(H):
        """logical columns (bf16 elements) the check reads in row 0, for sectors that hold at least one valid column"""
        cols = []
        for kb in range((H + 63) // 64):
            for sec in range(4):
                first = kb * 64 + 16 * sec
                if first < H:
                    cols.append(first + 2 * sec)              # word `sec` of the sector's first chunk = elements 2 sec, 2 sec + 1
        return cols

    for H in (16, 32, 800, 1024, 1760):
        assert all(c + 1 < H for c in checked_columns(H))
    assert any(c >= 36 for c in checked_columns(36))          # a partly padded sector: the word read is padding
