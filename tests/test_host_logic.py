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
