"""CPU: pins both oracles (oracle/torch_path.py, oracle/explicit.py) against the golden vectors
produced by the UNMODIFIED reference (oracle/make_golden.py -> tests/golden/*.pt)."""
import numpy as np
import pytest
import torch

from oracle import explicit, torch_path
from oracle.make_golden import checksum, sample_idx, synth_batch

SMALL = ["gru_small", "lstm_small", "lstm_c90"]


def _case(golden, name):
    g = golden(name)
    p = torch_path.init_params(g["rnn_type"], g["hidden"], g["layers"], g["C"])
    for k, cs in g["param_checksums"].items():
        assert torch.allclose(checksum(p[k]), cs, rtol=0, atol=0), f"weight RNG stream drifted at {k}"
    batch = synth_batch(g["seed"], g["B"], g["T"], g["U"], g["C"], g["lengths"])
    assert torch.equal(checksum(batch[0]), g["input_checksum"])
    assert torch.equal(batch[1], g["targets"])
    return g, p, batch


def _check_grads(grads, g, rtol, atol_scale):
    # absolute floor = 1e-6 of the largest gradient norm in the model: conv biases feeding a
    # BatchNorm have a mathematically ZERO gradient when no frame is masked, so the reference
    # value there is pure fp32 cancellation noise (1e-5 next to weight grads of 1e2..1e3)
    floor = 1e-6 * max(d["norm"] for d in g["grads"].values())
    for k, d in g["grads"].items():
        got = grads[k].flatten()
        ref = d["samples"]
        scale = d["norm"] / max(1.0, got.numel() ** 0.5)
        assert abs(got.double().norm().item() - d["norm"]) <= rtol * d["norm"] + floor, k
        err = (got[sample_idx(got.numel())] - ref).abs().max().item()
        assert err <= atol_scale * scale + rtol * ref.abs().max().item() + floor, (k, err, scale)


@pytest.mark.parametrize("name", SMALL + ["cfg1_gru800x5"])
def test_torch_path_matches_reference_golden(golden, name):
    g, p, batch = _case(golden, name)
    loss, logits, grads, dlogits, stats = torch_path.loss_and_grads(p, *batch, rnn_type=g["rnn_type"])
    assert torch.equal(torch_path.get_seq_lens((batch[2] * g["T"]).int()), g["output_sizes"])
    # same torch ops, same order => expect (near) bit equality; allow 1e-6 rel for thread-count effects
    assert abs(loss.item() - g["loss"].item()) <= 1e-6 * abs(g["loss"].item())
    assert torch.allclose(logits, g["logits"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(dlogits, g["dlogits"], rtol=1e-4, atol=1e-7)
    _check_grads(grads, g, rtol=1e-4, atol_scale=1e-3)
    for k, v in g["running_stats"].items():
        assert torch.allclose(stats[k], v, rtol=1e-5, atol=1e-7), k
    with torch.no_grad():
        probs, _ = torch_path.forward({**p, **stats}, batch[0], (batch[2] * g["T"]).int(), g["rnn_type"],
                                      training=False)
    idx = torch_path.greedy_indices(probs)
    for n, tn in enumerate(g["output_sizes"].tolist()):
        assert torch.equal(idx[n, :tn], g["eval_argmax"][n, :tn])


@pytest.mark.parametrize("name", SMALL)
def test_explicit_matches_reference_golden(golden, name):
    g, p, batch = _case(golden, name)
    q = {k: (v.double().requires_grad_(True) if k in torch_path.trainable(p) else
             (v.double() if v.is_floating_point() else v)) for k, v in p.items()}
    stats = {}
    lengths = (batch[2] * g["T"]).int()
    out, out_len = explicit.forward(q, batch[0].double(), lengths, g["rnn_type"], True, stats)
    assert torch.equal(out_len, g["output_sizes"])
    assert torch.allclose(out.float(), g["logits"], rtol=2e-4, atol=2e-5)
    lp = out.transpose(0, 1).log_softmax(2)
    nll, grad = explicit.ctc(lp.detach().numpy(), batch[1].numpy(), out_len.numpy(), batch[3].numpy())
    loss = nll.sum() / g["B"]
    assert abs(loss - g["loss"].item()) <= 2e-5 * abs(g["loss"].item())
    # torch's CTC backward yields d/dlogits directly; chain it through the explicit forward
    dlogits = torch.from_numpy(grad).transpose(0, 1) / g["B"]
    assert torch.allclose(dlogits.float(), g["dlogits"], rtol=1e-3, atol=2e-6)
    out.backward(dlogits)
    grads = {k: q[k].grad.float() for k in torch_path.trainable(p)}
    _check_grads(grads, g, rtol=2e-3, atol_scale=5e-3)
    for k, v in g["running_stats"].items():
        assert torch.allclose(stats[k].float(), v, rtol=1e-4, atol=1e-6), k


def test_explicit_ctc_matches_reference_cases(golden):
    for c in golden("ctc_cases"):
        lp = c["logits"].double().log_softmax(2).numpy()
        nll, grad = explicit.ctc(lp, c["targets"].numpy(), c["input_lengths"].numpy(), c["target_lengths"].numpy())
        ref = c["nll"].double().numpy()
        assert np.array_equal(np.isinf(nll), np.isinf(ref))
        fin = np.isfinite(ref)
        assert np.allclose(nll[fin], ref[fin], rtol=1e-5, atol=1e-5)
        if c["grad_logits"] is not None:
            assert np.allclose(grad, c["grad_logits"].double().numpy(), rtol=1e-3, atol=2e-6)
            # torch hands back the same tensor for d/dlog_probs (the 'logits shortcut')
            assert np.allclose(grad, c["grad_log_probs"].double().numpy(), rtol=1e-3, atol=2e-6)
            assert np.abs(grad.sum(-1)).max() < 1e-6


def test_seq_lens_and_misc_golden(golden):
    m = golden("misc")
    assert torch.equal(torch_path.get_seq_lens(m["seq_lens_in"]), m["seq_lens_out"])
    mc = m["maskconv"]
    y = explicit.time_mask(explicit.conv2d(mc["x"], mc["weight"], mc["bias"], (1, 1), (1, 1)), torch.tensor([10, 4]))
    assert torch.allclose(y, mc["y"], atol=1e-6)
    assert torch.count_nonzero(y[1, :, :, 4:]) == 0


def test_spectrogram_restatement_vs_torch_stft():
    """librosa is not installable here (parity unpinned vs librosa itself); pin the restatement
    against torch.stft with the librosa-0.11 defaults, on the reference's own fixture signal
    (tests/conftest.py:15-23 of the reference: 1 s 440 Hz sine @16 kHz) and on noise."""
    import scipy.signal

    t = np.linspace(0, 1, 16000, endpoint=False, dtype=np.float32)
    for y in (np.sin(2 * np.pi * 440 * t).astype(np.float32),
              np.random.default_rng(0).standard_normal(16000).astype(np.float32) * 0.1):
        s = explicit.spectrogram(y, normalize=False)
        assert s.dtype == torch.float32 and s.shape == (161, 101) and torch.isfinite(s).all()
        win = torch.from_numpy(scipy.signal.get_window("hamming", 320, fftbins=True)).float()
        D = torch.stft(torch.from_numpy(y), 320, 160, 320, window=win, center=True, pad_mode="constant",
                       return_complex=True)
        ref = torch.log1p(D.abs())
        assert torch.allclose(s, ref, atol=2e-5)
        sn = explicit.spectrogram(y, normalize=True)
        assert abs(sn.mean().item()) < 1e-3          # reference tests/test_spectrogram_dataset.py:50-58
        assert abs(sn.std().item() - 1) < 1e-3
