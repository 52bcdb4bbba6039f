"""The CPU oracle (oracle/torch_path.py) live against the UNMODIFIED reference package, whenever /root/reference is
present (the build container; the GPU box skips this file).  The golden files under tests/golden/ are recordings of
exactly these reference runs (oracle/make_golden.py); this test closes the loop without the recording."""
import os

import pytest
import torch

from oracle import torch_path
from oracle.make_golden import LABELS29, build_reference_model, reference_fit, synth_batch
from oracle.reference_loader import load_reference

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/asr_deepspeech"), reason="reference checkout not present")


@pytest.mark.parametrize("rnn_type,hidden,layers,C", [("gru", 24, 2, 29), ("lstm", 16, 2, 29)])
def test_torch_path_reproduces_the_reference(rnn_type, hidden, layers, C):
    load_reference()
    model = build_reference_model(rnn_type, hidden, layers, LABELS29[:C])
    p = torch_path.init_params(rnn_type, hidden, layers, C)
    sd = model.state_dict()
    assert list(sd.keys()) == list(p.keys())
    for k in sd:                                   # same seeded default initialisation, same creation order
        assert torch.equal(sd[k], p[k]), k
    batch = synth_batch(21, 3, 61, [5, 4, 3], C, [61, 47, 30])
    model.train()
    loss, out, output_sizes = reference_fit(model, batch)
    model.zero_grad()
    loss.backward()
    o_loss, o_out, o_grads, o_dlogits, o_stats = torch_path.loss_and_grads(p, *batch, rnn_type=rnn_type)
    assert abs(o_loss.item() - loss.item()) <= 1e-6 * abs(loss.item())
    assert torch.allclose(o_out, out.detach(), rtol=1e-5, atol=1e-6)
    assert torch.allclose(o_dlogits, out.grad, rtol=1e-4, atol=1e-7)
    for k, v in model.named_parameters():
        assert torch.allclose(o_grads[k], v.grad, rtol=1e-3, atol=1e-5 * float(v.grad.abs().max()) + 1e-9), k
    for k, v in o_stats.items():
        assert torch.allclose(v, model.state_dict()[k], rtol=1e-5, atol=1e-7), k
    # eval: probabilities and greedy indices (decoders/greedy_decoder.py:61)
    model.eval()
    with torch.no_grad():
        input_sizes = batch[2].clone().mul_(int(batch[0].size(3))).int()
        probs, sizes = model.forward(batch[0], input_sizes)
        o_probs, o_sizes = torch_path.forward({**p, **o_stats}, batch[0], input_sizes, rnn_type, training=False)
    assert sizes.tolist() == o_sizes.tolist() == torch_path.get_seq_lens(input_sizes).tolist()
    assert torch.allclose(o_probs, probs, rtol=1e-5, atol=1e-7)
    assert torch.equal(torch.max(probs, 2)[1], torch_path.greedy_indices(o_probs))
