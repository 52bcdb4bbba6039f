"""TEST INFRASTRUCTURE -- CPU oracle #1: the reference's hot path restated with the SAME
torch library calls the reference makes, in functional form over a flat parameter dict.

Nothing under ``asr_b200/`` imports this file.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs.

Why it exists: the reference (zakuro-ai/asr, ``asr_deepspeech`` v0.4.10) owns no arithmetic;
every FLOP on the path is a ``torch`` call (pinned torch 2.12.1 in ``uv.lock:6712``; 2.11.0
is installed here -- same operator semantics).  ``/root/reference`` cannot travel to the GPU
box, so the path is restated here call-for-call, citing the reference line each step follows,
and PINNED against golden vectors produced by importing the unmodified reference modules in
the build container (``oracle/make_golden.py`` -> ``tests/golden/*.pt``; checked by
``tests/test_oracle_golden.py``, and live against the reference by
``tests/test_oracle_vs_reference.py`` whenever ``/root/reference`` is present).

Parameter names are the reference's ``state_dict`` keys, so a reference
checkpoint feeds this oracle and the device path unchanged.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence

# Conv stack geometry: asr_deepspeech/modules/deepspeech.py:59-68
CONV_SPECS = (
    dict(idx=0, bn=1, kernel=(41, 11), stride=(2, 2), padding=(20, 5)),
    dict(idx=3, bn=4, kernel=(21, 11), stride=(2, 1), padding=(10, 5)),
)
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def get_seq_lens(lengths: torch.Tensor) -> torch.Tensor:
    """asr_deepspeech/modules/deepspeech.py:275-288 -- float division chained, one final int()."""
    seq = lengths.cpu().int()
    for spec in CONV_SPECS:
        k_t, s_t, p_t = spec["kernel"][1], spec["stride"][1], spec["padding"][1]
        seq = (seq + 2 * p_t - 1 * (k_t - 1) - 1) / s_t + 1
    return seq.int()


def _time_mask(x: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
    """asr_deepspeech/modules/blocks.py:49-55 -- zero x[i, :, :, len_i:] (applied after EVERY module)."""
    t = torch.arange(x.size(3), device=x.device)
    keep = t[None, :] < lengths.to(x.device)[:, None]
    return x * keep[:, None, None, :].to(x.dtype)


def _batch_norm(x, p, prefix, training, stats_out):
    """torch BatchNorm{1,2}d semantics: training -> biased batch variance for normalisation,
    unbiased for the running estimate, momentum 0.1 (nn defaults used at deepspeech.py:62,65,104
    and blocks.py:75)."""
    rm, rv = p.get(prefix + ".running_mean"), p.get(prefix + ".running_var")
    if training:
        rm = rm.clone() if rm is not None else None
        rv = rv.clone() if rv is not None else None
    y = F.batch_norm(x, rm, rv, p[prefix + ".weight"], p[prefix + ".bias"],
                     training=training, momentum=BN_MOMENTUM, eps=BN_EPS)
    if training and stats_out is not None and rm is not None:
        stats_out[prefix + ".running_mean"] = rm
        stats_out[prefix + ".running_var"] = rv
    return y


def mask_conv(p, x, out_lengths, training=True, stats_out=None):
    """MaskConv over the six-module conv stack (blocks.py:42-56 on deepspeech.py:59-68)."""
    for spec in CONV_SPECS:
        cw, cb = p[f"conv.seq_module.{spec['idx']}.weight"], p[f"conv.seq_module.{spec['idx']}.bias"]
        x = F.conv2d(x, cw, cb, stride=spec["stride"], padding=spec["padding"])
        x = _time_mask(x, out_lengths)
        x = _batch_norm(x, p, f"conv.seq_module.{spec['bn']}", training, stats_out)
        x = _time_mask(x, out_lengths)
        x = F.hardtanh(x, 0.0, 20.0)
        x = _time_mask(x, out_lengths)
    return x


def _rnn_module(rnn_type: str, input_size: int, hidden: int, bidirectional: bool):
    cls = {"gru": torch.nn.GRU, "lstm": torch.nn.LSTM}[rnn_type]
    return cls(input_size=input_size, hidden_size=hidden, bidirectional=bidirectional, bias=True)


def batch_rnn(p, layer, x, out_lengths, rnn_type, bidirectional=True, training=True, stats_out=None, rnn_modules=None):
    """BatchRNN.forward (blocks.py:84-93): [SequenceWise BN1d] -> pack -> 1-layer RNN -> pad -> sum dirs.
    rnn_modules: optional list of ready nn.GRU / nn.LSTM modules holding the layer weights (as the reference's BatchRNN
    does, blocks.py:75-78, with flatten_parameters() on CUDA): used instead of the functional call on `p`."""
    T, N, I = x.shape
    pre = f"rnns.{layer}"
    if pre + ".batch_norm.module.weight" in p:
        x = _batch_norm(x.reshape(T * N, I), p, pre + ".batch_norm.module", training, stats_out).reshape(T, N, I)
    hidden = p[pre + ".rnn.weight_hh_l0"].shape[1]
    packed = pack_padded_sequence(x, out_lengths.cpu())
    if rnn_modules is not None:
        y, _ = rnn_modules[layer](packed)
    else:
        mod = _rnn_module(rnn_type, I, hidden, bidirectional).to(x.dtype)
        names = [n for n, _ in mod.named_parameters()]
        weights = {n: p[f"{pre}.rnn.{n}"] for n in names}
        y, _ = torch.func.functional_call(mod, weights, (packed,))
    y, _ = pad_packed_sequence(y)
    if bidirectional:
        y = y.view(y.size(0), y.size(1), 2, -1).sum(2)
    return y


def n_rnn_layers(p) -> int:
    n = 0
    while f"rnns.{n}.rnn.weight_ih_l0" in p:
        n += 1
    return n


def forward(p, x, lengths, rnn_type="gru", bidirectional=True, training=True, stats_out=None, rnn_modules=None):
    """DeepSpeech.forward (deepspeech.py:130-149).  Returns (out[N,T',C], output_lengths)."""
    out_lengths = get_seq_lens(lengths)
    x = mask_conv(p, x, out_lengths, training, stats_out)
    b, c, d, t = x.shape
    x = x.view(b, c * d, t).transpose(1, 2).transpose(0, 1).contiguous()  # T x N x (C*D)
    for layer in range(n_rnn_layers(p)):
        x = batch_rnn(p, layer, x, out_lengths, rnn_type, bidirectional, training, stats_out, rnn_modules)
    T, N, H = x.shape
    x = _batch_norm(x.reshape(T * N, H), p, "fc.0.module.0", training, stats_out)
    x = F.linear(x, p["fc.0.module.1.weight"]).view(T, N, -1)  # Linear(bias=False), deepspeech.py:105
    x = x.transpose(0, 1)
    if not training:
        x = F.softmax(x, dim=-1)  # InferenceBatchSoftmax, blocks.py:59-64
    return x, out_lengths


def fit_loss(p, inputs, targets, input_percentages, target_sizes, rnn_type="gru", stats_out=None, rnn_modules=None):
    """The six arithmetic lines of DeepSpeechTrainer.fit (trainers/deepspeech_trainer.py:104-112)
    with criterion = CTCLoss(reduction='sum') (trainers/__main__.py:53).  Returns (loss/B, logits)."""
    input_sizes = (input_percentages * int(inputs.size(3))).int()
    out, output_sizes = forward(p, inputs, input_sizes, rnn_type, True, True, stats_out, rnn_modules)
    log_probs = out.transpose(0, 1).float().log_softmax(2)
    loss = F.ctc_loss(log_probs, targets, output_sizes, target_sizes, blank=0, reduction="sum",
                      zero_infinity=False)
    return loss / inputs.size(0), out


def greedy_indices(probs):
    """GreedyDecoder.decode's arithmetic: torch.max(probs, 2) (decoders/greedy_decoder.py:61)."""
    return torch.max(probs, 2)[1]


def init_params(rnn_type="gru", hidden=800, layers=5, num_classes=29, seed=123456, freq_bins=161,
                dtype=torch.float32):
    """Reference default initialisation (nn.Module defaults) under torch.manual_seed(seed)
    (seed mirrors asr_deepspeech/vars.py:13), built WITHOUT the reference package so that it
    runs on the GPU box; parameter creation order follows build_network (deepspeech.py:58-110)
    so the values are identical to instantiating the reference model after the same seed
    (checked by oracle/make_golden.py at generation time and by tests/test_oracle_vs_reference.py)."""
    import math
    from collections import OrderedDict

    nn = torch.nn
    state = torch.get_rng_state()
    torch.manual_seed(seed)
    try:
        conv = nn.Sequential(
            nn.Conv2d(1, 32, kernel_size=(41, 11), stride=(2, 2), padding=(20, 5)),
            nn.BatchNorm2d(32), nn.Hardtanh(0, 20),
            nn.Conv2d(32, 32, kernel_size=(21, 11), stride=(2, 1), padding=(10, 5)),
            nn.BatchNorm2d(32), nn.Hardtanh(0, 20))
        d = int(math.floor(freq_bins + 2 * 20 - 41) / 2 + 1)
        d = int(math.floor(d + 2 * 10 - 21) / 2 + 1)
        rnn_in = d * 32
        cls = {"gru": nn.GRU, "lstm": nn.LSTM}[rnn_type]
        rnns, bns = [], []
        for l in range(layers):
            isz = rnn_in if l == 0 else hidden
            bns.append(nn.BatchNorm1d(isz) if l > 0 else None)
            rnns.append(cls(input_size=isz, hidden_size=hidden, bidirectional=True, bias=True))
        fc_bn = nn.BatchNorm1d(hidden)
        fc = nn.Linear(hidden, num_classes, bias=False)
    finally:
        torch.set_rng_state(state)
    p = OrderedDict()
    for k, v in conv.state_dict().items():
        p[f"conv.seq_module.{k}"] = v
    for l in range(layers):
        if bns[l] is not None:
            for k, v in bns[l].state_dict().items():
                p[f"rnns.{l}.batch_norm.module.{k}"] = v
        for k, v in rnns[l].state_dict().items():
            p[f"rnns.{l}.rnn.{k}"] = v
    for k, v in fc_bn.state_dict().items():
        p[f"fc.0.module.0.{k}"] = v
    p["fc.0.module.1.weight"] = fc.weight.detach()
    return OrderedDict((k, v.detach().clone().to(dtype) if v.is_floating_point() else v.clone())
                       for k, v in p.items())


def trainable(p):
    """Names the reference exposes through model.parameters() (everything but BN buffers)."""
    return [k for k in p if not (k.endswith("running_mean") or k.endswith("running_var")
                                 or k.endswith("num_batches_tracked"))]


def loss_and_grads(p, inputs, targets, input_percentages, target_sizes, rnn_type="gru"):
    """fit() + loss.backward(): returns loss/B, logits, {name: grad}, dloss/dlogits, new BN stats."""
    q = {k: (v.detach().clone().requires_grad_(True) if k in set(trainable(p)) else v) for k, v in p.items()}
    stats = {}
    loss, out = fit_loss(q, inputs, targets, input_percentages, target_sizes, rnn_type, stats)
    out.retain_grad()
    loss.backward()
    grads = {k: q[k].grad for k in trainable(p)}
    return loss.detach(), out.detach(), grads, out.grad, stats
