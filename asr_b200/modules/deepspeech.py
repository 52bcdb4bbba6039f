"""Drop-in counterpart of asr_deepspeech.modules.deepspeech.DeepSpeech (same constructor, attributes,
state_dict keys, forward / get_seq_lens / __call__ contracts) running on the asr_b200 kernels."""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
from torch import nn

from .. import functional as F_
from ..decoders import GreedyDecoder
from .blocks import BatchRNN, InferenceBatchSoftmax, Lookahead, MaskConv, SequenceWise

_RNN_TYPES = {"lstm": nn.LSTM, "gru": nn.GRU, "rnn": nn.RNN}


def resolve_rnn_type(spec):
    """asr_deepspeech/vars.py:21-35: 'nn.LSTM' / 'LSTM' / 'lstm' / class -> class; unknown -> ValueError."""
    if isinstance(spec, type):
        return spec
    key = str(spec).split(".")[-1].lower()
    if key not in _RNN_TYPES:
        raise ValueError(f"Unsupported rnn_type {spec!r}; expected one of {sorted(_RNN_TYPES)}")
    return _RNN_TYPES[key]


def read_labels(label_path):
    """labels.csv (one 'label' column) -> {symbol: index}, as deepspeech.py:48 builds it."""
    import pandas as pd

    column = pd.read_csv(label_path)["label"]
    return {sym: idx for idx, sym in column.items()}


class DeepSpeech(nn.Module):
    def __init__(self, audio_conf, decoder, label_path, id="asr", rnn_type="nn.LSTM", rnn_hidden_size=768,
                 rnn_hidden_layers=5, bidirectional=True, context=20, version="0.0.1", model_path=None,
                 restart_from=None):
        super().__init__()
        self.version, self.id = version, id
        self.audio_conf = audio_conf
        self.context = context
        self.rnn_hidden_size = rnn_hidden_size
        self.rnn_hidden_layers = rnn_hidden_layers
        self.rnn_type = resolve_rnn_type(rnn_type)
        self.labels = read_labels(label_path)
        self.bidirectional = bidirectional
        self.sample_rate = audio_conf.sample_rate
        self.window_size = audio_conf.window_size
        self.num_classes = len(self.labels)
        self.model_path = model_path
        self.build_network()
        self.decoder = GreedyDecoder(self.labels)

    # -- architecture: deepspeech.py:58-110 (2 masked convs, L x BatchRNN, BN + bias-free Linear head)
    def build_network(self):
        self.conv = MaskConv(nn.Sequential(
            nn.Conv2d(1, 32, kernel_size=(41, 11), stride=(2, 2), padding=(20, 5)),
            nn.BatchNorm2d(32),
            nn.Hardtanh(0, 20, inplace=True),
            nn.Conv2d(32, 32, kernel_size=(21, 11), stride=(2, 1), padding=(10, 5)),
            nn.BatchNorm2d(32),
            nn.Hardtanh(0, 20, inplace=True),
        ))
        freq = int(math.floor((self.sample_rate * self.window_size) / 2) + 1)
        freq = int(math.floor(freq + 2 * 20 - 41) / 2 + 1)
        freq = int(math.floor(freq + 2 * 10 - 21) / 2 + 1)
        rnn_input_size = freq * 32
        layers = []
        for i in range(self.rnn_hidden_layers):
            layers.append((str(i), BatchRNN(input_size=rnn_input_size if i == 0 else self.rnn_hidden_size,
                                            hidden_size=self.rnn_hidden_size, rnn_type=self.rnn_type,
                                            bidirectional=self.bidirectional, batch_norm=i > 0)))
        self.rnns = nn.Sequential(OrderedDict(layers))
        self.lookahead = None if self.bidirectional else nn.Sequential(
            Lookahead(self.rnn_hidden_size, context=self.context), nn.Hardtanh(0, 20, inplace=True))
        self.fc = nn.Sequential(SequenceWise(nn.Sequential(
            nn.BatchNorm1d(self.rnn_hidden_size),
            nn.Linear(self.rnn_hidden_size, self.num_classes, bias=False))))
        self.inference_softmax = InferenceBatchSoftmax()

    def finetune_from(self, model_path, nlayers=1):
        """deepspeech.py:112-128: load shape-matching tensors, freeze all but the last `nlayers` parameters."""
        own = self.state_dict()
        for k, v in torch.load(model_path, map_location="cpu").items():
            if k in own and own[k].shape == v.shape:
                own[k] = v
            else:
                print(k, own[k].shape if k in own else None, v.shape)
        self.load_state_dict(own)
        if nlayers is not None:
            for prm in list(self.parameters())[:-nlayers]:
                prm.requires_grad = False

    def get_seq_lens(self, input_length):
        return F_.conv_seq_len(input_length, [m for m in self.conv.modules() if isinstance(m, nn.Conv2d)])

    def forward(self, x, lengths):
        """x [B,1,F,T] fp32 CUDA, lengths [B] -> (out [B,T',C], output_lengths int32 CPU)  (deepspeech.py:130-149)"""
        lengths = lengths.cpu().int()
        output_lengths = self.get_seq_lens(lengths)
        x, _ = self.conv(x, output_lengths)
        x = F_.NchwToTnf.apply(x)
        cb = getattr(self, "after_rnn_backward", None)
        if cb is not None and x.requires_grad:
            # fires when the gradient w.r.t. the recurrent stack's input exists, i.e. every gradient of the recurrent
            # layers and the head has been issued: data-parallel training starts their all-reduce here, under the conv
            # backward (asr_b200.distributed.OverlappedGradSync)
            x.register_hook(lambda g, cb=cb: (cb(), g)[1])
        for rnn in self.rnns:
            x = rnn(x, output_lengths)
        if not self.bidirectional:
            la, act = self.lookahead[0], self.lookahead[1]       # Lookahead + Hardtanh(0, 20): one fused pass
            x = la(x, act=(float(act.min_val), float(act.max_val)))
        x = self.fc(x)
        x = x.transpose(0, 1)
        x = self.inference_softmax(x)
        return x, output_lengths

    def get_loader(self, manifest, batch_size, num_workers, caching=False):
        """Host-side data pipeline is outside the accelerated path: delegate to the reference's loader."""
        try:
            from asr_deepspeech.data.loaders import get_loader
        except ImportError as exc:  # pragma: no cover
            raise RuntimeError("get_loader needs the reference package asr_deepspeech (host-side data pipeline)") from exc
        return get_loader(self.audio_conf, self.labels, manifest, batch_size, num_workers, caching=caching)

    def __call__(self, loader=None, manifest=None, batch_size=None, device="auto", num_workers=32, dist=None,
                 verbose=False, half=False, output_file=None, main_proc=True, restart_from=None, cuda=True):
        """Evaluation loop with the reference's contract (deepspeech.py:161-273): returns (WER%, CER%, output_data).
        Arithmetic on the device (forward + argmax); string collapse and edit distances on the host."""
        dev = torch.device("cuda") if str(device) in ("auto", "cuda", "gpu") else torch.device(device)
        with torch.no_grad():
            if loader is None:
                loader, _ = self.get_loader(manifest=manifest, batch_size=batch_size, num_workers=num_workers)
            self.eval()
            self.to(dev)
            total_cer = total_wer = num_tokens = num_chars = 0
            output_data = []
            for inputs, targets, input_percentages, target_sizes in loader:
                input_sizes = input_percentages.mul(int(inputs.size(3))).int()  # (the reference uses the in-place mul_)
                inputs = inputs.to(dev).float()
                split_targets, offset = [], 0
                for size in target_sizes:
                    split_targets.append(targets[offset:offset + size])
                    offset += size
                out, output_sizes = self.forward(inputs, input_sizes)
                decoded, _ = self.decoder.decode(out, output_sizes)
                refs = self.decoder.convert_to_strings(split_targets)
                if output_file is not None:
                    output_data.append((out.cpu().numpy(), output_sizes.numpy(), refs))
                for hyp, ref in zip(decoded, refs):
                    transcript, reference = hyp[0], ref[0]
                    total_wer += self.decoder.wer(transcript, reference)
                    total_cer += self.decoder.cer(transcript, reference)
                    num_tokens += len(reference.split())
                    num_chars += len(reference.replace(" ", ""))
                    if verbose:
                        print(f"Ref:{reference.lower()}\nHyp:{transcript.lower()}")
            wer = float(total_wer) / num_tokens
            cer = float(total_cer) / num_chars
            if main_proc and output_file is not None:
                with open(output_file, "w") as f:
                    f.write(f"===== {wer * 100:.2f}/{cer * 100:.2f} =====\n")
            return wer * 100, cer * 100, output_data
