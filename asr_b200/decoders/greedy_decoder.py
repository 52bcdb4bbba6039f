"""Greedy CTC decoding with the contract of asr_deepspeech.decoders (decoder.py:6-58, greedy_decoder.py:6-68).

The arithmetic -- the per-frame argmax over classes (greedy_decoder.py:61, `torch.max(probs, 2)`) -- runs in our
kernel (first maximum, int64 indices: bit-exact target of SURVEY.md section 8a a11).  Collapsing repeats /
stripping blanks and the edit distances are host string work, as in the reference.
"""
from __future__ import annotations

import torch

from .. import functional as F_
from .. import ops


def _edit_distance(a, b) -> int:
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


class Decoder:
    def __init__(self, labels, blank_index=0):
        self.labels = labels
        self.int_to_char = {i: c for i, c in enumerate(labels)}
        self.blank_index = blank_index
        self.space_index = list(labels).index(" ") if " " in labels else len(labels)

    def wer(self, s1, s2):
        vocab = {w: i for i, w in enumerate(set(s1.split() + s2.split()))}
        return _edit_distance([vocab[w] for w in s1.split()], [vocab[w] for w in s2.split()])

    def cer(self, s1, s2):
        return _edit_distance(s1.replace(" ", ""), s2.replace(" ", ""))

    def decode(self, probs, sizes=None):
        raise NotImplementedError


class GreedyDecoder(Decoder):
    def convert_to_strings(self, sequences, sizes=None, remove_repetitions=False, return_offsets=False):
        strings, offsets = [], []
        for x in range(len(sequences)):
            seq_len = sizes[x] if sizes is not None else len(sequences[x])
            string, string_offsets = self.process_string(sequences[x], seq_len, remove_repetitions)
            strings.append([string])
            offsets.append([string_offsets])
        return (strings, offsets) if return_offsets else strings

    def process_string(self, sequence, size, remove_repetitions=False):
        seq = [int(v) for v in (sequence.tolist() if isinstance(sequence, torch.Tensor) else sequence)][:int(size)]
        blank = self.int_to_char[self.blank_index]
        chars, offsets = [], []
        for i, idx in enumerate(seq):
            ch = self.int_to_char[idx]
            if ch == blank:
                continue
            if remove_repetitions and i != 0 and ch == self.int_to_char[seq[i - 1]]:
                continue
            chars.append(ch)
            offsets.append(i)
        return "".join(chars), torch.tensor(offsets, dtype=torch.int)

    def decode(self, probs, sizes=None):
        """probs [B,T,C] -> (strings, offsets); argmax AND the collapse (drop blanks and repeated frames) on the device,
        only the kept class indices come back to the host for the int -> char join."""
        ops.require_cuda(probs, "GreedyDecoder.decode")
        idx = F_.argmax_last_dim(probs)
        injective = len(set(self.int_to_char.values())) == len(self.int_to_char)
        if not injective:   # the reference compares CHARACTERS of neighbouring frames: only equivalent for distinct labels
            return self.convert_to_strings(idx.cpu(), sizes, remove_repetitions=True, return_offsets=True)
        sz = None if sizes is None else ops.lengths_to_device(torch.as_tensor(sizes), probs.device)
        labels, offs, counts = ops.greedy_collapse(idx.contiguous(), sz, self.blank_index)
        labels, offs, counts = labels.cpu(), offs.cpu(), counts.cpu().tolist()
        strings, offsets = [], []
        for n, k in enumerate(counts):
            strings.append(["".join(self.int_to_char[int(c)] for c in labels[n, :k].tolist())])
            offsets.append([offs[n, :k].clone()])
        return strings, offsets
