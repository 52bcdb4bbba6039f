"""Batch assembly on the GPU: raw waveforms cross PCIe, the spectrograms are made by the
sm_100a STFT kernels (asr_b200/csrc/stft.cu) and land directly in the padded batch tensor the model consumes.

Host-side mirror of two reference functions, with the same results:

  * `SpectrogramParser.parse_audio`  (asr_deepspeech/data/parsers/spectrogram_parser.py:45-60): n_fft = win_length =
    int(sample_rate * window_size), hop = int(sample_rate * window_stride), librosa.stft defaults (centre zero padding,
    periodic window, 1 + len // hop frames), |.|, log1p, (x - mean) / unbiased std over the utterance's own [F, T];
  * `_collate_fn`  (asr_deepspeech/functional.py:9-32): utterances sorted by frame count, longest first (stable),
    zero padded to the longest, `input_percentages = frames / max_frames` (float32), targets concatenated in that
    order (int32), `target_sizes` (int32).

The returned tuple is what `DeepSpeechTrainer.fit` (trainers/deepspeech_trainer.py:102-117) unpacks, except that
`inputs` is already on the device (its `.to(device)` at :106 is then a no-op).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


class GpuBatchAssembler:
    """assembler(batch) with batch = [(waveform float32 1-D numpy array or CPU tensor, target label ids), ...]
    -> (inputs [B,1,F,Tmax] on `device`, targets int32 [sum U] cpu, input_percentages f32 [B] cpu, target_sizes int32 [B] cpu)"""

    def __init__(self, audio_conf=None, device="cuda", normalize=True, sample_rate=16000, window_size=0.02,
                 window_stride=0.01, window="hamming"):
        if audio_conf is not None:
            get = (lambda k, d: audio_conf.get(k, d)) if isinstance(audio_conf, dict) else (lambda k, d: getattr(audio_conf, k, d))
            sample_rate, window_size = get("sample_rate", sample_rate), get("window_size", window_size)
            window_stride, window = get("window_stride", window_stride), get("window", window)
        self.device = torch.device(device)
        self.normalize = bool(normalize)
        self.n_fft = int(sample_rate * window_size)
        self.hop = int(sample_rate * window_stride)
        if self.n_fft <= 0 or self.hop <= 0 or self.n_fft % 2:
            raise ValueError(f"unsupported STFT geometry n_fft={self.n_fft} hop={self.hop}")
        import scipy.signal

        # librosa.stft -> librosa.filters.get_window -> scipy.signal.get_window(window, n_fft, fftbins=True)
        w = scipy.signal.get_window(window, self.n_fft, fftbins=True).astype(np.float32)
        self.window = torch.from_numpy(w).to(self.device)
        self.basis = ops.dft_basis(self.n_fft, self.device)
        self._pinned = None
        self._copied = None          # event after the last H2D copy out of the staging buffer

    def frames(self, n_samples: int) -> int:
        return 1 + n_samples // self.hop

    def _staging(self, B, S):
        """pinned host staging buffer, grown on demand and reused"""
        if self._pinned is None or self._pinned.shape[0] < B or self._pinned.shape[1] < S:
            pin = self.device.type == "cuda" and torch.cuda.is_available()
            self._pinned = torch.zeros(max(B, 1), max(S, 1), dtype=torch.float32, pin_memory=pin)
        return self._pinned[:B, :S]

    def __call__(self, batch):
        if len(batch) == 0:
            raise ValueError("empty batch")
        waves = []
        for w, _ in batch:
            w = torch.as_tensor(w)
            if w.dim() != 1 or w.numel() == 0:
                raise ValueError("waveforms must be non-empty 1-D arrays")
            waves.append(w.to(torch.float32))
        # functional.py:13: sorted(..., key=frames, reverse=True) -- stable, ties keep their order
        order = sorted(range(len(batch)), key=lambda i: self.frames(waves[i].numel()), reverse=True)
        B = len(order)
        S = max(w.numel() for w in waves)
        if self._copied is not None:     # the previous batch's copy still reads the staging buffer
            self._copied.synchronize()
        stage = self._staging(B, S)
        stage.zero_()
        n_samples = torch.empty(B, dtype=torch.int32)
        targets, target_sizes = [], torch.zeros(B, dtype=torch.int32)
        input_percentages = torch.zeros(B, dtype=torch.float32)
        max_frames = self.frames(waves[order[0]].numel())
        for x, i in enumerate(order):
            n = waves[i].numel()
            stage[x, :n].copy_(waves[i])
            n_samples[x] = n
            input_percentages[x] = self.frames(n) / float(max_frames)
            tgt = list(batch[i][1])
            target_sizes[x] = len(tgt)
            targets.extend(int(t) for t in tgt)
        wav_dev = stage.to(self.device, non_blocking=True).contiguous()
        if wav_dev.is_cuda:
            self._copied = torch.cuda.Event()
            self._copied.record()
        spec = ops.spectrogram(wav_dev, n_samples.to(self.device, non_blocking=True), self.window, self.basis, self.n_fft,
                               self.hop, self.normalize)
        # 1 + S // hop columns, S = the longest waveform: exactly the longest utterance's frame count
        return spec, torch.tensor(targets, dtype=torch.int32), input_percentages, target_sizes
