#!/usr/bin/env python
"""bench.py -- utterance-seconds of audio per second through the DeepSpeech2 hot path
(spectrogram -> MaskConv -> biGRU x5 -> FC -> log_softmax -> CTC, forward + backward to every parameter gradient
[+ one NCCL all-reduce of the flat gradient bucket when N > 1]).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our sm_100a path
  python bench.py --impl reference ...                           the reference's CPU path (oracle port) on host cores

Workload (BASELINE.json configs[1]): 5 x biGRU-800, batch 64 per GPU, 10 s @ 16 kHz (161 bins x 1001 frames),
29 labels, 100-character targets, synthetic randn spectrograms (seed 1234+2), reference default init (seed 123456).
One JSON line on stdout (rank 0).  The optimizer step is not part of the metric (fwd-bwd), see DESIGN.md.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG = dict(rnn_type="gru", hidden=800, layers=5, C=29, B=64, seconds=10, T=1001, U=100, seed=1236)
CPU_SAMPLE_B = 4   # utterances of the 64 used for the bounded CPU runs (about 10 s of CPU work per step on 8 cores)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d["hbm_gbs"], d["bf16_tflops_sustained"], "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


def flops_per_step(B, Tp, H, L, C, layers_in=1312):
    """Algorithmic FLOPs of SURVEY.md section 8d for fwd (x3 for fwd+bwd), padded frames counted."""
    G = 3 * H
    rows = Tp * B
    inproj = sum(2 * rows * (layers_in if l == 0 else H) * G * 2 for l in range(L))
    rec = 2 * rows * H * G * 2 * L
    conv1 = 2 * B * 32 * 81 * Tp * 451
    conv2 = 2 * B * 32 * 41 * Tp * 32 * 231
    fc = 2 * rows * H * C
    return dict(inproj=inproj, rec=rec, conv1=conv1, conv2=conv2, fc=fc)


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons during the timed region (B200_PROFILING.md recipe: the nvidia-smi clocks line).
    Read through NVML in-process (nvidia_ml_py: the library nvidia-smi itself sits on) every 50 ms: spawning `nvidia-smi`
    from a thread initialises NVML anew on every call, which was measured to stall this process's kernel launches for
    ~15 ms now and then -- 1-2 ms per step on a 30 ms step over a handful of timed steps.  Falls back to the nvidia-smi
    subprocess (every 0.5 s) when the NVML bindings are missing."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES-relative index -> NVML handle through the PCI bus id of the torch device
            bus = torch.cuda.get_device_properties(index).pci_bus_id if hasattr(torch.cuda.get_device_properties(index), "pci_bus_id") else None
            self.handle = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(h).bus == bus:
                        self.handle = h
                        break
            if self.handle is None:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        try:
            mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        except Exception:
            mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        self.rows.append((mhz, self.max_mhz, mask))

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
            r = [x.strip() for x in out.split(",")]
            mask = sum(bit for (_, bit), v in zip(self.REASONS, r[2:6]) if v.lower().startswith("active"))
            self.rows.append((float(r[0]), float(r[1]), mask))

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop_evt.wait(0.05 if self.nvml is not None else 0.5)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        sm = sorted(r[0] for r in self.rows)
        reasons = [name for name, bit in self.REASONS if any(r[2] & bit for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.rows[0][1] if self.rows else None,
                "reasons": reasons, "samples": len(self.rows), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def make_batch(B, cfg=CFG):
    from oracle.make_golden import synth_batch

    return synth_batch(cfg["seed"], B, cfg["T"], cfg["U"], cfg["C"])


def cpu_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown CPU"


def run_cpu_port(steps, warmup, cfg=CFG, best_of=False):
    """The reference's CPU implementation of the path (oracle/torch_path.py: the same torch calls the reference
    makes, pinned to the reference's golden vectors) on a bounded sample, all host threads.  best_of: report the
    fastest of the `steps` timed steps (1 warm-up + best-of-N) instead of their mean."""
    from oracle import torch_path

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    p = torch_path.init_params(cfg["rnn_type"], cfg["hidden"], cfg["layers"], cfg["C"])
    batch = make_batch(CPU_SAMPLE_B)
    for _ in range(warmup):
        torch_path.loss_and_grads(p, *batch, rnn_type=cfg["rnn_type"])
    times = []
    for _ in range(max(steps, 1)):
        t0 = time.perf_counter()
        torch_path.loss_and_grads(p, *batch, rnn_type=cfg["rnn_type"])
        times.append(time.perf_counter() - t0)
    dt = min(times) if best_of else sum(times) / len(times)
    sample = (f"{CPU_SAMPLE_B} of the {cfg['B']} utterances (full 10 s length, same model), fwd+CTC+bwd, "
              f"{'best' if best_of else 'mean'} of {len(times)} step(s) after {warmup} warm-up, torch {torch.__version__} CPU, "
              f"{cores} threads on {cpu_name()}")
    return CPU_SAMPLE_B * cfg["seconds"] / dt, dt, cores, sample


def run_cpu_extras():
    """configs[0] exactly as stated (B=4, 1 s, U=10, seed 1234+1; 1 warm-up + best of 5, without
    and with AdamW.step) and the CPU CTC of configs[4] on a quarter of its batch (N=64 of 256, linear in N), on the
    host cores of this box.  Reported next to `cpu_baseline`, never used as the headline ratio."""
    from oracle import torch_path
    from oracle.make_golden import synth_batch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    out = {"cores": cores, "cpu": cpu_name()}
    p = torch_path.init_params("gru", 800, 5, 29)
    batch = synth_batch(1235, 4, 101, 10, 29)
    names = torch_path.trainable(p)
    params = [p[k].clone().requires_grad_(True) for k in names]
    opt = torch.optim.AdamW(params, lr=1.5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5)   # trainers/__main__.py:41-47
    best = {"fwd_bwd": 1e30, "fwd_bwd_adamw": 1e30}
    for i in range(6):
        t0 = time.perf_counter()
        _, _, grads, _, _ = torch_path.loss_and_grads(p, *batch, rnn_type="gru")
        t1 = time.perf_counter()
        for q, k in zip(params, names):
            q.grad = grads[k]
        opt.step()
        t2 = time.perf_counter()
        if i:
            best["fwd_bwd"] = min(best["fwd_bwd"], t1 - t0)
            best["fwd_bwd_adamw"] = min(best["fwd_bwd_adamw"], t2 - t0)
    out["configs0_B4_1s"] = {"ms_fwd_bwd": best["fwd_bwd"] * 1e3, "utt_sec_per_s": 4.0 / best["fwd_bwd"],
                             "ms_with_adamw_step": best["fwd_bwd_adamw"] * 1e3,
                             "utt_sec_per_s_with_adamw_step": 4.0 / best["fwd_bwd_adamw"], "timing": "1 warm-up + best of 5"}
    T, N, C, U, ns = 2000, 256, 5000, 200, 64
    g = torch.Generator().manual_seed(1239)
    lp = (torch.randn(T, ns, C, generator=g) * 3).log_softmax(2).requires_grad_(True)
    tg = torch.randint(1, C, (ns * U,), generator=g, dtype=torch.int32)
    il, tl = torch.full((ns,), T, dtype=torch.int32), torch.full((ns,), U, dtype=torch.int32)
    bestc = 1e30
    for i in range(3):
        lp.grad = None
        t0 = time.perf_counter()
        torch.nn.functional.ctc_loss(lp, tg, il, tl, reduction="sum").backward()
        if i:
            bestc = min(bestc, time.perf_counter() - t0)
    out["configs4_ctc_cpu"] = {"ms_fwd_bwd_N64": bestc * 1e3, "ms_scaled_to_N256": bestc * 1e3 * N / ns,
                               "GBs_effective": 2.0 * T * ns * C * 4 / bestc / 1e9, "timing": "1 warm-up + best of 2, N=64 of 256"}
    return out


def run_torch_cudnn(dev, steps, warmup, cfg=CFG):
    """INFORMATIONAL (not the --impl reference arm): the reference path through torch + cuDNN on THIS B200 -- what
    trainers/deepspeech_trainer.py:80-91 runs with device="cuda" -- in fp32 and under fp16 autocast (device.py:43-46),
    same batch, same model, fwd + CTC + bwd.  The reference ships no kernels of its own, so this is its GPU path and
    the honest "existing Blackwell kernel" bar."""
    from oracle import torch_path

    p = {k: v.to(dev) for k, v in torch_path.init_params(cfg["rnn_type"], cfg["hidden"], cfg["layers"], cfg["C"]).items()}
    names = set(torch_path.trainable(p))
    q = {k: (v.clone().requires_grad_(True) if k in names and ".rnn." not in k else v) for k, v in p.items()}
    # the recurrent layers as real nn.GRU modules with flattened weights, as the reference's BatchRNN holds them
    # (modules/blocks.py:75-78,81-82): cuDNN's fast path, no per-call weight compaction
    cls = {"gru": torch.nn.GRU, "lstm": torch.nn.LSTM}[cfg["rnn_type"]]
    mods = []
    for l in range(cfg["layers"]):
        w = p[f"rnns.{l}.rnn.weight_ih_l0"]
        m = cls(input_size=w.shape[1], hidden_size=cfg["hidden"], bidirectional=True, bias=True).to(dev)
        m.load_state_dict({n: p[f"rnns.{l}.rnn.{n}"] for n, _ in m.named_parameters()})
        m.flatten_parameters()
        mods.append(m)
    host = make_batch(cfg["B"])
    x = host[0].to(dev)
    out = {}
    for name, amp in (("fp32", False), ("fp16_autocast", True)):
        def one():
            for t in q.values():
                if t.requires_grad:
                    t.grad = None
            for m in mods:
                m.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
                loss, _ = torch_path.fit_loss(q, x, host[1], host[2], host[3], cfg["rnn_type"], {}, mods)
            loss.backward()
            return loss.detach()
        try:
            for _ in range(max(warmup, 2)):
                loss = one()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                loss = one()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"ms_per_step": ms, "value": cfg["B"] * cfg["seconds"] / (ms * 1e-3), "loss": loss.item()}
        except Exception as e:      # informational leg: never takes the bench line down
            out[name] = {"error": f"{type(e).__name__}: {e}"[:200]}
        torch.cuda.empty_cache()
    out["note"] = ("oracle/torch_path.py (the reference's own torch calls) with parameters and batch on cuda: torch "
                   f"{torch.__version__} + cuDNN {torch.backends.cudnn.version()}, allow_tf32 matmul={torch.backends.cuda.matmul.allow_tf32} "
                   f"cudnn={torch.backends.cudnn.allow_tf32}; unit utterance-sec/s; timed with CUDA events after warm-up")
    return out


def ragged_dp_leg(model, bucket, sync, dev, rank, world, steps, warmup, cfg=CFG):
    """BASELINE.json configs[3] as stated: global batch 64 x N, durations U(5, 20) s, the utterances dealt to the ranks by
    `frame_balanced_shards` (equal counts, nearly equal frame sums), every rank's batch sorted by length and padded to its
    own longest utterance like _collate_fn (functional.py:9-32).  Reports per-rank compute time (before the all-reduce),
    the straggler ratio and the aggregate utterance-sec/s.  Extra key only: the metric line stays on the uniform config."""
    import torch.distributed as dist

    from asr_b200.distributed import frame_balanced_shards
    from asr_b200.trainers import CTCLoss, fit
    from oracle.make_golden import synth_batch

    g = torch.Generator().manual_seed(4321)
    secs = (5.0 + 15.0 * torch.rand(cfg["B"] * world, generator=g)).tolist()
    frames = [int(100 * sec) + 1 for sec in secs]
    shards = frame_balanced_shards(frames, world)
    whole_bins = [sorted(frames, reverse=True)[r * cfg["B"]:(r + 1) * cfg["B"]] for r in range(world)]   # the reference's rule
    mine = [frames[i] for i in shards[rank]]
    tmax = mine[0]
    U = [max(1, f // 10) for f in mine]                               # ~10 characters per second
    host = synth_batch(1240 + rank, cfg["B"], tmax, U, cfg["C"], mine)
    x = host[0].to(dev)
    criterion = CTCLoss(reduction="sum")

    def one():
        bucket.zero()
        _, loss, lv = fit(model, criterion, (x, host[1], host[2], host[3]), dev)
        loss.backward()
        mid = torch.cuda.Event(enable_timing=True)
        mid.record()
        sync.finish() if sync is not None else None
        return mid

    for _ in range(max(2, warmup)):
        one()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    comp = 0.0
    e0.record()
    starts, mids = [], []
    for _ in range(steps):
        st = torch.cuda.Event(enable_timing=True)
        st.record()
        starts.append(st)
        mids.append(one())
    e1.record()
    torch.cuda.synchronize()
    total = e0.elapsed_time(e1) / steps
    comp = sum(a.elapsed_time(b) for a, b in zip(starts, mids)) / steps
    t = torch.tensor([total, comp, float(sum(mine)), float(tmax)], device=dev)
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    rows = [v.tolist() for v in allt]
    step_ms = max(r[0] for r in rows)
    comps = [r[1] for r in rows]
    audio = sum(secs)
    return {"global_batch": cfg["B"] * world, "durations": "U(5,20) s, seed 4321", "sharding": "frame_balanced_shards",
            "ms_per_step": step_ms, "value": audio / (step_ms * 1e-3), "unit": "utterance-sec/s",
            "per_rank_compute_ms": [round(c, 2) for c in comps], "straggler_ratio": max(comps) / min(comps),
            "per_rank_frames": [int(r[2]) for r in rows], "per_rank_longest": [int(r[3]) for r in rows],
            "frame_imbalance_balanced": max(r[2] for r in rows) / min(r[2] for r in rows),
            "frame_imbalance_whole_bins": max(sum(b) for b in whole_bins) / min(sum(b) for b in whole_bins),
            "padded_frame_imbalance_whole_bins": max(b[0] * len(b) for b in whole_bins) / min(b[0] * len(b) for b in whole_bins)}


def configs2_leg(dev, tf_peak):
    """BASELINE.json configs[2] on one GPU, after the timed region: 7 x biLSTM-1024, batch 128, 15 s, 90 labels, U=150,
    fwd + CTC + bwd (TF32 forward products, bf16 recurrent and backward-GEMM operands, fp32 accumulation)."""
    from asr_b200.trainers import CTCLoss, fit
    from oracle.make_golden import synth_batch

    cfg = dict(rnn_type="lstm", hidden=1024, layers=7, C=90, B=128, seconds=15, T=1501, U=150, seed=1237)
    model = build_model(cfg, dev)
    host = synth_batch(cfg["seed"], cfg["B"], cfg["T"], cfg["U"], cfg["C"])
    x = host[0].to(dev)
    crit = CTCLoss(reduction="sum")

    def one():
        for p in model.parameters():
            p.grad = None
        _, loss, lv = fit(model, crit, (x, host[1], host[2], host[3]), dev)
        loss.backward()
        return lv

    for _ in range(2):
        lv = one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        lv = one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    Tp, B, H, L, G = 751, 128, 1024, 7, 4096
    rnn_flops = 3.0 * sum(2.0 * Tp * B * (1312 if l == 0 else H) * G * 2 + 2.0 * Tp * B * H * G * 2 for l in range(L))
    floor_ms = rnn_flops / (tf_peak * 1e12) * 1e3
    out = {"workload": "configs[2]: 7xbiLSTM-1024, batch 128, 15 s, 90 labels, U=150, fwd+CTC+bwd", "ms_per_step": ms,
           "value": B * 15 / (ms * 1e-3), "unit": "utterance-sec/s", "loss": lv,
           "rnn_gemm_floor_ms": floor_ms, "frac_of_rnn_gemm_roofline": floor_ms / ms,
           "peak_mem_GB": torch.cuda.max_memory_allocated(dev) / 1e9,
           "parity": "tests/test_gpu_fullsize_golden.py::cfg3_lstm1024x7_b16 (same model, batch 16, against the unmodified reference)"}
    del model, x
    torch.cuda.empty_cache()
    return out


def isolation_rooflines(dev, hbm_peak):
    """The HBM-bound named kernels at the shapes BASELINE.json quotes them on, run AFTER the timed region (stated):
    CTC forward / backward at configs[4] (T=2000, N=256, C=5000, U=200) and the spectrogram chain at the configs[1]
    shape (64 x 10 s of waveform).  achieved = algorithmic bytes / CUDA-event time per launch."""
    from asr_b200 import ops

    def timed(fn, n=5, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    res = {}
    T, N, C, U = 2000, 256, 5000, 200
    g = torch.Generator(device=dev).manual_seed(1239)
    lp = (torch.randn(T, N, C, device=dev, generator=g) * 3).log_softmax(2)
    tg = torch.randint(1, C, (N * U,), device=dev, generator=g, dtype=torch.int32)
    il = torch.full((N,), T, dtype=torch.int32, device=dev)
    tl = torch.full((N,), U, dtype=torch.int32, device=dev)
    one = torch.ones(1, device=dev)
    st = {}

    def fwd():
        st["f"] = ops.ctc_fwd(lp, tg, il, tl, U)

    def both():      # the backward consumes the alpha workspace: forward + backward pairs
        fwd()
        loss, nll, alpha = st["f"]
        st["g"] = ops.ctc_bwd(lp, tg, il, tl, alpha, nll, one, U)

    tf = timed(fwd)
    tb = timed(both) - tf
    half = 1.0 * T * N * C * 4
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    for key, ms, work, note in (
            ("asrb_ctc_cfg5_fwd", tf, half, "configs[4] alpha pass: read log_probs [2000,256,5000] once (10.24 GB)"),
            ("asrb_ctc_cfg5_bwd", tb, 2 * half, "configs[4] beta + gradient: read log_probs, write the dense gradient (20.48 GB)"),
            ("asrb_ctc_cfg5_fwd_bwd", tf + tb, 2 * half, "configs[4] as SURVEY 8d counts it: read log_probs + write grad = 20.48 GB over "
                                                       "BOTH launches (3.13 ms floor at the measured copy rate)")):
        ach = work / (ms * 1e-3) / 1e9
        res[key] = {"bound": "hbm", "achieved": round(ach, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(ach / hbm_peak, 4),
                    "launches": 1, "avg_launch_ms": round(ms, 4), "traffic": traffic.get(key), "algorithmic": note,
                    "in_timed_region": False}
    del lp, st
    torch.cuda.empty_cache()
    import scipy.signal
    B, S = 64, 160000
    wav = torch.randn(B, S, device=dev, generator=g) * 0.1
    ns = torch.full((B,), S, dtype=torch.int32, device=dev)
    win = torch.from_numpy(scipy.signal.get_window("hamming", 320, fftbins=True)).float().to(dev)
    basis = ops.dft_basis(320, dev)
    ts = timed(lambda: ops.spectrogram(wav, ns, win, basis, 320, 160, True))
    work = B * (4.0 * S + 4.0 * 161 * (1 + S // 160))
    ach = work / (ts * 1e-3) / 1e9
    res["asrb_spectrogram"] = {"bound": "hbm", "achieved": round(ach, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(ach / hbm_peak, 4),
                               "launches": 1, "avg_launch_ms": round(ts, 4), "traffic": traffic.get("asrb_spectrogram"),
                               "algorithmic": "64 x 10 s: read the waveform, write the [161,1001] log-spectrogram (82 MB); frames + "
                                              "3xTF32 DFT GEMM + log1p|.| + per-utterance normalisation", "in_timed_region": False}
    return res


def build_model(cfg, device):
    import pandas as pd

    from asr_b200.modules import DeepSpeech
    from oracle import torch_path
    from oracle.make_golden import LABELS29

    conf = SimpleNamespace(sample_rate=16000, window_size=0.02, window_stride=0.01, window="hamming")
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "labels.csv")
        labels = LABELS29[:cfg["C"]] if cfg["C"] <= 29 else [chr(0x3041 + i) for i in range(cfg["C"])]
        pd.DataFrame({"label": labels}).to_csv(path, index=False)
        model = DeepSpeech(audio_conf=conf, decoder=None, label_path=path, rnn_type=f"nn.{cfg['rnn_type'].upper()}",
                           rnn_hidden_size=cfg["hidden"], rnn_hidden_layers=cfg["layers"])
    model.load_state_dict(torch_path.init_params(cfg["rnn_type"], cfg["hidden"], cfg["layers"], cfg["C"]), strict=True)
    return model.to(device).train()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-isolation", action="store_true", help="skip the CTC configs[4] / spectrogram isolation rooflines and the torch+cuDNN leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = CFG
    workload = (f"configs[1]: 5xbiGRU-800, batch {cfg['B']}/GPU, 10 s @16 kHz (161x1001 spectrogram), 29 labels, "
                f"U=100, fwd+CTC+bwd to all parameter gradients")
    base = {"metric": "utterance-sec/s through conv+biRNN+CTC fwd-bwd", "unit": "utterance-sec/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "data": "synthetic",
            "config": {"workload": workload, "global_batch": cfg["B"] * max(args.gpus, 1), "frames": cfg["T"],
                       "optimizer_step": "excluded (metric is fwd-bwd)", "l2_policy": "activations (GBs) far exceed the 126 MB L2",
                       "weight_packs": "packed / concatenated weight copies are cached per parameter version: with no optimizer "
                                       "step in the metric they are built once (ASRB_WEIGHT_CACHE=0 rebuilds them every step: +0.45 ms)"}}

    if args.impl == "reference":
        if rank != 0:
            return
        steps, warmup = max(1, args.steps), max(1, args.warmup)
        v, dt, cores, sample = run_cpu_port(steps, warmup)
        line = dict(base, impl="reference", value=v, ms_per_step=dt * 1e3, dtype="f32",
                    cpu_baseline={"value": v, "unit": "utterance-sec/s", "cores": cores, "kind": "port", "sample": sample},
                    e2e={"value": v, "unit": "utterance-sec/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(line), flush=True)
        return

    import torch.distributed as dist

    from asr_b200 import ops
    from asr_b200.distributed import FlatGradBucket, OverlappedGradSync
    from asr_b200.trainers import CTCLoss, fit

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model = build_model(cfg, dev)
    criterion = CTCLoss(reduction="sum")
    bucket = FlatGradBucket(model.parameters())
    # N > 1: the all-reduce of the recurrent + head gradients starts under the conv backward, the conv gradients follow
    sync = OverlappedGradSync(bucket, model) if world > 1 else None
    host = make_batch(cfg["B"])
    pinned = host[0].pin_memory()
    resident = host[0].to(dev)
    stream = torch.cuda.current_stream()

    diag = [] if os.environ.get("ASRB_BENCH_DIAG") else None

    def step(inputs):
        t0 = time.perf_counter()
        bucket.zero()
        valid, loss, loss_value = fit(model, criterion, (inputs, host[1], host[2], host[3]), dev)
        t1 = time.perf_counter()
        loss.backward()
        if diag is not None:
            diag.append((t1 - t0, time.perf_counter() - t1))
        if sync is not None:
            sync.finish()
        return loss_value

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    copy_stream = torch.cuda.Stream(device=dev)
    dev_buf = [torch.empty_like(resident), torch.empty_like(resident)]

    def timed(inputs, steps, profile=False):
        """inputs on the device: the resident-input metric.  inputs in pinned host memory: the end-to-end metric -- every
        step's spectrogram batch crosses PCIe inside the timed region (the way a DataLoader with pin_memory and a
        prefetching copy stream feeds a trainer: the copy of step i+1 overlaps the compute of step i, the first copy
        does not) and every step's loss is read back with .item() inside fit()."""
        host_inputs = not inputs.is_cuda
        barrier()
        ops.PROFILE = {} if profile else None
        l0 = ops.LAUNCHES
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ready = [None, None]
        if host_inputs:
            copy_stream.wait_stream(stream)
            with torch.cuda.stream(copy_stream):
                dev_buf[0].copy_(inputs, non_blocking=True)
                ready[0] = torch.cuda.Event()
                ready[0].record(copy_stream)
        for i in range(steps):
            if host_inputs:
                stream.wait_event(ready[i % 2])
                if i + 1 < steps:
                    copy_stream.wait_stream(stream)          # buffer (i+1)%2 was last read by step i-1, already queued
                    with torch.cuda.stream(copy_stream):
                        dev_buf[(i + 1) % 2].copy_(inputs, non_blocking=True)
                        ready[(i + 1) % 2] = torch.cuda.Event()
                        ready[(i + 1) % 2].record(copy_stream)
                lv = step(dev_buf[i % 2])
            else:
                lv = step(inputs)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        prof, ops.PROFILE = ops.PROFILE, None
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms / steps, lv, ops.LAUNCHES - l0, prof

    for _ in range(max(args.warmup, 3)):
        step(resident)
    sampler = ClockSampler(local_rank) if rank == 0 and not os.environ.get("ASRB_BENCH_NO_SAMPLER") else None
    if sampler:
        sampler.start()
    ms_dev, loss_value, launches, _ = timed(resident, args.steps)
    # the host-input form has its own warm-up (copy stream, staging buffers, the allocator's blocks for that tensor pattern):
    # measured without it, its first leg ran 1.7 ms/step slower than a second one (32.1 against 30.4)
    # (a whole untimed leg of the same length: three warm-up steps were measured not to be enough -- a one-time ~30 ms stall,
    # the caching allocator settling on the host-input tensor pattern, landed inside the first 8 timed steps; the following
    # leg then runs at the resident form's pace)
    timed(pinned, max(args.warmup, args.steps, 10))     # (measured: five untimed steps of this form are not enough, ten are)
    ms_e2e, _, _, _ = timed(pinned, args.steps)          # host buffers: pinned H2D + loss D2H inside the region
    if os.environ.get("ASRB_BENCH_DIAG"):                # diagnostic: the two legs again, in the other order
        ms_e2e_b, _, _, _ = timed(pinned, args.steps)
        ms_dev_b, _, _, _ = timed(resident, args.steps)
        print(f"[diag] resident {ms_dev:.2f}, host {ms_e2e:.2f}, host again {ms_e2e_b:.2f}, resident again {ms_dev_b:.2f} ms/step", file=sys.stderr, flush=True)
        tail = diag[-args.steps:]
        print(f"[diag] host time per step: fit() incl. the loss read-back {1e3 * sum(a for a, _ in tail) / len(tail):.2f} ms, "
              f"backward() enqueue {1e3 * sum(b for _, b in tail) / len(tail):.2f} ms", file=sys.stderr, flush=True)
    # the clock sampler covers the two timed regions above and stops here: its periodic nvidia-smi spawn stalls the
    # launching thread for ~15 ms now and then, which is noise in `value` but lands on ONE kernel's event pair below
    clocks = sampler.stop() if sampler else None
    # separate, untimed-for-the-metric pass with a CUDA-event pair around every C-ABI call (the event records cost ~2 ms
    # of host time per step, which would otherwise leak into `value`): per-kernel device times for the rooflines
    # -- also with the side-stream overlap of the weight gradients switched off, so that every kernel is timed running
    # alone (what a roofline fraction means); `value` and `e2e` above are measured with the overlap on
    from asr_b200 import functional as F_
    overlap, F_.WGRAD_OVERLAP = F_.WGRAD_OVERLAP, False
    for _ in range(2):      # this form allocates from the main stream's pool: warm it before timing kernels
        step(resident)
    _, _, _, prof = timed(resident, args.steps, profile=True)
    F_.WGRAD_OVERLAP = overlap
    ragged = None
    if world > 1:
        try:
            ragged = ragged_dp_leg(model, bucket, sync, dev, rank, world, args.steps, args.warmup)
        except Exception as e:
            ragged = {"error": f"{type(e).__name__}: {e}"[:300]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_sec = cfg["B"] * cfg["seconds"] * world
    hbm_peak, tf_peak, peak_src = peaks()
    # per-entry-point device time over the timed region (CUDA events on the launching stream)
    # per step: the sum over the entry point's launches; over the K steps: the median (one host stall -- e.g. a
    # process spawn -- otherwise lands on a single kernel's event pair and skews its mean)
    def per_step_ms(v):
        n = len(v) // args.steps
        if n == 0 or len(v) % args.steps:
            return sum(s.elapsed_time(e) for s, e in v) / args.steps
        sums = sorted(sum(s.elapsed_time(e) for s, e in v[i * n:(i + 1) * n]) for i in range(args.steps))
        mid = len(sums) // 2
        return sums[mid] if len(sums) % 2 else 0.5 * (sums[mid - 1] + sums[mid])
    per_op = {k: (per_step_ms(v), len(v) // args.steps) for k, v in prof.items()}
    Tp = (cfg["T"] - 1) // 2 + 1
    fl = flops_per_step(cfg["B"], Tp, cfg["hidden"], cfg["layers"], cfg["C"])
    # every kernel group against its roofline: achieved = ALGORITHMIC work per launch / mean launch duration
    B_, C_ = cfg["B"], cfg["C"]
    groups = {   # entry point -> (bound, algorithmic work per STEP [FLOP or bytes], note)
        "asrb_rnn_fwd": ("tensor", fl["rec"], "2*T'*B*H*G*2dirs FLOP per layer-launch; latency chain, see DESIGN.md 6"),
        "asrb_rnn_bwd": ("tensor", fl["rec"], "2*T'*B*H*G*2dirs FLOP per layer-launch (dh = dgates W_hh)"),
        "asrb_gemm_tn": ("tensor", fl["inproj"] + 3 * fl["fc"],
                         "the tf32 GEMM launches of a step: in-proj fwd, FC fwd + dgrad + wgrad (tf32 operands run at half "
                         "the bf16 peak quoted)"),
        "asrb_gemm_tn_bf16": ("tensor", 2 * fl["inproj"] + fl["rec"],
                              "the bf16 GEMM launches of a step (gradient-only products): in-proj dgrad + wgrad, dW_hh"),
        "asrb_conv32_fwd": ("tensor", fl["conv2"], "conv2 forward"),
        "asrb_conv32_bwd_data": ("tensor", fl["conv2"], "conv2 input gradient"),
        "asrb_conv32_fwd_rows": ("tensor", fl["conv2"], "conv2 forward, 4 output rows per work item (tf32 operands: half the bf16 peak quoted)"),
        "asrb_conv32_bwd_data_rows": ("tensor", fl["conv2"], "conv2 input gradient, 4 output rows per work item (tf32 operands)"),
        "asrb_conv32_bwd_weight": ("tensor", fl["conv2"], "conv2 weight gradient"),
        "asrb_conv1_fwd": ("tensor", fl["conv1"], "conv1 forward"),
        "asrb_conv1_bwd_weight": ("tensor", fl["conv1"], "conv1 weight gradient"),
        "asrb_ctc_fwd": ("hbm", Tp * B_ * C_ * 4.0, "read log_probs once (alpha); launch-latency regime at C=29"),
        "asrb_ctc_bwd": ("hbm", 2.0 * Tp * B_ * C_ * 4.0, "read log_probs + write the dense gradient"),
    }
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")   # dram bytes per launch from the ncu --set full captures
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    rooflines = {}
    for k, (bound, work, note) in groups.items():
        if k not in per_op or per_op[k][0] <= 0:
            continue
        ms_step, n_launch = per_op[k]
        if bound == "tensor":
            ach, peak, unit = work / (ms_step * 1e-3) / 1e12, tf_peak, "TFLOP/s"
        else:
            ach, peak, unit = work / (ms_step * 1e-3) / 1e9, hbm_peak, "GB/s"
        rooflines[k] = {"bound": bound, "achieved": round(ach, 2), "peak": peak, "unit": unit, "frac": round(ach / peak, 4),
                        "launches": n_launch, "avg_launch_ms": round(ms_step / n_launch, 4),
                        "traffic": traffic.get(k), "algorithmic": note}
    top = max((k for k in per_op if k in rooflines), key=lambda k: per_op[k][0])
    what = f"{top}: {groups[top][2]}"
    avg_ms = per_op[top][0] / per_op[top][1]
    achieved = rooflines[top]["achieved"]
    base["config"] = dict(base["config"], streams="weight gradients of the recurrent layers on a second stream (value, "
                          "e2e); the per-kernel rooflines are timed in a separate single-stream pass")
    line = dict(base, value=total_sec / (ms_dev * 1e-3), ms_per_step=ms_dev, dtype="tf32 operands, f32 accumulate/storage",
                loss=loss_value, gpu_launches=launches,
                e2e={"value": total_sec / (ms_e2e * 1e-3), "unit": "utterance-sec/s", "ms_per_step": ms_e2e,
                     "h2d_bytes_per_step": pinned.numel() * 4 + host[1].numel() * 4 + 2 * cfg["B"] * 4 + 3 * cfg["B"] * 4,
                     "d2h_bytes_per_step": 4},
                clocks=clocks,
                roofline={"bound": rooflines[top]["bound"], "kernel": top, "achieved": achieved, "peak": rooflines[top]["peak"],
                          "unit": rooflines[top]["unit"], "frac": rooflines[top]["frac"], "traffic": traffic.get(top),
                          "peak_source": peak_src, "algorithmic": what, "avg_launch_ms": avg_ms,
                          "note": "tensor peak is the measured bf16 rate (tf32 operands run at half of it); the "
                                  "recurrence is a latency chain of T' grid-wide steps, see DESIGN.md section 6"},
                rooflines=rooflines,
                kernel_ms_per_step={k: round(v[0], 3) for k, v in sorted(per_op.items(), key=lambda kv: -kv[1][0])})
    if ragged is not None:
        line["ragged_dp"] = ragged
    if world == 1 and not args.no_isolation:
        # freed first: the isolation shapes need ~35 GB
        del model, bucket, resident, dev_buf
        torch.cuda.empty_cache()
        try:
            line["rooflines"].update(isolation_rooflines(dev, hbm_peak))
        except Exception as e:
            line["rooflines"]["isolation_error"] = f"{type(e).__name__}: {e}"[:200]
        line["torch_cudnn"] = run_torch_cudnn(dev, args.steps, args.warmup)
        try:
            line["configs2"] = configs2_leg(dev, tf_peak)
        except Exception as e:
            line["configs2"] = {"error": f"{type(e).__name__}: {e}"[:200]}
    if not args.no_cpu_baseline and world == 1:
        v, dt, cores, sample = run_cpu_port(3, 1, best_of=True)
        line["cpu_baseline"] = {"value": v, "unit": "utterance-sec/s", "cores": cores, "kind": "port", "sample": sample,
                                "extras": run_cpu_extras()}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
