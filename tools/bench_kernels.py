#!/usr/bin/env python
"""Isolation benchmarks of the HBM-bound named kernels against the measured copy bandwidth:
   CTC forward+backward at BASELINE.json configs[4] (T=2000, N=256, C=5000, U=200) and the spectrogram kernel chain
   at the configs[1] shape (64 x 10 s).  Prints one JSON line per kernel."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asr_b200 import ops

dev = "cuda"
peak = 6538.3
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p))["hbm_gbs"]


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


def ctc(T=2000, N=256, C=5000, U=200):
    g = torch.Generator(device=dev).manual_seed(1239)
    lp = (torch.randn(T, N, C, device=dev, generator=g) * 3).log_softmax(2)
    tg = torch.randint(1, C, (N * U,), device=dev, generator=g, dtype=torch.int32)
    il = torch.full((N,), T, dtype=torch.int32, device=dev)
    tl = torch.full((N,), U, dtype=torch.int32, device=dev)
    one = torch.ones(1, device=dev)
    state = {}

    def fwd():
        state["f"] = ops.ctc_fwd(lp, tg, il, tl, U)

    def bwd():
        loss, nll, alpha = state["f"]
        state["g"] = ops.ctc_bwd(lp, tg, il, tl, alpha, nll, one, U)

    def both():   # the backward consumes the alpha workspace: always time forward + backward pairs
        fwd()
        bwd()

    tf = timed(fwd)
    tb = timed(both) - tf
    # CPU reference of the same call (torch.nn.functional.ctc_loss, the reference's criterion) on a slice of the batch
    ns = 16
    lpc = lp[:, :ns].float().cpu().requires_grad_(True)
    t0 = time.perf_counter()
    l = torch.nn.functional.ctc_loss(lpc, tg[: ns * U].cpu(), il[:ns].cpu(), tl[:ns].cpu(), reduction="sum")
    l.backward()
    tcpu = (time.perf_counter() - t0) * N / ns
    loss = state["f"][0].item()
    ref = torch.nn.functional.ctc_loss(lp[:, :ns].float().cpu(), tg[: ns * U].cpu(), il[:ns].cpu(), tl[:ns].cpu(), reduction="sum").item()
    mine = state["f"][1][:ns].sum().item()
    gerr = (state["g"][:, :ns].cpu() - lpc.grad).abs().max().item()
    alg = 2.0 * T * N * C * 4
    print(json.dumps({"kernel": "ctc fwd+bwd", "shape": f"T={T} N={N} C={C} U={U}", "fwd_ms": tf, "bwd_ms": tb,
                      "algorithmic_GB": alg / 1e9, "achieved_GBs": alg / ((tf + tb) * 1e-3) / 1e9,
                      "frac_of_measured_hbm": alg / ((tf + tb) * 1e-3) / 1e9 / peak, "peak_GBs": peak,
                      "loss_rel_err_vs_torch_cpu": abs(mine - ref) / abs(ref), "grad_max_abs_err": gerr,
                      "cpu_torch_ms_scaled": tcpu * 1e3, "cpu_cores": os.cpu_count(),
                      "utt_sec_per_s": N * T * 0.02 / ((tf + tb) * 1e-3)}), flush=True)


def stft(B=64, S=160000):
    import scipy.signal
    g = torch.Generator(device=dev).manual_seed(7)
    wav = torch.randn(B, S, device=dev, generator=g) * 0.1
    ns = torch.full((B,), S, dtype=torch.int32, device=dev)
    win = torch.from_numpy(scipy.signal.get_window("hamming", 320, fftbins=True)).float().to(dev)
    basis = ops.dft_basis(320, dev)
    t = timed(lambda: ops.spectrogram(wav, ns, win, basis, 320, 160, True))
    alg = B * (4.0 * S + 4.0 * 161 * (1 + S // 160))
    print(json.dumps({"kernel": "spectrogram (frames + 3xTF32 DFT GEMM + log1p|.| + normalise)", "shape": f"B={B} S={S}",
                      "ms": t, "algorithmic_GB": alg / 1e9, "achieved_GBs": alg / (t * 1e-3) / 1e9,
                      "frac_of_measured_hbm": alg / (t * 1e-3) / 1e9 / peak, "utt_sec_per_s": B * S / 16000 / (t * 1e-3)}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["ctc", "stft"]
    if "ctctune" in which:
        from asr_b200 import _lib
        for bits in (0, 4):
            _lib.query("asrb_debug_ctc_dbg", bits)
            print("tuning", bits, 0, end=" ")
            ctc()
        _lib.query("asrb_debug_ctc_dbg", 0)
    if "ctc" in which:
        ctc()
    if "stft" in which:
        stft()
