"""TEST INFRASTRUCTURE -- generates tests/golden/*.pt by running the UNMODIFIED reference.

Run in the build container only:   python -m oracle.make_golden
(needs /root/reference; the GPU box never runs this).  Outputs are small: model weights and
inputs are regenerated from seeds by ``oracle.torch_path.init_params`` / ``synth_batch``
(their checksums are stored so a drifting RNG stream is detected, not silently accepted);
what is stored are the reference's OUTPUTS -- loss, logits, argmax indices, BN running
statistics, and for every parameter gradient its L2 norm, sum and 64 sampled entries.

Reference entry points exercised (all imported unmodified, see reference_loader.py):
  asr_deepspeech.modules.deepspeech.DeepSpeech.{__init__,forward,get_seq_lens}
  asr_deepspeech.modules.blocks.{MaskConv,BatchRNN,SequenceWise,InferenceBatchSoftmax}
  asr_deepspeech.functional.{check_loss,_collate_fn}
  asr_deepspeech.decoders.greedy_decoder.GreedyDecoder.decode
  the arithmetic of trainers/deepspeech_trainer.py:104-112 with torch.nn.CTCLoss(reduction="sum")
"""
from __future__ import annotations

import os
import sys
import tempfile
from types import SimpleNamespace

import numpy as np
import torch

from . import torch_path
from .reference_loader import load_reference

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# 29 symbols without a bare space (pandas drops it)
LABELS29 = ["_", "'"] + list("abcdefghijklmnopqrstuvwxyz") + ["#"]


def synth_batch(seed, B, T, U, C, lengths=None):
    """Synthetic batch in the layout _collate_fn produces (functional.py:9-32): lengths sorted
    descending, tail zero-padded, targets concatenated, uniform in [1, C-1]."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 1, 161, T, generator=g)
    if lengths is None:
        lengths = [T] * B
    lengths = sorted(lengths, reverse=True)
    for i, l in enumerate(lengths):
        x[i, :, :, l:] = 0
    pct = torch.tensor([l / float(T) for l in lengths], dtype=torch.float32)
    if isinstance(U, int):
        U = [U] * B
    tsz = torch.tensor(U, dtype=torch.int32)
    tgt = torch.randint(1, C, (int(tsz.sum()),), generator=g, dtype=torch.int64).to(torch.int32)
    return x, tgt, pct, tsz


def checksum(t: torch.Tensor):
    t = t.double()
    return torch.stack([t.sum(), (t * t).sum(), t.flatten()[:: max(1, t.numel() // 7)][:7].sum()])


def sample_idx(numel, k=64, seed=7):
    g = torch.Generator().manual_seed(seed + numel)
    return torch.randint(0, numel, (k,), generator=g)


def grad_digest(g: torch.Tensor):
    flat = g.flatten()
    return dict(norm=flat.double().norm().item(), sum=flat.double().sum().item(),
                samples=flat[sample_idx(flat.numel())].clone())


def _audio_conf():
    return SimpleNamespace(sample_rate=16000, window_size=0.02, window_stride=0.01, window="hamming",
                           speed_volume_perturb=False, spec_augment=False, noise_dir=None,
                           noise_prob=0.4, noise_levels=(0.0, 0.5))


def build_reference_model(rnn_type, hidden, layers, labels, seed=123456):
    import pandas as pd
    from asr_deepspeech.modules.deepspeech import DeepSpeech

    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "labels.csv")
        pd.DataFrame({"label": labels}).to_csv(path, index=False)
        torch.manual_seed(seed)
        model = DeepSpeech(audio_conf=_audio_conf(), decoder=None, label_path=path,
                           rnn_type=f"nn.{rnn_type.upper()}", rnn_hidden_size=hidden,
                           rnn_hidden_layers=layers, bidirectional=True)
    return model


def reference_fit(model, batch):
    """trainers/deepspeech_trainer.py:104-112 verbatim semantics on the reference model."""
    inputs, targets, input_percentages, target_sizes = batch
    criterion = torch.nn.CTCLoss(reduction="sum")                    # trainers/__main__.py:53
    input_sizes = input_percentages.clone().mul_(int(inputs.size(3))).int()
    out, output_sizes = model.forward(inputs, input_sizes)
    out.retain_grad()
    out_t = out.transpose(0, 1)
    float_out = out_t.float().log_softmax(2)
    loss = criterion(float_out, targets, output_sizes, target_sizes)
    loss = loss / inputs.size(0)
    return loss, out, output_sizes


def model_case(name, rnn_type, hidden, layers, C, seed, B, T, U, lengths=None):
    labels = LABELS29[:C] if C <= 29 else [chr(0x3041 + i) for i in range(C)]
    model = build_reference_model(rnn_type, hidden, layers, labels)
    assert model.num_classes == C, (model.num_classes, C)
    batch = synth_batch(seed, B, T, U, C, lengths)
    p = torch_path.init_params(rnn_type, hidden, layers, C)
    sd = model.state_dict()
    assert list(sd.keys()) == list(p.keys()), "state_dict key order differs from init_params"
    for k in sd:
        assert torch.equal(sd[k], p[k]), f"init_params != reference init at {k}"
    model.train()
    loss, out, output_sizes = reference_fit(model, batch)
    model.zero_grad()
    loss.backward()
    grads = {k: grad_digest(v.grad) for k, v in model.named_parameters()}
    stats = {k: v.clone() for k, v in model.state_dict().items() if "running_" in k}
    model.eval()
    with torch.no_grad():
        input_sizes = batch[2].clone().mul_(int(batch[0].size(3))).int()
        probs, _ = model.forward(batch[0], input_sizes)
        idx = torch.max(probs, 2)[1]                                  # greedy_decoder.py:61
        strings, offsets = model.decoder.decode(probs, output_sizes)
    rec = dict(
        name=name, rnn_type=rnn_type, hidden=hidden, layers=layers, C=C, seed=seed, B=B, T=T,
        U=U, lengths=lengths,
        param_checksums={k: checksum(v) for k, v in p.items() if v.is_floating_point()},
        input_checksum=checksum(batch[0]), targets=batch[1], input_percentages=batch[2],
        target_sizes=batch[3], output_sizes=output_sizes, loss=loss.detach(),
        logits=out.detach().clone(), dlogits=out.grad.detach().clone(), grads=grads,
        running_stats=stats, eval_argmax=idx, eval_probs_digest=grad_digest(probs),
        eval_strings=[s[0] for s in strings],
    )
    rec.update(eval64_fields(model, batch))
    torch.save(rec, os.path.join(GOLDEN_DIR, f"{name}.pt"))
    print(f"{name}: loss={loss.item():.6f} out={tuple(out.shape)} sizes={output_sizes.tolist()}")


def model_case_big(name, rnn_type, hidden, layers, C, seed, B, T, U, backward=True):
    """Full-size BASELINE.json configurations (configs[1], configs[2]) on the unmodified reference.  The tensors are too
    large to commit, so what is stored is: loss, output sizes, digests (norm / sum / 64 samples) of logits, d loss / d
    logits and every parameter gradient, the BN running statistics, and -- for the bit-exact index check -- the
    eval-mode argmax of every frame (int16) with the reference's own top-2 margin per frame (float16)."""
    labels = LABELS29[:C] if C <= 29 else [chr(0x3041 + i) for i in range(C)]
    model = build_reference_model(rnn_type, hidden, layers, labels)
    assert model.num_classes == C, (model.num_classes, C)
    batch = synth_batch(seed, B, T, U, C)
    p = torch_path.init_params(rnn_type, hidden, layers, C)
    sd = model.state_dict()
    for k in sd:
        assert torch.equal(sd[k], p[k]), f"init_params != reference init at {k}"
    import time
    t0 = time.time()
    model.train()
    rec = dict(name=name, rnn_type=rnn_type, hidden=hidden, layers=layers, C=C, seed=seed, B=B, T=T, U=U,
               input_checksum=checksum(batch[0]), targets=batch[1], input_percentages=batch[2], target_sizes=batch[3])
    if backward:
        loss, out, output_sizes = reference_fit(model, batch)
        model.zero_grad()
        loss.backward()
        rec.update(grads={k: grad_digest(v.grad) for k, v in model.named_parameters()},
                   dlogits_digest=grad_digest(out.grad.detach()))
    else:
        loss, out, output_sizes = reference_fit(model, batch)     # forward only: the graph is built but never walked
        loss, out = loss.detach(), out.detach()
    rec.update(loss=loss.detach(), output_sizes=output_sizes, logits_digest=grad_digest(out.detach()),
               running_stats={k: v.clone() for k, v in model.state_dict().items() if "running_" in k})
    print(f"{name}: train pass {time.time() - t0:.0f} s, loss={loss.item():.6f}", flush=True)
    model.eval()
    with torch.no_grad():
        input_sizes = batch[2].clone().mul_(int(batch[0].size(3))).int()
        probs, _ = model.forward(batch[0], input_sizes)
        top2 = torch.topk(probs, 2, dim=2)
        rec.update(eval_argmax=torch.max(probs, 2)[1].to(torch.int16),               # greedy_decoder.py:61
                   eval_margin=(top2.values[..., 0] - top2.values[..., 1]).to(torch.float16),
                   eval_probs_digest=grad_digest(probs))
    rec.update(eval64_fields(model, batch))
    torch.save(rec, os.path.join(GOLDEN_DIR, f"{name}.pt"))
    print(f"{name}: done {time.time() - t0:.0f} s out={tuple(out.shape)}", flush=True)


def eval64_fields(model, batch):
    """The eval-mode forward of the SAME unmodified reference modules in fp64 (BN on the running statistics the model
    holds): per frame the two most probable classes and the log-probability margin between them.  tests use it to
    settle every frame where a GPU argmax differs from the reference's fp32 argmax: either the GPU picked the fp64
    winner (the fp32 reference is the one on the wrong side), or the fp64 margin is below the arithmetic noise."""
    import copy

    m64 = copy.deepcopy(model).double().eval()
    with torch.no_grad():
        input_sizes = batch[2].clone().mul_(int(batch[0].size(3))).int()
        # eval mode has no cross-utterance arithmetic (BN on running statistics): 4 utterances at a time keeps the fp64
        # convolution's workspace small; the time axis of a chunk is the full padded one, as in the one-batch call
        outs = []
        for i in range(0, batch[0].size(0), 4):
            pr, _ = m64.forward(batch[0][i:i + 4].double(), input_sizes[i:i + 4])
            outs.append(pr)
        probs = torch.cat(outs, 0)
        lp = probs.log()
        top2 = torch.topk(lp, 2, dim=2)
    return dict(eval64_top2=top2.indices.to(torch.int16),
                eval64_margin=(top2.values[..., 0] - top2.values[..., 1]).float(),
                eval64_logp_digest=grad_digest(lp.float()))


def augment_eval64(name):
    """add the fp64 eval fields to an existing golden file (weights = the seeded init, BN statistics = the stored ones)"""
    path = os.path.join(GOLDEN_DIR, f"{name}.pt")
    rec = torch.load(path)
    C = rec["C"]
    labels = LABELS29[:C] if C <= 29 else [chr(0x3041 + i) for i in range(C)]
    model = build_reference_model(rec["rnn_type"], rec["hidden"], rec["layers"], labels)
    sd = model.state_dict()
    for k, v in rec["running_stats"].items():
        sd[k].copy_(v)
    batch = synth_batch(rec["seed"], rec["B"], rec["T"], rec["U"], C, rec.get("lengths"))
    assert torch.equal(checksum(batch[0]), rec["input_checksum"])
    model.eval()
    with torch.no_grad():       # the stored fp32 argmax is reproduced before anything is added
        input_sizes = batch[2].clone().mul_(int(batch[0].size(3))).int()
        probs, _ = model.forward(batch[0], input_sizes)
    assert torch.equal(torch.max(probs, 2)[1].to(rec["eval_argmax"].dtype), rec["eval_argmax"]), "fp32 eval argmax not reproduced"
    rec.update(eval64_fields(model, batch))
    torch.save(rec, path)
    flips = (rec["eval64_top2"][..., 0].long() != rec["eval_argmax"].long()).sum().item()
    print(f"{name}: fp64 eval fields added; the reference's own fp32 argmax differs from fp64 in {flips} frames", flush=True)


def ctc_cases():
    """torch.nn.CTCLoss(reduction='sum') exactly as constructed at trainers/__main__.py:53 and
    called at trainers/deepspeech_trainer.py:110-111 (log_softmax first)."""
    crit = torch.nn.CTCLoss(reduction="sum")
    cases = []
    g = torch.Generator().manual_seed(99)
    specs = [
        dict(T=12, N=3, C=6, in_len=[12, 9, 5], tgt=[[1, 2, 3], [2, 2], [5]]),        # repeats
        dict(T=7, N=2, C=4, in_len=[7, 7], tgt=[[1, 1, 1], [3, 2, 1]]),               # 1,1,1 needs T>=5
        dict(T=5, N=2, C=5, in_len=[5, 3], tgt=[[1, 2], [2, 2, 2]]),                  # 2nd infeasible -> inf
        dict(T=6, N=2, C=3, in_len=[6, 4], tgt=[[], [1]]),                            # empty target
        dict(T=40, N=4, C=29, in_len=[40, 33, 21, 20], tgt=None, U=[10, 9, 10, 3]),
        dict(T=30, N=2, C=90, in_len=[30, 30], tgt=None, U=[14, 15]),
    ]
    for sp in specs:
        T, N, C = sp["T"], sp["N"], sp["C"]
        logits = torch.randn(T, N, C, generator=g) * 2
        if sp["tgt"] is None:
            tl = sp["U"]
            tgt = [torch.randint(1, C, (u,), generator=g).tolist() for u in tl]
        else:
            tgt = sp["tgt"]
        flat = torch.tensor([c for t in tgt for c in t], dtype=torch.int32)
        tsz = torch.tensor([len(t) for t in tgt], dtype=torch.int32)
        isz = torch.tensor(sp["in_len"], dtype=torch.int32)
        x = logits.clone().requires_grad_(True)
        lp = x.float().log_softmax(2)
        lp.retain_grad()
        loss = crit(lp, flat, isz, tsz)
        per = torch.nn.functional.ctc_loss(lp.detach(), flat, isz, tsz, reduction="none")
        if torch.isfinite(loss):
            loss.backward()
            gl, glp = x.grad.clone(), lp.grad.clone()
        else:
            gl = glp = None
        cases.append(dict(logits=logits, targets=flat, input_lengths=isz, target_lengths=tsz,
                          loss=loss.detach(), nll=per, grad_logits=gl, grad_log_probs=glp))
        print("ctc case", T, N, C, "loss", loss.item())
    torch.save(cases, os.path.join(GOLDEN_DIR, "ctc_cases.pt"))


def misc_cases():
    from asr_deepspeech.functional import _collate_fn, check_loss
    from asr_deepspeech.modules.blocks import MaskConv

    model = build_reference_model("gru", 16, 1, LABELS29[:26])
    lens = torch.arange(1, 2100, dtype=torch.int32)
    seq = model.get_seq_lens(lens)
    # MaskConv value behaviour (tests/test_blocks_mask.py:6-14)
    torch.manual_seed(5)
    conv = torch.nn.Conv2d(1, 2, kernel_size=3, padding=1)
    mc = MaskConv(torch.nn.Sequential(conv))
    x = torch.randn(2, 1, 8, 10)
    y, _ = mc(x, torch.tensor([10, 4]))
    # collate ordering (tests/test_spectrogram_dataset.py:163-175)
    batch = [(torch.ones(161, 5), [1, 2, 3]), (torch.ones(161, 10) * 2, [4, 5])]
    inputs, targets, pct, tsz = _collate_fn(batch)
    cl = []
    for v in (1.5, float("inf"), float("-inf"), float("nan"), -0.5, 0.0):
        ok, err = check_loss(torch.tensor(v), v)
        cl.append((v, ok, err))
    torch.save(dict(seq_lens_in=lens, seq_lens_out=seq,
                    maskconv=dict(weight=conv.weight.detach(), bias=conv.bias.detach(), x=x, y=y.detach()),
                    collate=dict(inputs_sum=inputs.sum(), inputs_shape=tuple(inputs.shape), targets=targets,
                                 pct=pct, tsz=tsz),
                    check_loss=cl), os.path.join(GOLDEN_DIR, "misc.pt"))
    print("misc: seq_lens sample", seq[[100, 1000, 1500, 2000]].tolist())


def reference_checkpoint(model, optimizer, scheduler=None, epoch=3, metrics=None):
    """the dict DeepSpeechTrainer.save() writes (trainers/deepspeech_trainer.py:176-188)"""
    return {"epoch": epoch, "metrics": metrics if metrics is not None else {"wer": 0.5, "cer": 0.25},
            "optimizer": optimizer.state_dict(), "scheduler": scheduler, "state_dict": model.state_dict()}


def reference_optimizer(model):
    """trainers/__main__.py:41-47 with config.yml:41-47"""
    return torch.optim.AdamW(model.parameters(), lr=1.5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5)


def fake_grads(model, seed):
    g = torch.Generator().manual_seed(seed)
    for p in model.parameters():
        p.grad = torch.randn(p.shape, generator=g) * 0.01


def checkpoint_manifest():
    """Structure of a checkpoint written from the unmodified reference model after two AdamW steps: parameter order,
    state_dict keys/shapes/dtypes, optimizer state layout -- a few KB instead of the 3 MB file itself
    (tests/test_host_logic.py regenerates the real file from /root/reference when it is present)."""
    model = build_reference_model("gru", 8, 2, LABELS29[:26])
    opt = reference_optimizer(model)
    for s in (1, 2):
        fake_grads(model, s)
        opt.step()
    ck = reference_checkpoint(model, opt)
    osd = ck["optimizer"]
    man = dict(ckpt_keys=sorted(ck.keys()),
               param_order=[k for k, _ in model.named_parameters()],
               state_dict=[(k, tuple(v.shape), str(v.dtype)) for k, v in ck["state_dict"].items()],
               opt_group_keys=sorted(k for k in osd["param_groups"][0].keys()),
               opt_group_params=list(osd["param_groups"][0]["params"]),
               opt_state_keys=sorted(osd["state"][0].keys()),
               opt_step_type=type(osd["state"][0]["step"]).__name__,
               digest={k: checksum(v.float()) for k, v in ck["state_dict"].items() if v.dtype.is_floating_point})
    torch.save(man, os.path.join(GOLDEN_DIR, "ref_checkpoint_manifest.pt"))
    print("checkpoint manifest:", len(man["param_order"]), "parameters,", len(man["state_dict"]), "state_dict entries")


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    load_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "checkpoint":
        checkpoint_manifest()
        return
    if len(sys.argv) > 2 and sys.argv[1] == "eval64":
        torch.set_num_threads(os.cpu_count())
        for name in sys.argv[2:]:
            augment_eval64(name)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "cfg2":
        # BASELINE.json configs[1] = the benchmarked shape (bench.py CFG: seed 1234+2), fwd + bwd once (~6 min of CPU)
        torch.set_num_threads(os.cpu_count())
        model_case_big("cfg2_gru800x5", "gru", 800, 5, 29, seed=1236, B=64, T=1001, U=100)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "cfg3":
        # BASELINE.json configs[2]: 7 x biLSTM-1024, 15 s, 90 labels; batch reduced 128 -> 16 for the CPU run (stated), fwd + bwd
        torch.set_num_threads(os.cpu_count())
        model_case_big("cfg3_lstm1024x7_b16", "lstm", 1024, 7, 90, seed=1237, B=16, T=1501, U=150, backward=True)
        return
    torch.set_num_threads(os.cpu_count())
    misc_cases()
    ctc_cases()
    model_case("gru_small", "gru", 24, 2, 29, seed=11, B=3, T=61, U=[5, 4, 3], lengths=[61, 47, 30])
    model_case("lstm_small", "lstm", 24, 2, 29, seed=12, B=3, T=61, U=[5, 4, 3], lengths=[61, 48, 29])
    model_case("lstm_c90", "lstm", 16, 3, 90, seed=13, B=2, T=41, U=[6, 2], lengths=[41, 22])
    # BASELINE.json configs[0]: 2-conv + 5 x biGRU-800, batch 4, 1 s, 29 labels (seed 1234+1)
    model_case("cfg1_gru800x5", "gru", 800, 5, 29, seed=1235, B=4, T=101, U=10)


if __name__ == "__main__":
    main()
