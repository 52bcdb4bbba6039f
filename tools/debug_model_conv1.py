import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from types import SimpleNamespace
import pandas as pd
from asr_b200 import ops
from asr_b200.modules import DeepSpeech
from asr_b200.trainers import CTCLoss, fit
from oracle import torch_path
from oracle.make_golden import LABELS29, synth_batch
g = torch.load('tests/golden/gru_small.pt', weights_only=False)
conf = SimpleNamespace(sample_rate=16000, window_size=0.02, window_stride=0.01, window="hamming")
res = {}
for flags in (4, 0):
    ops.set_debug_flags(flags)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "labels.csv"); pd.DataFrame({"label": LABELS29}).to_csv(path, index=False)
        model = DeepSpeech(audio_conf=conf, decoder=None, label_path=path, rnn_type="nn.GRU", rnn_hidden_size=24, rnn_hidden_layers=2)
    model.load_state_dict(torch_path.init_params("gru", 24, 2, 29), strict=True)
    model.to("cuda").train()
    batch = synth_batch(g['seed'], g['B'], g['T'], g['U'], g['C'], g['lengths'])
    _, loss, lv = fit(model, CTCLoss(), batch, "cuda")
    loss.backward(); torch.cuda.synchronize()
    res[flags] = {k: p.grad.detach().cpu().clone() for k, p in model.named_parameters()}
    print("flags", flags, "loss", lv)
for k in list(res[0])[:6]:
    a, b = res[0][k].double(), res[4][k].double()
    err = (a - b).abs()
    print(k, "max err", err.max().item(), "rms", b.pow(2).mean().sqrt().item())
k = 'conv.seq_module.0.weight'
err = (res[0][k] - res[4][k]).abs()[:, 0]
print("err by kw:", [round(err[:, :, i].max().item(), 4) for i in range(11)])
print("err by kh:", [round(err[:, i, :].max().item(), 3) for i in range(41)])
print("err by co:", [round(err[i].max().item(), 3) for i in range(32)])
