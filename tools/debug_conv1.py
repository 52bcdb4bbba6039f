import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from asr_b200 import ops
from oracle import torch_path
from oracle.make_golden import synth_batch
g = torch.load('tests/golden/gru_small.pt', weights_only=False)
batch = synth_batch(g['seed'], g['B'], g['T'], g['U'], g['C'], g['lengths'])
x = batch[0]
torch.manual_seed(1)
w = torch.randn(32, 1, 41, 11) * 0.1
lens = torch.tensor([31, 24, 15], dtype=torch.int32)
for scale in (1.0, 0.018):
    for xmode in ("synth", "randn"):
        xx = x if xmode == "synth" else torch.randn_like(x)
        dy = torch.randn(3, 32, 81, 31) * scale
        dy = dy * (torch.arange(31)[None, :] < lens[:, None])[:, None, None, :]
        ref = torch.nn.grad.conv2d_weight(xx.double(), w.shape, dy.double(), stride=(2, 2), padding=(20, 5))
        got = ops.conv1_bwd_weight(xx.cuda(), dy.cuda(), tuple(w.shape), (20, 5)).cpu().double()
        err = (got - ref).abs()
        idx = err.flatten().argmax().item()
        print(f"scale {scale} x={xmode}: max err {err.max().item():.4e} ref rms {ref.pow(2).mean().sqrt().item():.4e} at {idx} "
              f"(co,kh,kw)={(idx // 451, (idx % 451) // 11, idx % 11)}; err by kw: {[round(err[:, 0, :, k].max().item(), 4) for k in range(11)]}")
