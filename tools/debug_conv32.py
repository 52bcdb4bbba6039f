import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from asr_b200 import ops
which, flags = sys.argv[1], int(sys.argv[2])
KH, KW = int(sys.argv[3]), int(sys.argv[4])
ops.set_debug_flags(flags)
torch.manual_seed(0)
B,H,W = 2, 12, 40
SH, PH, PW = 1, KH//2, KW//2
x = torch.randn(B,32,H,W); w = torch.randn(32,32,KH,KW)*0.05; b = torch.randn(32)
lens = torch.tensor([W, W-7], dtype=torch.int32)
y_ref = F.conv2d(x.double(), w.double(), b.double(), stride=(SH,1), padding=(PH,PW))
mask = (torch.arange(y_ref.shape[-1])[None,:] < lens[:,None])[:,None,None,:]
y_ref = y_ref*mask
dy = torch.randn_like(y_ref).float()*mask
xd, wd, ld = x.cuda(), w.cuda(), lens.cuda()
pf, pd = ops.conv32_pack_weights(wd)
if which == 'fwd':
    y = ops.conv32_fwd(ops.nchw_to_nhwc(xd), pf, b.cuda(), ld, tuple(w.shape), (SH,1), (PH,PW)); torch.cuda.synchronize()
    print(which, flags, KH, KW, 'maxerr', (y.cpu().double()-y_ref).abs().max().item(), 'refmax', y_ref.abs().max().item())
elif which == 'dgrad':
    ref = torch.nn.grad.conv2d_input(x.shape, w.double(), dy.double(), stride=(SH,1), padding=(PH,PW))
    dx = ops.conv32_bwd_data(ops.nchw_to_nhwc(dy.cuda()), pd, tuple(x.shape), tuple(w.shape), (SH,1), (PH,PW)); torch.cuda.synchronize()
    print(which, flags, KH, KW, 'maxerr', (dx.cpu().double()-ref).abs().max().item(), 'refmax', ref.abs().max().item())
else:
    ref = torch.nn.grad.conv2d_weight(x.double(), w.shape, dy.double(), stride=(SH,1), padding=(PH,PW))
    dw = ops.conv32_bwd_weight(xd, dy.cuda(), tuple(w.shape), (SH,1), (PH,PW)); torch.cuda.synchronize()
    print(which, flags, KH, KW, 'maxerr', (dw.cpu().double()-ref).abs().max().item(), 'refmax', ref.abs().max().item())
