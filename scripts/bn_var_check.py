"""Debug aid: accuracy of the forward BatchNorm sums (var = E[x^2] - mean^2) for near-constant channels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from simclr_b200 import _lib
st = _lib.stream_ptr()
def mk(path):
    l = _lib._Lib(); keep = _lib.LIB_PATH; _lib.LIB_PATH = path; l.load(); _lib.LIB_PATH = keep; return l
libs = {'main': mk(os.path.join(ROOT, 'simclr_b200', 'libsimclr_b200.so'))}
for n in ():   # side builds for A/B runs go here
    p = os.path.join(ROOT, 'build_tmp', 'libsimclr_%s.so' % n)
    if os.path.exists(p): libs[n] = mk(p)
torch.manual_seed(0)
for rows, C in [(16384, 32), (4096, 64), (64, 2048), (16, 2048), (16, 32)]:
    for mean, sd in [(0.0, 1.0), (3.0, 1e-2), (3.0, 1e-3), (1.0, 0.0)]:
        x = (mean + sd * torch.randn(rows, C, device='cuda')).float()
        xd = x.double()
        m = xd.mean(0); var = ((xd - m) ** 2).mean(0)
        line = 'rows=%6d C=%5d mean=%g sd=%g |' % (rows, C, mean, sd)
        for name, lib in libs.items():
            s = torch.empty(2 * C, dtype=torch.float64, device='cuda')
            lib.bn_stats(x, 0, rows, C, s, st)
            torch.cuda.synchronize()
            mk_ = s[:C] / rows; vk = (s[C:] / rows - mk_ * mk_).clamp(min=0)
            r_true = torch.rsqrt(var + 1e-5); r_k = torch.rsqrt(vk + 1e-5)
            line += ' %s: rstd err max %.1e' % (name, float(((r_k - r_true) / r_true).abs().max()))
        print(line, flush=True)
